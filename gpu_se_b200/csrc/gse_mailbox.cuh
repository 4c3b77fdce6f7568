// Peer mailboxes: all-gather of one small record per shard over NVLink peer memory, inside whatever kernel needs it
// (no NCCL launch, no host involvement).  Shard t's mailbox is GSE_MAILBOX_BYTES of gse_peer_alloc memory that every
// rank has mapped: two parities x GSE_MAX_SHARDS slots of MBOX_SLOT_BYTES.  An exchange with sequence number `epoch`
// (identical on every rank, never 0, +1 per exchange) uses the slots of parity epoch & 1: slot [rank] of every peer's
// mailbox receives this rank's record (payload first, system-scope fence, then the epoch as the flag); a rank reads
// slot [t] of its OWN mailbox once the flag shows `epoch`.  A peer can only reuse a slot two exchanges later, which
// needs this rank's flag of the exchange in between, written after it has read -- so no slot is overwritten unread.
// The wait is bounded (GSE_MBOX_TIMEOUT_NS of %globaltimer): a rank that died or issued its collectives in another
// order makes the others set GSE_ERR_PEER_TIMEOUT in their error word instead of hanging the GPU.
#pragma once

#include "gse_common.cuh"

#define MBOX_SLOT_BYTES 512
#define MBOX_FLAG_WORD 60                 // payload: words [0, 60) of 8 bytes; flag (the epoch) in word 60
#define GSE_MBOX_TIMEOUT_NS 4000000000ull

struct MailboxTable {
    unsigned char* box[GSE_MAX_SHARDS];
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long mbox_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ volatile unsigned long long* mbox_slot(const MailboxTable& mb, int owner, int writer,
                                                                  unsigned int epoch) {
    return (volatile unsigned long long*)(mb.box[owner] + ((size_t)(epoch & 1u) * GSE_MAX_SHARDS + writer) * MBOX_SLOT_BYTES);
}
// spin until peer `t`'s record of `epoch` has landed in this rank's mailbox; false (and the error bit) on time-out
__device__ __forceinline__ bool mbox_wait(const MailboxTable& mb, int rank, int t, unsigned int epoch, unsigned int* err) {
    volatile unsigned long long* src = mbox_slot(mb, rank, t, epoch);
    const unsigned long long t0 = mbox_timer_ns();
    unsigned int spins = 0;
    while (*(volatile unsigned int*)(src + MBOX_FLAG_WORD) != epoch) {
        __nanosleep(40);
        if ((++spins & 1023u) == 0u && mbox_timer_ns() - t0 > GSE_MBOX_TIMEOUT_NS) {
            if (err) atomicOr(err, GSE_ERR_PEER_TIMEOUT);
            return false;
        }
    }
    __threadfence_system();
    return true;
}

// record of two words, one warp: lane t < nshards talks to peer t
__device__ __forceinline__ void mbox_exchange(const MailboxTable& mb, int rank, int nshards, unsigned int epoch,
                                              unsigned long long w0, unsigned long long w1, int lane,
                                              unsigned long long& r0, unsigned long long& r1, unsigned int* err) {
    r0 = 0; r1 = 0;
    if (lane < nshards) {
        volatile unsigned long long* dst = mbox_slot(mb, lane, rank, epoch);
        dst[0] = w0;
        dst[1] = w1;
        __threadfence_system();
        *(volatile unsigned int*)(dst + MBOX_FLAG_WORD) = epoch;
        if (mbox_wait(mb, rank, lane, epoch, err)) {
            volatile unsigned long long* src = mbox_slot(mb, rank, lane, epoch);
            r0 = src[0];
            r1 = src[1];
        }
    }
}

// barrier without a payload: "everything I wrote before my system-scope fence is out" / wait until every peer has said
// the same.  The CALLER issues the one __threadfence_system() before it; nothing is read afterwards, so no acquire fence.
__device__ __forceinline__ void mbox_signal_wait(const MailboxTable& mb, int rank, int nshards, unsigned int epoch, int lane,
                                                 unsigned int* err) {
    if (lane < nshards) {
        *(volatile unsigned int*)(mbox_slot(mb, lane, rank, epoch) + MBOX_FLAG_WORD) = epoch;
        volatile unsigned long long* src = mbox_slot(mb, rank, lane, epoch);
        const unsigned long long t0 = mbox_timer_ns();
        unsigned int spins = 0;
        while (*(volatile unsigned int*)(src + MBOX_FLAG_WORD) != epoch) {
            __nanosleep(20);
            if ((++spins & 1023u) == 0u && mbox_timer_ns() - t0 > GSE_MBOX_TIMEOUT_NS) {
                if (err) atomicOr(err, GSE_ERR_PEER_TIMEOUT);
                break;
            }
        }
    }
}

// one 64-bit value per shard WITHOUT a fence: the value travels as two words that carry the epoch in their upper halves
// (slot words 61 and 62, written by nothing else), each a single 8-byte store -- a reader that sees the epoch in both has
// the value.  Saves the two system-scope fences of the generic exchange on the critical path of the fused resample.
#define MBOX_TAGGED_WORD 61
__device__ __forceinline__ void mbox_exchange_tagged(const MailboxTable& mb, int rank, int nshards, unsigned int epoch,
                                                     unsigned long long value, int lane, unsigned long long& r0,
                                                     unsigned int* err) {
    r0 = 0;
    if (lane < nshards) {
        volatile unsigned long long* dst = mbox_slot(mb, lane, rank, epoch) + MBOX_TAGGED_WORD;
        const unsigned long long tag = (unsigned long long)epoch << 32;
        dst[0] = tag | (value & 0xffffffffull);
        dst[1] = tag | (value >> 32);
        volatile unsigned long long* src = mbox_slot(mb, rank, lane, epoch) + MBOX_TAGGED_WORD;
        const unsigned long long t0 = mbox_timer_ns();
        unsigned int spins = 0;
        unsigned long long a = src[0], b = src[1];
        while ((unsigned int)(a >> 32) != epoch || (unsigned int)(b >> 32) != epoch) {
            __nanosleep(20);
            if ((++spins & 1023u) == 0u && mbox_timer_ns() - t0 > GSE_MBOX_TIMEOUT_NS) {
                if (err) atomicOr(err, GSE_ERR_PEER_TIMEOUT);
                return;
            }
            a = src[0];
            b = src[1];
        }
        r0 = (b << 32) | (a & 0xffffffffull);
    }
}

// record of nwords <= MBOX_FLAG_WORD words, one warp: the warp writes the record to every peer in turn, then lane t
// waits for peer t.  Afterwards the records of all shards sit in this rank's own mailbox (mbox_slot(mb, rank, t, epoch)).
__device__ __forceinline__ void mbox_exchange_block(const MailboxTable& mb, int rank, int nshards, unsigned int epoch,
                                                    const unsigned long long* rec, int nwords, int lane, unsigned int* err) {
    for (int t = 0; t < nshards; ++t) {
        volatile unsigned long long* dst = mbox_slot(mb, t, rank, epoch);
        for (int k = lane; k < nwords; k += 32) dst[k] = rec[k];
    }
    __threadfence_system();
    __syncwarp();
    if (lane < nshards) *(volatile unsigned int*)(mbox_slot(mb, lane, rank, epoch) + MBOX_FLAG_WORD) = epoch;
    if (lane < nshards) mbox_wait(mb, rank, lane, epoch, err);
    __syncwarp();
}
#endif

static inline int gse_build_mailboxes(void* const boxes[GSE_MAX_SHARDS], int rank, int nshards, MailboxTable* mb) {
    GSE_REQUIRE(boxes != NULL && nshards >= 1 && nshards <= GSE_MAX_SHARDS && rank >= 0 && rank < nshards,
                "bad mailbox arguments");
    memset(mb, 0, sizeof(*mb));
    for (int t = 0; t < nshards; ++t) {
        GSE_REQUIRE(boxes[t] != NULL, "mailbox pointer is NULL");
        mb->box[t] = (unsigned char*)boxes[t];
    }
    return GSE_OK;
}
