/* Prints the layout of the structs of include/gse.h as the C compiler sees it; tests/test_abi.py
 * compares it with the ctypes mirrors in gpu_se_b200/_lib.py. */
#include <stddef.h>
#include <stdio.h>
#include "gse.h"

#define F(S, f) printf("\"%s.%s\": [%zu, %zu],\n", #S, #f, offsetof(S, f), sizeof(((S*)0)->f))

int main(void) {
    printf("{\n");
    printf("\"sizeof.gse_mixture\": %zu,\n", sizeof(gse_mixture));
    F(gse_mixture, nd); F(gse_mixture, nx); F(gse_mixture, weights); F(gse_mixture, means); F(gse_mixture, covs);
    printf("\"sizeof.gse_shards\": %zu,\n", sizeof(gse_shards));
    F(gse_shards, nshards); F(gse_shards, rows); F(gse_shards, cumsum_dev); F(gse_shards, state_dev);
    F(gse_shards, ld); F(gse_shards, offsets_dev); F(gse_shards, idx_dev); F(gse_shards, rank);
    printf("\"sizeof.gse_step_params\": %zu,\n", sizeof(gse_step_params));
    F(gse_step_params, u); F(gse_step_params, dt); F(gse_step_params, z); F(gse_step_params, r);
    F(gse_step_params, step); F(gse_step_params, reserved);
    printf("\"GSE_ABI_VERSION\": %d, \"GSE_MAX_SHARDS\": %d, \"GSE_MAX_ND\": %d, \"GSE_MAILBOX_BYTES\": %d,\n",
           GSE_ABI_VERSION, GSE_MAX_SHARDS, GSE_MAX_ND, (int)GSE_MAILBOX_BYTES);
    printf("\"GSE_IPC_HANDLE_BYTES\": %d\n}\n", GSE_IPC_HANDLE_BYTES);
    return 0;
}
