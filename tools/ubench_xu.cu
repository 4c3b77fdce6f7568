// Micro-benchmark: issue throughput of the XU-pipe operations the kernels use (MUFU flavours and
// the float64 / 64-bit conversions), in lane-ops per clock per SM.  Build + run:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_xu tools/ubench_xu.cu && /tmp/ubench_xu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(float* out, float seed, unsigned long long useed) {
    float x[ILP];
    double d[ILP];
    unsigned long long u[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = seed + threadIdx.x * 1e-3f + i; d[i] = x[i]; u[i] = useed + threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 2) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 3) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 4) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 6) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 7) { asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(d[i]) : "l"(u[i])); u[i] += (unsigned long long)__double_as_longlong(d[i]) & 1; }
            if (OP == 8) asm volatile("cvt.rmi.f64.f64 %0, %0;" : "+d"(d[i]));
            if (OP == 9) { int r; asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(r) : "d"(d[i])); d[i] += r; }
            if (OP == 10) { unsigned long long r; asm volatile("cvt.rni.u64.f32 %0, %1;" : "=l"(r) : "f"(x[i])); x[i] += (float)(r & 1); }
            if (OP == 11) { unsigned r; asm volatile("cvt.rni.u32.f32 %0, %1;" : "=r"(r) : "f"(x[i])); x[i] += __uint_as_float(r & 1); }
            if (OP == 12) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            if (OP == 13) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(d[i]));
            if (OP == 14) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x[i]) : "r"(__float_as_uint(x[i]))); }
            if (OP == 15) asm volatile("div.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % ILP]));
            if (OP == 16) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(x[i])); x[i] += __int_as_float(__double2loint(d[i]) & 1); }
            if (OP == 17) asm volatile("add.rn.f64 %0, %0, %0;" : "+d"(d[i]));
            if (OP == 18) asm volatile("cvt.rni.f32.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 19) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x[i]) : "d"(d[i])); d[i] += __longlong_as_double((long long)(__float_as_uint(x[i]) & 1)); }
            if (OP == 20) { unsigned long long r; asm volatile("cvt.rni.u64.f64 %0, %1;" : "=l"(r) : "d"(d[i])); d[i] += __longlong_as_double((long long)(r & 1)); }
            if (OP == 21) { int r; asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(r) : "f"(x[i])); x[i] += __int_as_float(r & 1); }
        }
    }
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i] + (float)d[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms, float clk_ghz, float* out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(out, 1.5f, 3ull);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(out, 1.5f, 3ull);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * ITERS * ILP;
    printf("%-22s %8.3f ms  %7.1f lane-ops/clk/SM (at %.3f GHz)\n", name, ms, ops / (ms * 1e-3) / (clk_ghz * 1e9) / sms, clk_ghz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz * 1e-6f;
    float* out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<0>("mufu.ex2", p.multiProcessorCount, ghz, out);
    run<1>("mufu.lg2", p.multiProcessorCount, ghz, out);
    run<2>("mufu.sin", p.multiProcessorCount, ghz, out);
    run<3>("mufu.cos", p.multiProcessorCount, ghz, out);
    run<4>("mufu.sqrt", p.multiProcessorCount, ghz, out);
    run<5>("mufu.rcp", p.multiProcessorCount, ghz, out);
    run<6>("mufu.rsqrt", p.multiProcessorCount, ghz, out);
    run<7>("cvt f64<-u64", p.multiProcessorCount, ghz, out);
    run<8>("floor f64", p.multiProcessorCount, ghz, out);
    run<9>("cvt s32<-f64", p.multiProcessorCount, ghz, out);
    run<10>("cvt u64<-f32", p.multiProcessorCount, ghz, out);
    run<11>("cvt u32<-f32", p.multiProcessorCount, ghz, out);
    run<12>("ffma", p.multiProcessorCount, ghz, out);
    run<13>("dfma", p.multiProcessorCount, ghz, out);
    run<14>("cvt f32<-u32", p.multiProcessorCount, ghz, out);
    run<15>("div.rn.f64", p.multiProcessorCount, ghz, out);
    run<16>("cvt f64<-f32", p.multiProcessorCount, ghz, out);
    run<17>("dadd", p.multiProcessorCount, ghz, out);
    run<18>("rint f32", p.multiProcessorCount, ghz, out);
    run<19>("cvt f32<-f64", p.multiProcessorCount, ghz, out);
    run<20>("cvt u64<-f64", p.multiProcessorCount, ghz, out);
    run<21>("cvt s32<-f32", p.multiProcessorCount, ghz, out);
    return 0;
}
