// microbench_mbar2.cu -- latency of the individual synchronisation operations a producer/consumer
// pipeline inside one CTA can be built from (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void single_ops(int iters, long long* out)
{
    __shared__ unsigned long long bar[2];
    __shared__ volatile int flag;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, (1 << 20) - 1); flag = 1; }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(bar); // phase 0 of bar[0] complete
    __syncthreads();
    const int lane = threadIdx.x;
    long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < iters; ++i) { while (!mbar_try_wait(bar, 0)) {} acc += i; }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_arrive(bar + 1); }
    t1 = clock64();
    if (lane == 0) out[1] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { while (flag == 0) {} acc += i; }
    t1 = clock64();
    if (lane == 0) out[2] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { __syncwarp(); if (lane == 0) flag = i + 1; }
    t1 = clock64();
    if (lane == 0) out[3] = t1 - t0 + (acc == 12345);
    // try_wait issued early, consumed after independent work (latency hidden?)
    double x = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const bool ok = mbar_try_wait(bar, 0);
#pragma unroll
        for (int k = 0; k < 12; ++k) x = fma(x, 0.999, 1.0);
        if (!ok) x += 1.0;
    }
    t1 = clock64();
    if (lane == 0) out[4] = t1 - t0 + (x == 12345.0);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) x = fma(x, 0.999, 1.0);
    }
    t1 = clock64();
    if (lane == 0) out[5] = t1 - t0 + (x == 12345.0);
}
// flag-based hand-shake (volatile shared memory, no mbarrier) between a compute warp and a loader lane
template <int NS>
__global__ void handshake_flags(int steps, long long* out)
{
    __shared__ volatile int ready[NS]; // step number + 1 that the stage holds
    __shared__ volatile int progress;  // steps the compute warp has finished
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) ready[i] = 0; progress = 0; }
    __syncthreads();
    if (warp == 0) {
        const long long t0 = clock64();
        int st = 0;
        for (int t = 0; t < steps; ++t) {
            while (ready[st] != t + 1) {}
            __syncwarp();
            if (lane == 0) progress = t + 1;
            if (++st == NS) st = 0;
        }
        const long long t1 = clock64();
        if (lane == 0) out[6] = t1 - t0;
    } else if (warp == 4 && lane == 0) {
        int st = 0;
        for (int t = 0; t < steps; ++t) {
            while (progress < t + 1 - NS) {}
            ready[st] = t + 1;
            if (++st == NS) st = 0;
        }
    }
}
int main()
{
    long long* out;
    cudaMallocManaged(&out, 128);
    const int iters = 4096;
    single_ops<<<1, 32>>>(iters, out); cudaDeviceSynchronize();
    printf("try_wait on a completed phase + branch: %.1f cycles\n", (double)out[0] / iters);
    printf("arrive (lane 0, phase never completes):  %.1f cycles\n", (double)out[1] / iters);
    printf("volatile LDS flag test + branch:         %.1f cycles\n", (double)out[2] / iters);
    printf("syncwarp + volatile STS flag:            %.1f cycles\n", (double)out[3] / iters);
    printf("try_wait issued before 12 dependent DFMA, tested after: %.1f cycles (12 DFMA alone: %.1f)\n",
           (double)out[4] / iters, (double)out[5] / iters);
    handshake_flags<4><<<1, 256>>>(iters, out); cudaDeviceSynchronize();
    printf("flag hand-shake compute <-> loader, empty step: %.1f cycles per step\n", (double)out[6] / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
