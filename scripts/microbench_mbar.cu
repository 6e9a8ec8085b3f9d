// microbench_mbar.cu -- cost of a full/empty mbarrier hand-shake between a loader warp and a compute
// warp of one CTA, as the chunk sweeps use it (not part of the product).
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
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// mode 0: try_wait spin ; 1: test_wait spin ; 2: test_wait + nanosleep(32) in the LOADER only
template <int NS>
__global__ void handshake(int steps, int loader_warp, int mode, int work, long long* out, double* sink)
{
    __shared__ unsigned long long full[NS], empty[NS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    __syncthreads();
    auto wait = [&](unsigned long long* b, unsigned par, bool loader) {
        if (mode == 0) { while (!mbar_try_wait(b, par)) {} }
        else if (mode == 1 || !loader) { while (!mbar_test_wait(b, par)) {} }
        else { while (!mbar_test_wait(b, par)) __nanosleep(32); }
    };
    if (warp == 0) { // compute warp
        double x = lane;
        const long long t0 = clock64();
        int st = 0; unsigned ph = 0;
        for (int t = 0; t < steps; ++t) {
            wait(full + st, ph, false);
            for (int k = 0; k < work; ++k) x = fma(x, 0.999, 1.0); // dependent DFMA chain = the step's work
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
            if (++st == NS) { st = 0; ph ^= 1; }
        }
        const long long t1 = clock64();
        if (lane == 0) out[0] = t1 - t0;
        sink[lane] = x;
    } else if (warp == loader_warp && lane == 0) {
        int st = 0; unsigned round = 0;
        for (int t = 0; t < steps; ++t) {
            if (round > 0) wait(empty + st, (round - 1) & 1, true);
            mbar_arrive(full + st);
            if (++st == NS) { st = 0; ++round; }
        }
    }
}
int main()
{
    long long* out; double* sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 256);
    const int steps = 4096;
    for (int work : {0, 12, 40})
        for (int mode : {0, 1, 2})
            for (int lw : {4, 1}) {
                handshake<4><<<1, 256>>>(steps, lw, mode, work, out, sink);
                cudaDeviceSynchronize();
                printf("work %2d DFMA, wait mode %d (%s), loader on %s scheduler: %.1f cycles per step\n", work, mode,
                       mode == 0 ? "try_wait spin" : mode == 1 ? "test_wait spin" : "loader test_wait+nanosleep",
                       lw == 4 ? "the SAME" : "another", (double)out[0] / steps);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
