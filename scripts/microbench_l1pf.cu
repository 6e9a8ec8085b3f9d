// microbench_l1pf.cu -- can a helper warp keep a compute warp's record stream in the L1 so that the
// compute warp's step (45 coalesced LDG.64 + a 12-deep DFMA chain + ring hand-over) runs at L1-hit
// speed with NO barrier between the two?  (design study for the chunk sweeps; not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kRec = 9728, kLines = kRec / 128;
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// mode 0: no helper ; 1: prefetch.global.L1 ; 2: real ld.global.nc touches ; bit 4: + bulk L2 prefetch 16 ahead
__global__ void __launch_bounds__(256) run(const unsigned char* stream, int steps, int mode, int dist, int nwarps, long long* out, double* sink)
{
    __shared__ double ring[4][3 * 128];
    __shared__ volatile int progress[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = warp & 3;
    if (threadIdx.x < 4) progress[threadIdx.x] = 0;
    __syncthreads();
    if (wi >= nwarps) return;
    const unsigned char* base = stream + ((size_t)blockIdx.x * 4 + wi) * (size_t)steps * kRec;
    if (warp < 4) {
        double r0 = 1, r1 = 2, r2 = 3;
        const long long t0 = clock64();
        for (int t = 0; t < steps; ++t) {
            const double* rec = reinterpret_cast<const double*>(base + (size_t)t * kRec + 512) + lane;
            double x0 = ring[wi][(lane + t) & 127], x1 = ring[wi][128 + ((lane + t) & 127)], x2 = ring[wi][256 + ((lane + t) & 127)];
            double n0 = r0 * 1e-3, n1 = r1 * 1e-3, n2 = r2 * 1e-3;
            double bv[36];
#pragma unroll
            for (int e = 0; e < 36; ++e)
                bv[e] = __ldg(rec + e * 32);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double* b = bv + s * 9;
                n0 = fma(-b[0], x0, n0); n1 = fma(-b[3], x0, n1); n2 = fma(-b[6], x0, n2);
                n0 = fma(-b[1], x1, n0); n1 = fma(-b[4], x1, n1); n2 = fma(-b[7], x1, n2);
                n0 = fma(-b[2], x2, n0); n1 = fma(-b[5], x2, n1); n2 = fma(-b[8], x2, n2);
            }
            r0 = n0; r1 = n1; r2 = n2;
            ring[wi][(lane + t + 1) & 127] = r0; ring[wi][128 + ((lane + t + 1) & 127)] = r1; ring[wi][256 + ((lane + t + 1) & 127)] = r2;
            __syncwarp();
            if (lane == 0) progress[wi] = t + 1;
        }
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[wi] = t1 - t0;
        sink[(blockIdx.x * 4 + wi) * 32 + lane] = r0 + r1 + r2;
    } else if ((mode & 3) != 0) {
        int pf = 0, pf2 = 0;
        while (pf < steps) {
            const int p = progress[wi];
            if (p >= steps) break;
            if ((mode & 4) && lane == 0)
                for (; pf2 < min(steps, p + 16); ++pf2) l2_prefetch_bulk(base + (size_t)pf2 * kRec, kRec);
            if (pf < p + dist) {
                const unsigned char* rec = base + (size_t)pf * kRec;
                for (int i = lane; i < kLines; i += 32) {
                    if ((mode & 3) == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + i * 128));
                    else { int v; (void)v; asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(rec + i * 128)); }
                }
                ++pf;
            } else {
                __nanosleep(40);
            }
        }
    }
}
int main()
{
    const int steps = 90, nblk = 148;
    const size_t bytes = (size_t)nblk * 4 * steps * kRec;
    unsigned char* stream; long long* out; double* sink;
    cudaMalloc(&stream, bytes + kRec); cudaMemset(stream, 0, bytes + kRec);
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, nblk * 4 * 32 * 8);
    unsigned char* flush; cudaMalloc(&flush, 256 << 20);
    for (int grid : {1, 148})
        for (int nwarps : {1, 4})
            for (int mode : {0, 1, 2, 5})
                for (int dist : {2, 4}) {
                    if (mode == 0 && dist != 2) continue;
                    cudaMemset(flush, 1, 256 << 20); // evict the stream from the L2
                    cudaDeviceSynchronize();
                    run<<<grid, 256>>>(stream, steps, mode, dist, nwarps, out, sink);
                    cudaDeviceSynchronize();
                    printf("grid %3d, %d chunk(s)/SM, helper mode %d, distance %d: %.0f cycles per step (cold L2)", grid, nwarps, mode, dist,
                           (double)out[0] / steps);
                    run<<<grid, 256>>>(stream, steps, mode, dist, nwarps, out, sink);
                    cudaDeviceSynchronize();
                    printf("   %.0f (L2-warm)\n", (double)out[0] / steps);
                }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
