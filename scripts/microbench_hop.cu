// microbench_hop.cu -- measures the producer->consumer "hop" latencies the sweep design depends on
// (not part of the product).  nvcc -arch=sm_100a -O3 -o microbench_hop microbench_hop.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 1. dependent chain of relaxed gpu-scope loads (L2 round trip)
__global__ void chase(const unsigned long long* buf, int iters, long long* out)
{
    unsigned long long idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
        idx = ld_relaxed(buf + idx);
    long long t1 = clock64();
    out[0] = t1 - t0;
    out[1] = (long long)idx;
}

// 2. ping-pong through global memory between CTA 0 and CTA `peer`
__global__ void pingpong(unsigned long long* flags, int peer, int iters, long long* out)
{
    if (threadIdx.x != 0)
        return;
    unsigned long long* fa = flags;
    unsigned long long* fb = flags + 64; // different lines
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            st_relaxed(fa, (unsigned long long)i);
            while (ld_relaxed(fb) != (unsigned long long)i) {}
        }
        out[0] = clock64() - t0;
    } else if ((int)blockIdx.x == peer) {
        for (int i = 1; i <= iters; ++i) {
            while (ld_relaxed(fa) != (unsigned long long)i) {}
            st_relaxed(fb, (unsigned long long)i);
        }
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        out[1] = smid;
    }
}

// 3. ping-pong through shared memory between two warps of one CTA
__global__ void pingpong_smem(int iters, long long* out)
{
    __shared__ volatile unsigned long long fa, fb;
    if (threadIdx.x == 0) { fa = 0; fb = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            fa = i;
            while (fb != (unsigned long long)i) {}
        }
        out[0] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        for (int i = 1; i <= iters; ++i) {
            while (fa != (unsigned long long)i) {}
            fb = i;
        }
    }
}

// 4. ping-pong through distributed shared memory inside a cluster of 2
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int iters, long long* out)
{
    __shared__ volatile unsigned long long flag;
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) flag = 0;
    cl.sync();
    const unsigned r = cl.block_rank();
    volatile unsigned long long* peer = cl.map_shared_rank((unsigned long long*)&flag, r ^ 1);
    if (threadIdx.x == 0) {
        if (r == 0) {
            long long t0 = clock64();
            for (int i = 1; i <= iters; ++i) {
                *peer = i;                                   // write into CTA 1's smem
                while (flag != (unsigned long long)i) {}     // wait for CTA 1's answer in my smem
            }
            out[0] = clock64() - t0;
        } else {
            for (int i = 1; i <= iters; ++i) {
                while (flag != (unsigned long long)i) {}
                *peer = i;
            }
        }
    }
    cl.sync();
}

// 5. hop with a warp-wide payload: producer warp stores 32x3 doubles, consumer warp polls all of them
__global__ void pingpong_wide(double* bufA, double* bufB, int peer, int iters, long long* out, int nwords)
{
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 && (int)blockIdx.x != peer)
        return;
    double* mine = blockIdx.x == 0 ? bufA : bufB;
    double* other = blockIdx.x == 0 ? bufB : bufA;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (blockIdx.x == 0) {
            for (int w = 0; w < nwords; ++w)
                asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(mine + w * 4096 + lane), "d"((double)i) : "memory");
        }
        bool ok;
        do {
            ok = true;
            double v[9];
            for (int w = 0; w < nwords; ++w)
                asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[w]) : "l"(other + w * 4096 + lane) : "memory");
            for (int w = 0; w < nwords; ++w)
                ok = ok && (v[w] == (double)i);
        } while (!__all_sync(0xffffffffu, ok));
        if (blockIdx.x != 0) {
            for (int w = 0; w < nwords; ++w)
                asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(mine + w * 4096 + lane), "d"((double)i) : "memory");
        }
    }
    if (blockIdx.x == 0 && lane == 0)
        out[0] = clock64() - t0;
}

int main()
{
    const int iters = 2000;
    long long* out;
    cudaMallocManaged(&out, 64);
    unsigned long long* buf;
    cudaMalloc(&buf, 1 << 20);
    cudaMemset(buf, 0, 1 << 20);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (max) %d kHz\n", clk);
    chase<<<1, 1>>>(buf, iters, out);
    cudaDeviceSynchronize();
    printf("ld.relaxed.gpu dependent chain (L2 hit): %.1f cycles per load\n", (double)out[0] / iters);
    for (int peer : {1, 2, 5, 20, 74, 100, 147}) {
        cudaMemset(buf, 0, 4096);
        pingpong<<<148, 32>>>(buf, peer, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("global ping-pong CTA0 <-> CTA%-3d (smid %lld): %.1f cycles per hop  (%s)\n", peer, out[1],
               (double)out[0] / (2.0 * iters), cudaGetErrorString(e));
    }
    pingpong_smem<<<1, 64>>>(iters, out);
    cudaDeviceSynchronize();
    printf("shared-memory ping-pong (2 warps, 1 CTA): %.1f cycles per hop\n", (double)out[0] / (2.0 * iters));
    pingpong_dsmem<<<2, 32>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("DSMEM ping-pong (cluster of 2): %.1f cycles per hop (%s)\n", (double)out[0] / (2.0 * iters), cudaGetErrorString(e));
    double *bA, *bB;
    cudaMalloc(&bA, 9 * 4096 * 8);
    cudaMalloc(&bB, 9 * 4096 * 8);
    for (int nw : {1, 3, 9}) {
        for (int peer : {1, 74}) {
            cudaMemset(bA, 0, 9 * 4096 * 8);
            cudaMemset(bB, 0, 9 * 4096 * 8);
            pingpong_wide<<<148, 32>>>(bA, bB, peer, iters, out, nw);
            e = cudaDeviceSynchronize();
            printf("warp-wide hop, %d x 256B lines, CTA0 <-> CTA%-3d: %.1f cycles per hop (%s)\n", nw, peer,
                   (double)out[0] / (2.0 * iters), cudaGetErrorString(e));
        }
    }
    return 0;
}
