// microbench_lat.cu -- instruction latencies the chunk-sweep step model depends on (not part of the
// product).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_lat microbench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double a, double b, int iters, long long* out, double* sink)
{
    double x = a;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            x = fma(x, b, a);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x;
}
// 3 independent chains (the 3 block rows of a 3x3 block)
__global__ void dfma_chain3(double a, double b, int iters, long long* out, double* sink)
{
    double x = a, y = a + 1, z = a + 2;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x = fma(x, b, a);
            y = fma(y, b, a);
            z = fma(z, b, a);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x + y + z;
}
__global__ void lds_chain(int iters, long long* out, int* sink)
{
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            p = next[p];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = p;
}
// store -> __syncwarp -> load by the neighbouring lane (the ring hand-over of the chunk sweeps)
__global__ void sts_sync_lds(int iters, long long* out, double* sink)
{
    __shared__ double ring[128];
    const int lane = threadIdx.x;
    ring[lane] = lane;
    __syncwarp();
    double v = 1.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        ring[(lane + i) & 127] = v;
        __syncwarp();
        v = ring[(lane + i + 127) & 127] + 1.0;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = v;
}
__global__ void clock_overhead(int iters, long long* out)
{
    long long acc = 0, prev = clock64();
    const long long t0 = prev;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        const long long now = clock64();
        acc += now - prev;
        prev = now;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}
// 36 independent LDS.64 then a 12-deep DFMA chain: the shape of one sweep step
__global__ void step_shape(int iters, long long* out, double* sink)
{
    __shared__ double stage[36 * 32];
    for (int i = threadIdx.x; i < 36 * 32; i += 32) stage[i] = 1.0 / (i + 1);
    __syncwarp();
    const int lane = threadIdx.x;
    double r0 = 1, r1 = 2, r2 = 3;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        double x0 = r0, x1 = r1, x2 = r2;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double* b = stage + s * 9 * 32 + lane;
            double n0 = r0, n1 = r1, n2 = r2;
            n0 = fma(-b[0 * 32], x0, n0); n1 = fma(-b[3 * 32], x0, n1); n2 = fma(-b[6 * 32], x0, n2);
            n0 = fma(-b[1 * 32], x1, n0); n1 = fma(-b[4 * 32], x1, n1); n2 = fma(-b[7 * 32], x1, n2);
            n0 = fma(-b[2 * 32], x2, n0); n1 = fma(-b[5 * 32], x2, n1); n2 = fma(-b[8 * 32], x2, n2);
            r0 = n0; r1 = n1; r2 = n2;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = r0 + r1 + r2;
}

int main()
{
    long long* out; double* sink; int* isink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 1024 * 8); cudaMalloc(&isink, 1024 * 4);
    const int iters = 4096;
    dfma_chain<<<1, 32>>>(1.0, 0.999, iters, out, sink); cudaDeviceSynchronize();
    printf("DFMA dependent chain, 1 warp: %.2f cycles per DFMA\n", (double)out[0] / (16.0 * iters));
    dfma_chain3<<<1, 32>>>(1.0, 0.999, iters, out, sink); cudaDeviceSynchronize();
    printf("3 interleaved DFMA chains, 1 warp: %.2f cycles per round of 3\n", (double)out[0] / (16.0 * iters));
    lds_chain<<<1, 32>>>(iters, out, isink); cudaDeviceSynchronize();
    printf("LDS.32 dependent chain: %.2f cycles per load\n", (double)out[0] / (16.0 * iters));
    sts_sync_lds<<<1, 32>>>(iters, out, sink); cudaDeviceSynchronize();
    printf("STS -> syncwarp -> LDS -> DADD round: %.2f cycles\n", (double)out[0] / iters);
    clock_overhead<<<1, 32>>>(iters, out); cudaDeviceSynchronize();
    printf("clock64 read + accumulate: %.2f cycles per mark\n", (double)out[0] / iters);
    step_shape<<<1, 32>>>(iters, out, sink); cudaDeviceSynchronize();
    printf("step shape (36 LDS.64 + 4 x 3-deep DFMA, 1 warp): %.2f cycles per step\n", (double)out[0] / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
