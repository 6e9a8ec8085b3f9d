// extra_kernels.cuh -- the operators either side of the smoother:
//
//   well_z_kernel          z_w = D_w^-1 (B_w x) per standard well          (StandardWellEquations.cpp:132-144;
//                                                                           gpubridge/cuda/cuWellContributions.cu:37-112)
//                          (the C^T z_w half runs inside spmv_kernel<.., WELLS>, kernels.cuh)
//   cpr_weights_kernel     quasi-IMPES weights from the diagonal blocks     (getQuasiImpesWeights.hpp:64-111;
//                                                                           gpuistl/detail/cpr_amg_operations.cu:35-76)
//   cpr_coarse_kernel      pressure-matrix entries of every block           (PressureTransferPolicy.hpp calculateCoarseEntries;
//                                                                           cpr_amg_operations.cu:79-123)
//   cpr_restrict_kernel    fine residual -> pressure residual               (cpr_amg_operations.cu:126-151)
//   cpr_prolongate_kernel  pressure correction -> fine vector               (cpr_amg_operations.cu:154-178)
//
// The matrix kernels read the SELL-32 copy of A that the solver already keeps (layout.hpp): a warp-wide load of
// one block element is one contiguous 256-byte line, and the coarse-entry kernel touches only the b of b*b
// elements it needs (one column / one row of every block): a third of the matrix for 3x3 blocks.
#pragma once
#include "kernels.cuh"

namespace opmb200 {

struct WellZArgs {
    int n_wells, dw;
    const int* wptr;     // [n_wells+1] perforation ranges
    const int* wpos;     // [n_perf] POSITION of the perforated cell
    const double* B;     // [n_perf][dw][b]
    const double* Dinv;  // [n_wells][dw][dw]
    const double* x;     // position-ordered, component-major
    int64_t n;
    double* z;           // [n_wells][dw]
    const Scalars* sc;
    int check_done;
};

constexpr int kMaxWellEq = 8;

// one warp per well: lane l takes perforations l, l+32, ... (dw partial sums each), a fixed-order shuffle tree
// adds the lanes, lanes < dw apply D^-1.  Run-to-run deterministic; the grouping of the sum over perforations
// differs from the reference's running sum (parity 1e-10 relative, tests/test_gpu_wells_cpr.py).
template <int B>
__global__ void __launch_bounds__(kCtaThreads) well_z_kernel(WellZArgs a)
{
    if (a.check_done && a.sc->done)
        return;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= a.n_wells)
        return;
    double z1[kMaxWellEq];
#pragma unroll
    for (int r = 0; r < kMaxWellEq; ++r)
        z1[r] = 0.0;
    const int p1 = a.wptr[w + 1];
    for (int p = a.wptr[w] + lane; p < p1; p += 32) {
        const int q = a.wpos[p];
        double xv[B];
#pragma unroll
        for (int c = 0; c < B; ++c)
            xv[c] = a.x[VIDX(a.n, q, c)];
        const double* Bp = a.B + (size_t)p * a.dw * B;
#pragma unroll
        for (int r = 0; r < kMaxWellEq; ++r)
            if (r < a.dw) {
#pragma unroll
                for (int c = 0; c < B; ++c)
                    z1[r] += Bp[r * B + c] * xv[c];
            }
    }
#pragma unroll
    for (int r = 0; r < kMaxWellEq; ++r)
        if (r < a.dw)
            z1[r] = warp_sum(z1[r]);
    // every lane now holds (or lane 0 holds) the sums: broadcast lane 0's
#pragma unroll
    for (int r = 0; r < kMaxWellEq; ++r)
        z1[r] = __shfl_sync(0xffffffffu, z1[r], 0);
    if (lane < a.dw) {
        const double* Dw = a.Dinv + ((size_t)w * a.dw + lane) * a.dw;
        double z2 = 0.0;
#pragma unroll
        for (int c = 0; c < kMaxWellEq; ++c)
            if (c < a.dw)
                z2 += Dw[c] * z1[c];
        a.z[(size_t)w * a.dw + lane] = z2;
    }
}

// ---- CPR -------------------------------------------------------------------------------------------
// solve M w = e_p for one b x b block through the inverse the factorisation kernels use (MatrixBlock::invert
// expression trees, kernels.cuh blk_invert); w = column p of M^-1.
template <int B, bool TRANSPOSE>
__global__ void __launch_bounds__(kCtaThreads) cpr_weights_kernel(int nslices, const SliceMeta* slices, const double* A,
                                                                  const int* r2n, int p_index, double* w_nat, int* bad)
{
    constexpr int BB = B * B;
    const int lane = threadIdx.x & 31;
    const int S = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (S >= nslices)
        return;
    const SliceMeta m = slices[S];
    if (lane >= m.count)
        return;
    double D[BB], M[BB];
#pragma unroll
    for (int e = 0; e < BB; ++e)
        D[e] = A[elem_index<BB>(m.base + m.wl, lane, e)];
    // transpose == false: the TRANSPOSED diagonal block is solved (getQuasiImpesWeights.hpp:100-103)
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j)
            M[i * B + j] = TRANSPOSE ? D[i * B + j] : D[j * B + i];
    if (!blk_invert<B>(M))
        atomicExch(bad, 1);
    double bw[B], mx = 0.0;
#pragma unroll
    for (int i = 0; i < B; ++i) {
        bw[i] = 0.0;
#pragma unroll
        for (int j = 0; j < B; ++j)
            bw[i] += (j == p_index) ? M[i * B + j] : 0.0;
        mx = fmax(mx, fabs(bw[i]));
    }
    const size_t row = (size_t)r2n[m.q0 + lane];
#pragma unroll
    for (int i = 0; i < B; ++i)
        w_nat[row * B + i] = bw[i] / mx;
}

// coarse[k] for block k of the caller's BCSR: transpose == false  sum_j A_k[j][p] w_row[j]
//                                             transpose == true   sum_j A_k[p][j] w_col[j]
template <int B, bool TRANSPOSE>
__global__ void __launch_bounds__(kCtaThreads) cpr_coarse_kernel(int nslices, const SliceMeta* slices, const int* slot_col,
                                                                 const int* slot_src, const double* A, const int* r2n,
                                                                 const double* w_nat, int p_index, double* coarse)
{
    constexpr int BB = B * B;
    const int lane = threadIdx.x & 31;
    const int S = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (S >= nslices)
        return;
    const SliceMeta m = slices[S];
    const bool active = lane < m.count;
    double wr[B];
    if (!TRANSPOSE && active) {
        const size_t row = (size_t)r2n[m.q0 + lane];
#pragma unroll
        for (int j = 0; j < B; ++j)
            wr[j] = w_nat[row * B + j];
    }
    const int nsr = m.wl + 1 + m.wu;
    for (int sr = 0; sr < nsr; ++sr) {
        const size_t g = (size_t)(m.base + sr) * 32 + lane;
        const int src = active ? __ldg(slot_src + g) : -1;
        if (src < 0)
            continue;
        if (TRANSPOSE) {
            const size_t colrow = (size_t)r2n[__ldg(slot_col + g)];
#pragma unroll
            for (int j = 0; j < B; ++j)
                wr[j] = w_nat[colrow * B + j];
        }
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int e = TRANSPOSE ? p_index * B + j : j * B + p_index;
            v += A[elem_index<BB>(m.base + sr, lane, e)] * wr[j];
        }
        coarse[src] = v;
    }
}

template <int B, bool TRANSPOSE>
__global__ void cpr_restrict_kernel(int64_t n, const double* __restrict__ fine, const double* __restrict__ w,
                                    int p_index, double* __restrict__ coarse)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (TRANSPOSE) {
            v = fine[i * B + p_index];
        } else {
#pragma unroll
            for (int k = 0; k < B; ++k)
                v += fine[i * B + k] * w[i * B + k];
        }
        coarse[i] = v;
    }
}

template <int B, bool TRANSPOSE>
__global__ void cpr_prolongate_kernel(int64_t n, const double* __restrict__ coarse, const double* __restrict__ w,
                                      int p_index, double* __restrict__ fine)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (TRANSPOSE) {
#pragma unroll
            for (int k = 0; k < B; ++k)
                fine[i * B + k] = coarse[i] * w[i * B + k];
        } else {
            fine[i * B + p_index] = coarse[i];
        }
    }
}

} // namespace opmb200
