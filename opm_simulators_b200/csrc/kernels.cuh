// kernels.cuh -- hand-written sm_100a kernels of the hot path.
//
//   relayout_kernel      caller's BCSR values -> level-ordered SELL-32 (+ ghost rows -> identity)
//   spmv_kernel          y = A x | y += alpha A x, fused (u,y) and (y,y) partial dots
//   dilu_factor_kernel   Dinv_i = (A_ii - sum A_ij Dinv_j A_ji)^-1        (DILU.hpp:186-206)
//   ilu0_factor_kernel   block ILU(0), left-looking, stored inverse        (ParallelOverlappingILU0_impl.hpp:42-99)
//   sweep_kernel         lower / upper triangular sweeps of DILU and ILU0  (DILU.hpp:253-304, ..ILU0_impl.hpp:383-411)
//   vec_* kernels        fused BiCGSTAB vector updates + dots, fixed-order reductions
//
// All matrix kernels use one lane per block row and one warp per 32-row slice; every matrix
// load is a warp-wide contiguous 256-byte line (layout.hpp).  The triangular kernels are
// "sync-free": slices are started strictly in schedule order (atomic ticket), a row spins on
// the values of the rows it depends on (a NaN-pattern sentinel marks "not yet written"), so no
// grid-wide barrier and no kernel launch per level is paid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opmb200 {

constexpr int kWarpsPerCta = 8;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kPrefetch = 3; // block slots a lane holds in registers / stages before a dependency wait
constexpr unsigned long long kSentinelBits = 0xFFF8B200DEADC0DEull; // quiet NaN never produced by arithmetic

struct SliceMeta { // 32 bytes
    int q0, count, base, wl, wu, level, lrank, pad;
};

// device-resident scalar state of one BiCGSTAB solve (Dune::BiCGSTABSolver::apply locals)
struct Scalars {
    double rho, rho_new, alpha, omega, beta, h, norm, norm0;
    double reduction; // requested |r|/|r0|
    double it;        // Dune's half-step counter
    int maxiter;
    int converged;
    int abort_code; // 0 ok, 1 breakdown (rho/omega/h), 2 NaN/Inf defect
    int done;       // converged || abort || out of iterations  => all later kernels return at once
    int hist_count;
    int hist_cap;
    int factor_error; // singular diagonal block seen by a factorisation kernel
    int pad;
};

enum Epilogue { EPI_NONE = 0, EPI_INIT = 1, EPI_H = 2, EPI_NORM1 = 3, EPI_OMEGA = 4, EPI_NORM2 = 5, EPI_DOT = 6 };

// -------------------------------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_relaxed(const double* p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// weak L1-bypassing load.  NOT usable for the dependency polls: a weak 64-bit load that races with
// the producer's store may return a torn value (observed: a NaN that is neither the sentinel nor
// the final value when producer and consumer share an SM), so every poll is a strong
// ld.relaxed.gpu.
__device__ __forceinline__ double ld_cg(const double* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int* p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool is_sentinel(double v)
{
    return (unsigned long long)__double_as_longlong(v) == kSentinelBits;
}
__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long)kSentinelBits); }
// a genuine result that happens to carry the sentinel bit pattern is stored as the canonical NaN
__device__ __forceinline__ double guard(double v)
{
    return is_sentinel(v) ? __longlong_as_double(0x7FF8000000000000ll) : v;
}

// Level-ordered vectors are stored component-major ("SoA"): component c of row q at c*n + q, so
// that a warp touching 32 consecutive rows moves whole 256-byte lines (loads, stores and the
// dependency polls of the sweeps alike).
#define VIDX(n, q, c) ((size_t)(c) * (size_t)(n) + (size_t)(q))

// Dependency records.  The values the sweeps hand from row to row live in arrays of one aligned
// record per row (2 doubles for b <= 2, 4 doubles = one 32-byte sector for b = 3, 4), so that a
// consumer samples a dependency with ONE strong vector load and a producer publishes it with ONE
// strong vector store: the hardware keeps only a few strong loads in flight per thread
// (scripts/microbench_hop.cu: nine polled words cost 3000 cycles per hop, three cost 1100).
// Every word still validates itself against the sentinel, so no atomicity beyond 8 bytes is needed.
template <int B>
struct Rec {
    static constexpr int W = (B <= 2) ? 2 : 4;
};
template <int B>
__device__ __forceinline__ void rec_load_strong(const double* base, size_t q, double (&x)[B])
{
    const double* p = base + q * Rec<B>::W;
    if constexpr (Rec<B>::W == 2) {
        double w0, w1;
        asm volatile("ld.relaxed.gpu.global.v2.f64 {%0,%1}, [%2];" : "=d"(w0), "=d"(w1) : "l"(p) : "memory");
        x[0] = w0;
        if constexpr (B == 2)
            x[1] = w1;
    } else {
        double w0, w1, w2, w3;
        asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(w0), "=d"(w1), "=d"(w2), "=d"(w3)
                     : "l"(p)
                     : "memory");
        x[0] = w0;
        x[1] = w1;
        x[2] = w2;
        if constexpr (B == 4)
            x[3] = w3;
    }
}
// weak load of a record written by an EARLIER kernel (the lower sweep's y read by the upper sweep).
// NOT __ldcg: ld.global.cg compiles to LDG.STRONG.GPU, and the hardware serialises strong loads
// (~330 cycles each beyond the first few in flight, scripts/microbench_hop.cu)
template <int B>
__device__ __forceinline__ void rec_load_weak(const double* base, size_t q, double (&x)[B])
{
    const double2* p = reinterpret_cast<const double2*>(base + q * Rec<B>::W);
    const double2 a = __ldcs(p);
    x[0] = a.x;
    if constexpr (B >= 2)
        x[1] = a.y;
    if constexpr (B >= 3) {
        const double2 b = __ldcs(p + 1);
        x[2] = b.x;
        if constexpr (B == 4)
            x[3] = b.y;
    }
}

template <int B>
__device__ __forceinline__ void rec_store_strong(double* base, size_t q, const double (&x)[B])
{
    double* p = base + q * Rec<B>::W;
    if constexpr (Rec<B>::W == 2) {
        const double w1 = (B == 2) ? x[B - 1] : 0.0;
        asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x[0]), "d"(w1) : "memory");
    } else {
        const double w3 = (B == 4) ? x[B - 1] : 0.0;
        asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(w3)
                     : "memory");
    }
}
// plain store of a record nobody polls while this kernel runs (a later kernel reads it)
template <int B>
__device__ __forceinline__ void rec_store_weak(double* base, size_t q, const double (&x)[B])
{
    double2* p = reinterpret_cast<double2*>(base + q * Rec<B>::W);
    p[0] = make_double2(x[0], B >= 2 ? x[B >= 2 ? 1 : 0] : 0.0);
    if constexpr (Rec<B>::W == 4)
        p[1] = make_double2(x[B >= 3 ? 2 : 0], B == 4 ? x[B - 1] : 0.0);
}
template <int B>
__device__ __forceinline__ void rec_store_sentinel(double* base, size_t q)
{
    double x[B];
#pragma unroll
    for (int r = 0; r < B; ++r)
        x[r] = __longlong_as_double((long long)kSentinelBits);
    rec_store_strong<B>(base, q, x);
}
template <int B>
__device__ __forceinline__ bool rec_valid(const double (&x)[B])
{
    bool ok = true;
#pragma unroll
    for (int r = 0; r < B; ++r)
        ok = ok && !is_sentinel(x[r]);
    return ok;
}

// element e of block slot (slot_row, lane)
template <int BB>
__device__ __forceinline__ size_t elem_index(int slot_row, int lane, int e)
{
    return (size_t)slot_row * (32 * BB) + (size_t)e * 32 + lane;
}
template <int BB>
__device__ __forceinline__ size_t elem_index_slot(int g, int e)
{
    return (size_t)(g & ~31) * BB + (size_t)e * 32 + (g & 31);
}

// y -= A x ; y += A x ; y = A x     (row-major b x b, Dune mmv / umv / mv)
template <int B>
__device__ __forceinline__ void blk_mmv(const double* A, const double* x, double* y)
{
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
        for (int c = 0; c < B; ++c)
            y[r] -= A[r * B + c] * x[c];
}
template <int B>
__device__ __forceinline__ void blk_umv(const double* A, const double* x, double* y)
{
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
        for (int c = 0; c < B; ++c)
            y[r] += A[r * B + c] * x[c];
}
template <int B>
__device__ __forceinline__ void blk_mv(const double* A, const double* x, double* y)
{
#pragma unroll
    for (int r = 0; r < B; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < B; ++c)
            s += A[r * B + c] * x[c];
        y[r] = s;
    }
}
template <int B>
__device__ __forceinline__ void blk_mm(const double* A, const double* Bm, double* C)
{
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < B; ++k)
                s += A[i * B + k] * Bm[k * B + j];
            C[i * B + j] = s;
        }
}

// Opm::MatrixBlock::invert (matrixblock.hh:255-283).  Returns false for an exactly singular 4x4
// block (the only case in which the reference throws Dune::MatrixBlockError).
template <int B>
__device__ __forceinline__ bool blk_invert(double* a);

template <>
__device__ __forceinline__ bool blk_invert<1>(double* a)
{
    a[0] = 1.0 / a[0];
    return true;
}
template <>
__device__ __forceinline__ bool blk_invert<2>(double* a)
{
    const double det_1 = 1.0 / (a[0] * a[3] - a[1] * a[2]);
    const double a00 = a[0];
    a[0] = a[3] * det_1;
    a[1] = -a[1] * det_1;
    a[2] = -a[2] * det_1;
    a[3] = a00 * det_1;
    return true;
}
template <>
__device__ __forceinline__ bool blk_invert<3>(double* a)
{
    // Dune::FMatrixHelp::invertMatrix 3x3 closed form (same expression tree as
    // gpuistl/detail/deviceBlockOperations.hpp:51-72)
    const double m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[3], m11 = a[4], m12 = a[5], m20 = a[6], m21 = a[7],
                 m22 = a[8];
    const double p0011 = m00 * m11, p0012 = m00 * m12, p0110 = m01 * m10, p0210 = m02 * m10, p0120 = m01 * m20,
                 p0220 = m02 * m20;
    const double rdet
        = 1.0 / (p0011 * m22 - p0012 * m21 - p0110 * m22 + p0210 * m21 + p0120 * m12 - p0220 * m11);
    a[0] = (m11 * m22 - m12 * m21) * rdet;
    a[1] = -(m01 * m22 - m02 * m21) * rdet;
    a[2] = (m01 * m12 - m02 * m11) * rdet;
    a[3] = -(m10 * m22 - m12 * m20) * rdet;
    a[4] = (m00 * m22 - p0220) * rdet;
    a[5] = -(p0012 - p0210) * rdet;
    a[6] = (m10 * m21 - m11 * m20) * rdet;
    a[7] = -(m00 * m21 - p0120) * rdet;
    a[8] = (p0011 - p0110) * rdet;
    return true;
}
// Dune DenseMatrix::invert: LU with partial pivoting, singular iff a pivot is exactly zero
__device__ inline bool blk_invert_lu4(double* M)
{
    double A[16];
    int piv[4];
    for (int i = 0; i < 16; ++i)
        A[i] = M[i];
    for (int i = 0; i < 4; ++i) {
        double pivmax = fabs(A[i * 4 + i]);
        int imax = i;
        for (int k = i + 1; k < 4; ++k) {
            const double v = fabs(A[k * 4 + i]);
            if (v > pivmax) {
                pivmax = v;
                imax = k;
            }
        }
        if (imax != i)
            for (int j = 0; j < 4; ++j) {
                const double t = A[i * 4 + j];
                A[i * 4 + j] = A[imax * 4 + j];
                A[imax * 4 + j] = t;
            }
        piv[i] = imax;
        if (!(pivmax != 0.0))
            return false;
        for (int k = i + 1; k < 4; ++k) {
            const double f = A[k * 4 + i] / A[i * 4 + i];
            A[k * 4 + i] = f;
            for (int j = i + 1; j < 4; ++j)
                A[k * 4 + j] -= f * A[i * 4 + j];
        }
    }
    for (int i = 0; i < 16; ++i)
        M[i] = 0.0;
    for (int i = 0; i < 4; ++i)
        M[i * 4 + i] = 1.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < i; ++j)
            for (int k = 0; k < 4; ++k)
                M[i * 4 + k] -= A[i * 4 + j] * M[j * 4 + k];
    for (int i = 3; i >= 0; --i)
        for (int k = 0; k < 4; ++k) {
            for (int j = i + 1; j < 4; ++j)
                M[i * 4 + k] -= A[i * 4 + j] * M[j * 4 + k];
            M[i * 4 + k] /= A[i * 4 + i];
        }
    for (int i = 3; i >= 0; --i)
        if (i != piv[i])
            for (int j = 0; j < 4; ++j) {
                const double t = M[j * 4 + i];
                M[j * 4 + i] = M[j * 4 + piv[i]];
                M[j * 4 + piv[i]] = t;
            }
    return true;
}
template <>
__device__ __forceinline__ bool blk_invert<4>(double* a)
{
    // adjugate / determinant with the term order of matrixblock.hh:72-190; pivoted-LU fallback
    // for |det| < 1e-40 (:205-224)
    double m[16], inv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        m[i] = a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // rows != j, cols != i
            const int r0 = (j == 0) ? 1 : 0, r1 = (j <= 1) ? 2 : 1, r2 = (j <= 2) ? 3 : 2;
            const int c0 = (i == 0) ? 1 : 0, c1 = (i <= 1) ? 2 : 1, c2 = (i <= 2) ? 3 : 2;
            const double sg = ((i + j) & 1) ? -1.0 : 1.0;
            double s = sg * m[r0 * 4 + c0] * m[r1 * 4 + c1] * m[r2 * 4 + c2];
            s -= sg * m[r0 * 4 + c0] * m[r1 * 4 + c2] * m[r2 * 4 + c1];
            s -= sg * m[r1 * 4 + c0] * m[r0 * 4 + c1] * m[r2 * 4 + c2];
            s += sg * m[r1 * 4 + c0] * m[r0 * 4 + c2] * m[r2 * 4 + c1];
            s += sg * m[r2 * 4 + c0] * m[r0 * 4 + c1] * m[r1 * 4 + c2];
            s -= sg * m[r2 * 4 + c0] * m[r0 * 4 + c2] * m[r1 * 4 + c1];
            inv[i * 4 + j] = s;
        }
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (fabs(det) < 1e-40) {
        if (!blk_invert_lu4(a)) {
            for (int k = 0; k < 16; ++k)
                a[k] = __longlong_as_double(0x7FF8000000000000ll);
            return false;
        }
        return true;
    }
    const double rdet = 1.0 / det;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        a[k] = inv[k] * rdet;
    return true;
}

// -------------------------------------------------------------------------------------------------
// fixed-order reductions.  Every CTA reduces its threads' contributions with a shuffle tree and
// a fixed cross-warp order and writes one partial per dot; the LAST CTA to finish (atomic
// counter) sums the partials in index order with the same tree and runs the scalar epilogue.
// The summation order therefore depends on the launch geometry only, never on scheduling.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int ND>
__device__ __forceinline__ void cta_sum(double (&v)[ND], double* smem /* [ND][32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        v[d] = warp_sum(v[d]);
        if (lane == 0)
            smem[d * 32 + warp] = v[d];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double t = lane < nw ? smem[d * 32 + lane] : 0.0;
            v[d] = warp_sum(t);
        }
    }
}

// ---- peer-to-peer collectives over NVLink (one process per GPU, peers' arenas mapped with CUDA IPC) ----
constexpr int kMaxRanks = 16;
constexpr int kMaxNeighbors = 16;
struct P2PDev {
    int rank, size;
    double* mbox[kMaxRanks]; // rank p's mailboxes: records [2 parities][size senders][4 doubles]
    // Sequence number of the reductions and epoch of the halo exchanges live in device memory and are advanced by
    // the kernels themselves (the same on every rank: every rank runs the same kernel sequence), so that no kernel
    // argument changes from launch to launch and the multi-rank iteration can be replayed as a CUDA graph.
    unsigned long long* seq;
    int* halo_epoch;
    int* abort; // set when a peer did not answer within kSpinTimeoutCycles: the solve reports it instead of hanging
};
constexpr long long kSpinTimeoutCycles = 40000000000ll; // ~20 s at 1.965 GHz

struct ReduceCtx {
    double* partials;      // [ND][max grid]
    unsigned int* counter; // self-resetting arrival counter
    int stride;            // max grid
    double* sums;          // multi-rank over NCCL: the local sums land here, ncclAllReduce and
    int defer;             //   epilogue_kernel follow on the stream (defer != 0)
    const P2PDev* p2p;     // multi-rank over peer memory: the all-reduce happens inside this kernel
    // a reduction spread over two launches (the SpMV split into interior rows and rows that read ghost values, the
    // halo copy running beside the first): the first launch only stores its partials (partial_only), the second
    // stores its own behind them (offset) and its last CTA sums all `total` of them, in index order as ever
    int offset;
    int total;
    int partial_only;
};

__device__ __forceinline__ void st_sys(double* p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys(int* p, int v)
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_sys(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_sys(const int* p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All-reduce (sum) of ND <= 2 doubles across the ranks, executed by warp 0 of ONE CTA per rank:
// lane p pushes this rank's partial into rank p's mailbox over NVLink (values, system fence, then
// the sequence number as the flag), then waits for rank p's partial in its own mailbox.  The
// partials are added in rank order on every rank, so all ranks obtain bitwise identical sums.
// Mailboxes are double-buffered on the parity of the sequence number: a rank can be at most one
// reduction ahead of the slowest rank, because it needs everybody's partial to finish one.
template <int ND>
__device__ __forceinline__ void p2p_allreduce(const P2PDev& c, double (&acc)[ND])
{
    const int lane = threadIdx.x & 31;
    unsigned long long seq = 0;
    if (lane == 0)
        seq = ++(*c.seq); // this CTA is the only one of the rank that reduces
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int par = (int)(seq & 1ull);
    double v[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
        v[d] = 0.0;
    if (lane < c.size) {
        double* dst = c.mbox[lane] + (size_t)(par * c.size + c.rank) * 4;
#pragma unroll
        for (int d = 0; d < ND; ++d)
            st_sys(dst + d, acc[d]);
        __threadfence_system();
        st_sys(reinterpret_cast<unsigned long long*>(dst + 3), seq);
        const double* src = c.mbox[c.rank] + (size_t)(par * c.size + lane) * 4;
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_sys(reinterpret_cast<const unsigned long long*>(src + 3)) != seq) {
            if ((++spins & 1023u) == 0 && clock64() - t0 > kSpinTimeoutCycles) { // a dead peer must not hang the others
                st_relaxed(c.abort, 1);
                break;
            }
        }
        __threadfence_system();
#pragma unroll
        for (int d = 0; d < ND; ++d)
            v[d] = ld_sys(src + d);
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double s = 0.0;
        for (int p = 0; p < c.size; ++p)
            s += __shfl_sync(0xffffffffu, v[d], p);
        acc[d] = s;
    }
}

__device__ void run_epilogue(int epi, Scalars* sc, double* hist, const double* s, double* dot_out)
{
    const double EPSILON = 1e-80; // dune-istl BiCGSTABSolver breakdown threshold
    switch (epi) {
    case EPI_INIT: { // norm0 = |r|, rt = r  =>  rho_new = (rt, r) = |r|^2
        const double nrm = sqrt(s[0]);
        sc->norm0 = nrm;
        sc->norm = nrm;
        sc->it = 0.0;
        sc->rho = 1.0;
        sc->alpha = 1.0;
        sc->omega = 1.0;
        sc->beta = 0.0;
        sc->rho_new = s[0];
        sc->hist_count = 0;
        if (sc->hist_cap > 0)
            hist[sc->hist_count++] = nrm;
        if (!isfinite(nrm)) {
            sc->abort_code = 2;
            sc->done = 1;
        } else if (nrm < nrm * sc->reduction || nrm < 1e-30) {
            sc->converged = 1;
            sc->done = 1;
        } else if (!(0.5 < sc->maxiter)) {
            sc->done = 1;
        }
        break;
    }
    case EPI_H: { // h = (rt, v); alpha = rho_new / h
        sc->h = s[0];
        if (fabs(s[0]) < EPSILON) {
            sc->abort_code = 1;
            sc->done = 1;
        }
        sc->alpha = (sc->norm == 0.0) ? 0.0 : sc->rho_new / s[0];
        break;
    }
    case EPI_NORM1:   // first half step done
    case EPI_NORM2: { // second half step done; s[1] = (rt, r)
        const double nrm = sqrt(s[0]);
        sc->norm = nrm;
        sc->it += 0.5;
        if (sc->hist_count < sc->hist_cap)
            hist[sc->hist_count++] = nrm;
        if (!isfinite(nrm)) {
            sc->abort_code = 2;
            sc->done = 1;
            break;
        }
        if (nrm < sc->norm0 * sc->reduction || nrm < 1e-30) {
            sc->converged = 1;
            sc->done = 1;
            break;
        }
        if (epi == EPI_NORM2) {
            sc->rho = sc->rho_new;
            sc->rho_new = s[1];
            if (!(sc->it + 0.5 < sc->maxiter)) { // for (it = 0.5; it < maxit; it += .5)
                sc->done = 1;
                break;
            }
            // breakdown tests of the next loop head
            if (fabs(sc->rho) <= EPSILON || fabs(sc->omega) <= EPSILON) {
                sc->abort_code = 1;
                sc->done = 1;
                break;
            }
            sc->beta = (nrm == 0.0) ? 0.0 : (sc->rho_new / sc->rho) * (sc->alpha / sc->omega);
        }
        break;
    }
    case EPI_OMEGA: { // s[0] = (t, r), s[1] = (t, t)
        sc->omega = (sc->norm == 0.0) ? 0.0 : s[0] / s[1];
        break;
    }
    case EPI_DOT:
        dot_out[0] = s[0];
        break;
    default:
        break;
    }
}

// called by all threads of every CTA with the CTA's thread-local sums
template <int ND>
__device__ __forceinline__ void grid_reduce(double (&v)[ND], const ReduceCtx& rc, int epi, Scalars* sc, double* hist,
                                            double* dot_out)
{
    __shared__ double red_smem[ND * 32];
    __shared__ bool is_last;
    cta_sum<ND>(v, red_smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
            rc.partials[d * rc.stride + rc.offset + blockIdx.x] = v[d];
        is_last = false;
        if (!rc.partial_only) {
            __threadfence();
            const unsigned int t = atomicAdd(rc.counter, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last)
        return;
    __threadfence();
    const unsigned int nparts = rc.total > 0 ? (unsigned int)rc.total : gridDim.x;
    // the other CTAs' partials must come from the L2 (this SM's L1 may hold the previous reduction's);
    // L1-bypassing loads are slow one by one (~300 cycles each beyond a few in flight), so four
    // partials per strong vector load (rc.stride is a multiple of 4), in a fixed order
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        acc[d] = 0.0;
        const double* p = rc.partials + (size_t)d * rc.stride;
        const unsigned int n4 = nparts >> 2;
        for (unsigned int g = threadIdx.x; g < n4; g += blockDim.x) {
            double w0, w1, w2, w3;
            asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(w0), "=d"(w1), "=d"(w2), "=d"(w3)
                         : "l"(p + 4 * (size_t)g));
            acc[d] += (w0 + w1) + (w2 + w3);
        }
        for (unsigned int i = (n4 << 2) + threadIdx.x; i < nparts; i += blockDim.x)
            acc[d] += __ldcg(p + i);
    }
    __syncthreads();
    cta_sum<ND>(acc, red_smem);
    if (rc.p2p && threadIdx.x < 32)
        p2p_allreduce<ND>(*rc.p2p, acc);
    if (threadIdx.x == 0) {
        *rc.counter = 0u;
        if (rc.defer) {
#pragma unroll
            for (int d = 0; d < ND; ++d)
                rc.sums[d] = acc[d];
        } else {
            run_epilogue(epi, sc, hist, acc, dot_out);
        }
    }
}

// multi-rank: runs the scalar epilogue on the all-reduced sums
__global__ void epilogue_kernel(int epi, Scalars* sc, double* hist, const double* sums, double* dot_out, int check_done)
{
    if (check_done && sc->done)
        return;
    run_epilogue(epi, sc, hist, sums, dot_out);
}

// -------------------------------------------------------------------------------------------------
// relayout: caller's BCSR values -> SELL-32 slots (A, and optionally a second copy for ILU0)
// -------------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) relayout_kernel(int64_t nslots, const int* __restrict__ slot_src,
                                                       const double* __restrict__ vals, double* __restrict__ A,
                                                       double* __restrict__ F)
{
    constexpr int BB = B * B;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nslots; g += (int64_t)gridDim.x * blockDim.x) {
        const int src = slot_src[g];
        double blk[BB];
        if (src >= 0) {
            const double* p = vals + (size_t)src * BB;
#pragma unroll
            for (int e = 0; e < BB; ++e)
                blk[e] = __ldg(p + e);
        } else {
#pragma unroll
            for (int e = 0; e < BB; ++e)
                blk[e] = (src == -2 && (e / B) == (e % B)) ? 1.0 : 0.0;
        }
        const size_t o = (size_t)(g & ~(int64_t)31) * BB + (g & 31);
#pragma unroll
        for (int e = 0; e < BB; ++e) {
            A[o + (size_t)e * 32] = blk[e];
            if (F)
                F[o + (size_t)e * 32] = blk[e];
        }
    }
}

// natural <-> level order for vectors
template <int B>
__global__ void permute_in_kernel(int64_t n, const int* __restrict__ r2n, const double* __restrict__ nat,
                                  double* __restrict__ lvl)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / B;
        const int c = (int)(t - q * B);
        lvl[VIDX(n, q, c)] = nat[(size_t)r2n[q] * B + c];
    }
}
template <int B>
__global__ void permute_out_kernel(int64_t n, const int* __restrict__ n2r, const double* __restrict__ lvl,
                                   double* __restrict__ nat)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / B;
        const int c = (int)(t - i * B);
        nat[t] = lvl[VIDX(n, n2r[i], c)];
    }
}

// -------------------------------------------------------------------------------------------------
// SpMV (MatrixAdapter / GhostLastMatrixAdapter::apply, applyscaleadd; WellOperators.hpp:432-456)
// -------------------------------------------------------------------------------------------------
struct SpmvArgs {
    int nslices;
    const int* slice_list; // null: all slices in order; else the nslices slices of this launch (interior / boundary pass)
    const SliceMeta* slices;
    const int* slot_col;
    const double* A;
    const int* r2n;      // only read when n_interior < n
    int64_t n, n_interior;
    const double* x;
    double* y;
    double alpha;        // SCALEADD: y += alpha A x
    const double* u;     // dot 0 partner: (u, y_new)   (NDOT >= 1)
    double* copy_out;    // optional second destination of y_new (rt = r)
    ReduceCtx rc;
    int epi;
    Scalars* sc;
    double* hist;
    double* dot_out;
    int check_done;
    // standard wells kept outside the matrix (WELLS instantiations only; well_z_kernel ran on x before):
    //   well_head[q]: -1, or t with the perforations e in [well_ptr[t], well_ptr[t+1]) of the cell at position q
    //   well_perf[e] = perforation (its C block: well_C[perf][dw][B]), well_of[e] = its well (z2 = Dinv B x: well_z[well][dw])
    const int* well_head;
    const int* well_ptr;
    const int* well_perf;
    const int* well_of;
    const double* well_C;
    const double* well_z;
    int well_dw;
};

// WELLS: y = (A - C^T D^-1 B) x, the operator of WellModelMatrixAdapter / WellModelGhostLastMatrixAdapter
// (WellOperators.hpp:244-262, 325-356; the reference's GPU twin is a separate kernel after the SpMV,
// gpubridge/cuda/cuWellContributions.cu:37-130): the few perforated rows subtract C_p^T z_w in the order the
// reference's wells visit them, inside the SpMV, so the fused dots see the corrected result.
template <int B, bool SCALEADD, int NDOT, bool WELLS = false>
__global__ void __launch_bounds__(kCtaThreads) spmv_kernel(SpmvArgs a)
{
    constexpr int BB = B * B;
    if (a.check_done && a.sc->done)
        return;
    const int lane = threadIdx.x & 31;
    const int S = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    double acc[B];
#pragma unroll
    for (int r = 0; r < B; ++r)
        acc[r] = 0.0;
    double dots[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int d = 0; d < (NDOT > 0 ? NDOT : 1); ++d)
        dots[d] = 0.0;
    if (S < a.nslices) {
        const SliceMeta m = a.slices[a.slice_list ? a.slice_list[S] : S];
        const bool active = lane < m.count;
        const int q = m.q0 + lane;
        const int nsr = m.wl + 1 + m.wu;
#pragma unroll 2
        for (int sr = 0; sr < nsr; ++sr) {
            const int c = active ? __ldg(a.slot_col + (size_t)(m.base + sr) * 32 + lane) : -1;
            if (c >= 0) {
                double blk[BB], xv[B];
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    blk[e] = __ldcs(a.A + elem_index<BB>(m.base + sr, lane, e));
#pragma unroll
                for (int r = 0; r < B; ++r)
                    xv[r] = __ldg(a.x + VIDX(a.n, c, r));
                blk_umv<B>(blk, xv, acc);
            }
        }
        if (WELLS && active) {
            const int t = __ldg(a.well_head + q);
            if (t >= 0) {
                const int e1 = a.well_ptr[t + 1];
                for (int e = a.well_ptr[t]; e < e1; ++e) {
                    const double* Cp = a.well_C + (size_t)a.well_perf[e] * a.well_dw * B;
                    const double* z = a.well_z + (size_t)a.well_of[e] * a.well_dw;
                    for (int r = 0; r < a.well_dw; ++r) { // BCRSMatrix::mmtv: y[c] -= C[r][c] z[r]
                        const double zr = z[r];
#pragma unroll
                        for (int c = 0; c < B; ++c)
                            acc[c] -= Cp[r * B + c] * zr;
                    }
                }
            }
        }
        if (active) {
            const bool ghost = (a.n_interior < a.n) && (a.r2n[q] >= a.n_interior);
            double out[B];
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double v = ghost ? 0.0 : acc[r];
                if (SCALEADD)
                    v = ghost ? 0.0 : a.y[VIDX(a.n, q, r)] + a.alpha * v;
                out[r] = v;
                a.y[VIDX(a.n, q, r)] = v;
                if (a.copy_out)
                    a.copy_out[VIDX(a.n, q, r)] = v;
            }
            if (NDOT >= 1) {
#pragma unroll
                for (int r = 0; r < B; ++r)
                    dots[0] += out[r] * (a.u ? a.u[VIDX(a.n, q, r)] : out[r]);
            }
            if (NDOT >= 2) {
#pragma unroll
                for (int r = 0; r < B; ++r)
                    dots[NDOT >= 2 ? 1 : 0] += out[r] * out[r];
            }
        }
    }
    if (NDOT > 0)
        grid_reduce<(NDOT > 0 ? NDOT : 1)>(dots, a.rc, a.epi, a.sc, a.hist, a.dot_out);
}

// -------------------------------------------------------------------------------------------------
// in-order slice scheduling for the dependency-carrying kernels
// -------------------------------------------------------------------------------------------------
struct Ticket {
    unsigned int* next; // next chunk to hand out
    unsigned int* done; // CTAs finished (the last one rewinds both counters)
};

// returns the chunk index of this CTA: CTAs obtain chunks in the order they START running, so a
// CTA only ever waits for chunks held by CTAs that are already resident (no deadlock, whatever
// order the hardware dispatches blockIdx in).
__device__ __forceinline__ unsigned int take_ticket(const Ticket& t)
{
    __shared__ unsigned int chunk;
    if (threadIdx.x == 0)
        chunk = atomicAdd(t.next, 1u);
    __syncthreads();
    return chunk;
}
__device__ __forceinline__ void return_ticket(const Ticket& t)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(t.done, 1u) == gridDim.x - 1) {
            *t.next = 0u;
            *t.done = 0u;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// DILU factorisation  (MultithreadDILU::update, DILU.hpp:186-206 / 208-251)
//   Dinv_i = (A_ii - sum_{j<i, A_ji stored} (A_ij Dinv_j) A_ji)^-1
// readiness of Dinv_j is published through row_flag[j] == epoch (release/acquire).
// -------------------------------------------------------------------------------------------------
struct FactorArgs {
    int nslices;
    const SliceMeta* slices;
    const int* order;       // ticket k works on slice order[k] (nullptr: k)
    const int* slot_col;
    const double* A;        // DILU: matrix values ; ILU0: unused
    double* F;              // ILU0: factor values, updated in place
    const int* l_transpose; // DILU
    const int* trip_ptr;    // ILU0
    const int* trip_src;
    const int* trip_dst;
    const int* row_static;  // ILU0: the row's U blocks are never modified (by position)
    double* dinv;           // [n][b*b] by position
    double* dinv_rec;       // Dinv as dependency records [n][b][Rec<b>::W] (sentinel-armed before the launch)
    int* row_flag;          // [n]
    int epoch;
    Ticket ticket;
    Scalars* sc;
};

template <int BB>
__device__ __forceinline__ void wait_row(const int* flag, int epoch)
{
    while (ld_relaxed(flag) != epoch)
        __nanosleep(32);
    __threadfence();
}

// Dinv_j travels from row to row as B dependency records (one per block row, see Rec<B>) in
// `dinv_rec`, armed with sentinels by fill_kernel before the launch: a consumer samples a finished
// Dinv_j with B strong vector loads, no flag and no fence on the producer side.
template <int B>
__global__ void __launch_bounds__(kCtaThreads) dilu_factor_kernel(FactorArgs a)
{
    constexpr int BB = B * B;
    const unsigned int chunk = take_ticket(a.ticket);
    const int lane = threadIdx.x & 31;
    const int T = (int)chunk * kWarpsPerCta + (threadIdx.x >> 5);
    if (T < a.nslices) {
        const int S = a.order ? a.order[T] : T;
        const SliceMeta m = a.slices[S];
        if (lane < m.count) {
            const int q = m.q0 + lane;
            double D[BB];
#pragma unroll
            for (int e = 0; e < BB; ++e)
                D[e] = a.A[elem_index<BB>(m.base + m.wl, lane, e)];
            for (int s0 = 0; s0 < m.wl; s0 += kPrefetch) {
                // everything that does not depend on other rows first ...
                int cj[kPrefetch];
                double Aij[kPrefetch][BB], Aji[kPrefetch][BB];
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k) {
                    const int s = s0 + k;
                    cj[k] = -1;
                    if (s < m.wl) {
                        const int c = a.slot_col[(size_t)(m.base + s) * 32 + lane];
                        const int tr = c >= 0 ? a.l_transpose[(size_t)(m.lrank + s) * 32 + lane] : -1;
                        if (tr >= 0) { // A_ji stored: the term exists (DILU.hpp:196-201)
                            cj[k] = c;
#pragma unroll
                            for (int e = 0; e < BB; ++e) {
                                Aij[k][e] = a.A[elem_index<BB>(m.base + s, lane, e)];
                                Aji[k][e] = a.A[elem_index_slot<BB>(tr, e)];
                            }
                        }
                    }
                }
                // ... then wait for the Dinv_j, all outstanding ones per round trip
                double Dj[kPrefetch][BB];
                unsigned pending = 0;
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k)
                    if (cj[k] >= 0)
                        pending |= 1u << k;
                while (pending) {
#pragma unroll
                    for (int k = 0; k < kPrefetch; ++k)
                        if (pending & (1u << k)) {
#pragma unroll
                            for (int r = 0; r < B; ++r) {
                                double row[B];
                                rec_load_strong<B>(a.dinv_rec, (size_t)cj[k] * B + r, row);
#pragma unroll
                                for (int c2 = 0; c2 < B; ++c2)
                                    Dj[k][r * B + c2] = row[c2];
                            }
                        }
                    unsigned got = 0;
#pragma unroll
                    for (int k = 0; k < kPrefetch; ++k)
                        if (pending & (1u << k)) {
                            bool ok = true;
#pragma unroll
                            for (int e = 0; e < BB; ++e)
                                ok = ok && !is_sentinel(Dj[k][e]);
                            if (ok)
                                got |= 1u << k;
                        }
                    pending &= ~got;
                }
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k)
                    if (cj[k] >= 0) { // Dinv_temp -= (A_ij Dinv_j) A_ji, slots in ascending column order
                        double T1[BB], T2[BB];
                        blk_mm<B>(Aij[k], Dj[k], T1);
                        blk_mm<B>(T1, Aji[k], T2);
#pragma unroll
                        for (int e = 0; e < BB; ++e)
                            D[e] -= T2[e];
                    }
            }
            if (!blk_invert<B>(D))
                a.sc->factor_error = 1;
#pragma unroll
            for (int e = 0; e < BB; ++e) {
                D[e] = guard(D[e]);
                a.dinv[(size_t)q * BB + e] = D[e];
            }
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double row[B];
#pragma unroll
                for (int c2 = 0; c2 < B; ++c2)
                    row[c2] = D[r * B + c2];
                rec_store_strong<B>(a.dinv_rec, (size_t)q * B + r, row);
            }
        }
    }
    return_ticket(a.ticket);
}

// -------------------------------------------------------------------------------------------------
// block ILU(0) factorisation, left-looking with stored inverse diagonal
// (detail::ghost_last_bilu0_decomposition, ParallelOverlappingILU0_impl.hpp:42-99 ==
//  Dune::ILU::blockILU0Decomposition).  F holds a copy of A on entry and L \ D^-1 \ U on exit.
// -------------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(kCtaThreads) ilu0_factor_kernel(FactorArgs a)
{
    constexpr int BB = B * B;
    const unsigned int chunk = take_ticket(a.ticket);
    const int lane = threadIdx.x & 31;
    const int T = (int)chunk * kWarpsPerCta + (threadIdx.x >> 5);
    if (T < a.nslices) {
        const int S = a.order ? a.order[T] : T;
        const SliceMeta m = a.slices[S];
        {   // pull this slice's blocks (one contiguous region) into the L2 while the warp waits
            const char* region = reinterpret_cast<const char*>(a.F + (size_t)m.base * 32 * BB);
            const int lines = (m.wl + 1 + m.wu) * 32 * BB * 8 / 128;
            for (int l = lane; l < lines; l += 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(region + (size_t)l * 128));
        }
        if (lane < m.count) {
            const int q = m.q0 + lane;
            // Fast path (7-point-like rows): at most kPrefetch lower blocks, each of which updates only
            // the diagonal, with rows whose own elimination only ever touched THEIR diagonal.  Then
            // U_ji still holds the caller's value and is read -- like the row's own blocks -- before
            // the wait; the only thing that travels from row to row is Dinv_j, as sentinel-armed
            // dependency records sampled with B strong vector loads (the DILU factorisation's scheme:
            // 27 + 27 separate strong 8-byte loads behind a flag made this kernel 3.8 ms on C3).
            bool fast = m.wl <= kPrefetch;
            int cj[kPrefetch], tsrc[kPrefetch];
            const int gdiag = (m.base + m.wl) * 32 + lane;
#pragma unroll
            for (int k = 0; k < kPrefetch; ++k) {
                cj[k] = -1;
                tsrc[k] = -1;
                if (fast && k < m.wl) {
                    cj[k] = a.slot_col[(size_t)(m.base + k) * 32 + lane];
                    if (cj[k] >= 0) {
                        const size_t cl = (size_t)(m.lrank + k) * 32 + lane;
                        const int t0 = a.trip_ptr[cl], t1 = a.trip_ptr[cl + 1];
                        if (t1 - t0 > 1 || (t1 - t0 == 1 && a.trip_dst[t0] != gdiag))
                            fast = false;
                        else if (t1 - t0 == 1) {
                            tsrc[k] = a.trip_src[t0];
                            if (!a.row_static[cj[k]])
                                fast = false;
                        }
                    }
                }
            }
            if (fast) {
                double Aij[kPrefetch][BB], Uji[kPrefetch][BB], D[BB];
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k)
                    if (cj[k] >= 0) {
#pragma unroll
                        for (int e = 0; e < BB; ++e) {
                            Aij[k][e] = a.F[elem_index<BB>(m.base + k, lane, e)];
                            Uji[k][e] = tsrc[k] >= 0 ? a.F[elem_index_slot<BB>(tsrc[k], e)] : 0.0;
                        }
                    }
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    D[e] = a.F[elem_index<BB>(m.base + m.wl, lane, e)];
                double Dj[kPrefetch][BB];
                unsigned pending = 0;
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k)
                    if (cj[k] >= 0)
                        pending |= 1u << k;
                while (pending) {
#pragma unroll
                    for (int k = 0; k < kPrefetch; ++k)
                        if (pending & (1u << k)) {
#pragma unroll
                            for (int r = 0; r < B; ++r) {
                                double row[B];
                                rec_load_strong<B>(a.dinv_rec, (size_t)cj[k] * B + r, row);
#pragma unroll
                                for (int c2 = 0; c2 < B; ++c2)
                                    Dj[k][r * B + c2] = row[c2];
                            }
                        }
                    unsigned got = 0;
#pragma unroll
                    for (int k = 0; k < kPrefetch; ++k)
                        if (pending & (1u << k)) {
                            bool ok = true;
#pragma unroll
                            for (int e = 0; e < BB; ++e)
                                ok = ok && !is_sentinel(Dj[k][e]);
                            if (ok)
                                got |= 1u << k;
                        }
                    pending &= ~got;
                }
#pragma unroll
                for (int k = 0; k < kPrefetch; ++k)
                    if (cj[k] >= 0) {
                        double Lij[BB], P[BB];
                        blk_mm<B>(Aij[k], Dj[k], Lij);
#pragma unroll
                        for (int e = 0; e < BB; ++e)
                            a.F[elem_index<BB>(m.base + k, lane, e)] = Lij[e];
                        if (tsrc[k] >= 0) {
                            blk_mm<B>(Lij, Uji[k], P);
#pragma unroll
                            for (int e = 0; e < BB; ++e)
                                D[e] -= P[e];
                        }
                    }
                if (!blk_invert<B>(D))
                    a.sc->factor_error = 1;
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    D[e] = guard(D[e]);
#pragma unroll
                for (int r = 0; r < B; ++r) { // publish Dinv first: that is what the next rows wait for
                    double row[B];
#pragma unroll
                    for (int c2 = 0; c2 < B; ++c2)
                        row[c2] = D[r * B + c2];
                    rec_store_strong<B>(a.dinv_rec, (size_t)q * B + r, row);
                }
#pragma unroll
                for (int e = 0; e < BB; ++e) {
                    a.F[elem_index<BB>(m.base + m.wl, lane, e)] = D[e];
                    a.dinv[(size_t)q * BB + e] = D[e];
                    }
                __threadfence(); // rows on the general path wait for the flag and read F and dinv
                st_relaxed(a.row_flag + q, a.epoch);
                goto row_done;
            }
            // 1. all rows this row eliminates with must be complete: poll their flags (strong loads),
            //    then ONE acquire fence; everything after it may use ordinary L2 loads
            for (int s = 0; s < m.wl; ++s) {
                const int c = a.slot_col[(size_t)(m.base + s) * 32 + lane];
                if (c >= 0)
                    while (ld_relaxed(a.row_flag + c) != a.epoch) {}
            }
            __threadfence();
            // 2. eliminate left to right; this row's own blocks are touched by this thread only
            for (int s = 0; s < m.wl; ++s) {
                const int c = a.slot_col[(size_t)(m.base + s) * 32 + lane];
                if (c < 0)
                    continue;
                double Aij[BB], Dj[BB], Lij[BB];
#pragma unroll
                for (int e = 0; e < BB; ++e) {
                    Aij[e] = a.F[elem_index<BB>(m.base + s, lane, e)];
                    Dj[e] = __ldcg(a.dinv + (size_t)c * BB + e);
                }
                blk_mm<B>(Aij, Dj, Lij); // L_ij = A_ij A_jj^-1   (rightmultiply, :63)
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    a.F[elem_index<BB>(m.base + s, lane, e)] = Lij[e];
                const size_t cl = (size_t)(m.lrank + s) * 32 + lane;
                for (int t = a.trip_ptr[cl]; t < a.trip_ptr[cl + 1]; ++t) { // A_ik -= L_ij A_jk (:66-86)
                    const int gs = a.trip_src[t], gd = a.trip_dst[t];
                    double Ujk[BB], P[BB];
#pragma unroll
                    for (int e = 0; e < BB; ++e)
                        Ujk[e] = __ldcg(a.F + elem_index_slot<BB>(gs, e));
                    blk_mm<B>(Lij, Ujk, P);
#pragma unroll
                    for (int e = 0; e < BB; ++e)
                        a.F[elem_index_slot<BB>(gd, e)] -= P[e];
                }
            }
            double D[BB];
#pragma unroll
            for (int e = 0; e < BB; ++e)
                D[e] = a.F[elem_index<BB>(m.base + m.wl, lane, e)];
            if (!blk_invert<B>(D))
                a.sc->factor_error = 1;
#pragma unroll
            for (int e = 0; e < BB; ++e) {
                D[e] = guard(D[e]);
                a.F[elem_index<BB>(m.base + m.wl, lane, e)] = D[e];
                a.dinv[(size_t)q * BB + e] = D[e];
            }
            // 3. publish: release fence, then the flag and the Dinv records the fast rows wait for
            __threadfence();
            st_relaxed(a.row_flag + q, a.epoch);
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double row[B];
#pragma unroll
                for (int c2 = 0; c2 < B; ++c2)
                    row[c2] = D[r * B + c2];
                rec_store_strong<B>(a.dinv_rec, (size_t)q * B + r, row);
            }
        row_done:;
        }
    }
    return_ticket(a.ticket);
}

// -------------------------------------------------------------------------------------------------
// triangular sweeps
//   DILU  lower: y_i = Dinv_i (d_i - sum_{j<i} A_ij y_j)          upper: v_i = y_i - Dinv_i sum_{j>i} A_ij v_j
//   ILU0  lower: y_i = d_i - sum_{j<i} L_ij y_j                   upper: v_i = w Dinv_i (y_i - sum_{j>i} U_ij v_j)
// `tmp` carries y between the two kernels.  Protocol: tmp is all-sentinel between applies; the
// lower kernel arms v with sentinels and fills tmp; the upper kernel consumes tmp (re-arming it)
// and fills v.  A consumer simply re-reads the value it needs until it is not the sentinel.
// -------------------------------------------------------------------------------------------------
struct SweepArgs {
    int nslices;
    const SliceMeta* slices;
    const int* slot_col;
    const double* M;     // block values (A for DILU, F for ILU0)
    const double* dinv;  // [n][b*b]
    const double* d;     // right-hand side (lower)
    double* tmp;         // y, dependency records [n][Rec<B>::W]
    double* vpoll;       // result as dependency records (what the upper sweep's consumers poll)
    double* v;           // result, component-major like every solver vector
    const int* level_q0; // [n_levels+1]
    int n_levels;
    int throttle;        // how many levels behind the front fine-grained polling starts
    double relax;        // ILU0 relaxation w (1.0: none)
    const int* r2n;      // ghost detection (ILU0, parallel)
    int64_t n, n_interior;
    int ghost_zero;      // ILU0 ghost rows: 1 = v enters as 0 (the solver's y = 0), 0 = keep v's input
    Ticket ticket;
    Scalars* sc;
    int check_done;
};


template <int B, bool ILU0, bool UPPER>
__global__ void __launch_bounds__(kCtaThreads, B <= 3 ? 2 : 1) sweep_kernel(SweepArgs a)
{
    constexpr int BB = B * B;
    const unsigned int chunk = take_ticket(a.ticket);
    const bool skip = a.check_done && a.sc->done;
    const int lane = threadIdx.x & 31;
    const int Sfwd = (int)chunk * kWarpsPerCta + (threadIdx.x >> 5);
    const int S = UPPER ? a.nslices - 1 - Sfwd : Sfwd;
    if (!skip && Sfwd < a.nslices) {
        const SliceMeta m = a.slices[S];
        const bool active = lane < m.count;
        const int q = m.q0 + lane;
        const int w = UPPER ? m.wu : m.wl;
        const int sr0 = UPPER ? m.base + m.wl + 1 : m.base;
        double* out = UPPER ? a.vpoll : a.tmp; // record array this sweep produces and its consumers poll

        // ---- 1. prefetch everything that does not depend on other rows --------------------------
        int cj[kPrefetch];
        double blk[kPrefetch][BB];
#pragma unroll
        for (int s = 0; s < kPrefetch; ++s)
            cj[s] = (active && s < w) ? __ldg(a.slot_col + (size_t)(sr0 + s) * 32 + lane) : -1;
#pragma unroll
        for (int s = 0; s < kPrefetch; ++s)
            if (cj[s] >= 0) {
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    blk[s][e] = __ldcs(a.M + elem_index<BB>(sr0 + s, lane, e));
            }
        // first slot beyond the register window (e.g. the ghost plane below a rank's slab, which the
        // ghost-last numbering turns into a 4th upper entry of 13 200 rows): its dependency is sampled
        // with the others and its block copied into shared memory now (cp.async, no registers held)
        // -- handled one after the other behind the wait it cost the upper sweep of such a rank +45 %
        __shared__ double xblk[kWarpsPerCta][BB * 32];
        int cjx = -1;
        if (active && w > kPrefetch) {
            cjx = __ldg(a.slot_col + (size_t)(sr0 + kPrefetch) * 32 + lane);
            if (cjx >= 0) {
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(
                                     &xblk[threadIdx.x >> 5][e * 32 + lane])),
                                 "l"(a.M + elem_index<BB>(sr0 + kPrefetch, lane, e))
                                 : "memory");
            }
        }
        double di[BB], rhs[B], yi[B];
        bool ghost = false;
        if (active) {
            if (!(ILU0 && !UPPER)) {
#pragma unroll
                for (int e = 0; e < BB; ++e)
                    di[e] = __ldcs(a.dinv + (size_t)q * BB + e);
            }
            if (ILU0 && a.n_interior < a.n)
                ghost = a.r2n[q] >= a.n_interior;
            if (!UPPER) {
                // ParallelOverlappingILU0 never touches ghost rows: their v keeps its input value
#pragma unroll
                for (int r = 0; r < B; ++r)
                    rhs[r] = ghost ? (a.ghost_zero ? 0.0 : a.v[VIDX(a.n, q, r)]) : a.d[VIDX(a.n, q, r)];
                rec_store_sentinel<B>(a.vpoll, (size_t)q); // arm the upper sweep's records
            } else {
                rec_load_weak<B>(a.tmp, (size_t)q, yi); // y_i of the lower sweep (complete: previous kernel)
#pragma unroll
                for (int r = 0; r < B; ++r)
                    rhs[r] = (ILU0) ? yi[r] : 0.0;
            }
        }

        // ---- 2. stay asleep while the front is far away (keeps polling traffic off the L2) -------
        if (a.throttle > 0) {
            const int lv = UPPER ? m.level + a.throttle : m.level - a.throttle;
            if (lv >= 0 && lv < a.n_levels) {
                const int pq = UPPER ? a.level_q0[lv] : a.level_q0[lv + 1] - 1;
                const double* pp = out + (size_t)pq * Rec<B>::W + (B - 1);
                if (lane == 0)
                    while (is_sentinel(ld_relaxed(pp)))
                        __nanosleep(200);
                __syncwarp();
            }
        }

        // ---- 3. wait for the rows this row depends on, then accumulate in slot order ------------
        if (active) {
            double xv[kPrefetch][B], xvx[B];
            unsigned pending = 0;
#pragma unroll
            for (int s = 0; s < kPrefetch; ++s)
                if (cj[s] >= 0)
                    pending |= 1u << s;
            if (cjx >= 0)
                pending |= 1u << kPrefetch;
            auto sample = [&](double (&x)[kPrefetch][B], double (&xx)[B]) {
                // all outstanding dependencies are sampled together, one strong vector load each
#pragma unroll
                for (int s = 0; s < kPrefetch; ++s)
                    if (pending & (1u << s))
                        rec_load_strong<B>(out, (size_t)cj[s], x[s]);
                if (pending & (1u << kPrefetch))
                    rec_load_strong<B>(out, (size_t)cjx, xx);
            };
            auto arrived = [&](double (&x)[kPrefetch][B], double (&xx)[B]) -> unsigned {
                // every word validates itself against the sentinel
                unsigned got = 0;
#pragma unroll
                for (int s = 0; s < kPrefetch; ++s)
                    if ((pending & (1u << s)) && rec_valid<B>(x[s]))
                        got |= 1u << s;
                if ((pending & (1u << kPrefetch)) && rec_valid<B>(xx))
                    got |= 1u << kPrefetch;
                return got;
            };
            // (two sample sets in flight half a round trip apart were measured: 540 us instead of 361 --
            // strong loads queue behind each other, fewer in flight is faster)
            while (pending) {
                sample(xv, xvx);
                pending &= ~arrived(xv, xvx);
            }
#pragma unroll
            for (int s = 0; s < kPrefetch; ++s)
                if (cj[s] >= 0) {
                    if (UPPER && !ILU0)
                        blk_umv<B>(blk[s], xv[s], rhs);
                    else
                        blk_mmv<B>(blk[s], xv[s], rhs);
                }
            // rows wider than the register window (NNC / well rows): stream the rest
            for (int s = kPrefetch; s < w; ++s) {
                const int c = s == kPrefetch ? cjx : __ldg(a.slot_col + (size_t)(sr0 + s) * 32 + lane);
                if (c < 0)
                    continue;
                double bl[BB], xs[B];
                if (s == kPrefetch) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < BB; ++e)
                        bl[e] = xblk[threadIdx.x >> 5][e * 32 + lane]; // this lane's own copies
#pragma unroll
                    for (int r = 0; r < B; ++r)
                        xs[r] = xvx[r]; // sampled with the first kPrefetch dependencies
                } else {
#pragma unroll
                    for (int e = 0; e < BB; ++e)
                        bl[e] = __ldcs(a.M + elem_index<BB>(sr0 + s, lane, e));
                    do {
                        rec_load_strong<B>(out, (size_t)c, xs);
                    } while (!rec_valid<B>(xs));
                }
                if (UPPER && !ILU0)
                    blk_umv<B>(bl, xs, rhs);
                else
                    blk_mmv<B>(bl, xs, rhs);
            }

            // ---- 4. finish the row and publish it ------------------------------------------------
            double res[B];
            if (!UPPER) {
                if (ILU0) {
#pragma unroll
                    for (int r = 0; r < B; ++r)
                        res[r] = rhs[r]; // L_ii = I
                } else {
                    blk_mv<B>(di, rhs, res); // y_i = Dinv_i rhs
                }
            } else {
                if (ILU0) {
                    if (ghost) {
#pragma unroll
                        for (int r = 0; r < B; ++r)
                            res[r] = rhs[r];
                    } else {
                        blk_mv<B>(di, rhs, res); // v_i = Dinv_i (y_i - sum); relaxation is applied afterwards
                    }
                } else {
                    blk_mmv<B>(di, rhs, yi); // v_i = y_i - Dinv_i rhs
#pragma unroll
                    for (int r = 0; r < B; ++r)
                        res[r] = yi[r];
                }
                rec_store_sentinel<B>(a.tmp, (size_t)q); // re-arm for the next apply
            }
#pragma unroll
            for (int r = 0; r < B; ++r)
                res[r] = guard(res[r]);
            rec_store_strong<B>(out, (size_t)q, res);
            if (UPPER) {
#pragma unroll
                for (int r = 0; r < B; ++r)
                    a.v[VIDX(a.n, q, r)] = res[r];
            }
        }
    }
    return_ticket(a.ticket);
}

// -------------------------------------------------------------------------------------------------
// shared-memory / mbarrier / TMA helpers of the tile walkers (tile_kernels.cuh)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// non-blocking test of a phase (try_wait may suspend the thread until a time limit)
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// -------------------------------------------------------------------------------------------------
// fused BiCGSTAB vector kernels (Dune::BiCGSTABSolver::apply; one pass each, dots fused)
// -------------------------------------------------------------------------------------------------
struct VecArgs {
    int64_t len; // n * b
    double* x;
    double* r;
    double* p;
    const double* v;
    const double* t;
    const double* y;
    const double* rt;
    ReduceCtx rc;
    Scalars* sc;
    double* hist;
};

// p = r + beta (p - omega v)          (first iteration: beta = 0, p = v = 0  =>  p = r)
__global__ void __launch_bounds__(256) vec_p_update_kernel(VecArgs a)
{
    if (a.sc->done)
        return;
    const double beta = a.sc->beta, omega = a.sc->omega;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += (int64_t)gridDim.x * blockDim.x) {
        double pn = a.p[i] - omega * a.v[i];
        pn *= beta;
        a.p[i] = pn + a.r[i];
    }
}

// x += alpha y ; r -= alpha v ; |r|^2
__global__ void __launch_bounds__(256) vec_half1_kernel(VecArgs a)
{
    if (a.sc->done)
        return;
    const double alpha = a.sc->alpha;
    double s[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += (int64_t)gridDim.x * blockDim.x) {
        a.x[i] += alpha * a.y[i];
        const double rn = a.r[i] - alpha * a.v[i];
        a.r[i] = rn;
        s[0] += rn * rn;
    }
    grid_reduce<1>(s, a.rc, EPI_NORM1, a.sc, a.hist, nullptr);
}

// x += omega y ; r -= omega t ; |r|^2 ; (rt, r)
__global__ void __launch_bounds__(256) vec_half2_kernel(VecArgs a)
{
    if (a.sc->done)
        return;
    const double omega = a.sc->omega;
    double s[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += (int64_t)gridDim.x * blockDim.x) {
        a.x[i] += omega * a.y[i];
        const double rn = a.r[i] - omega * a.t[i];
        a.r[i] = rn;
        s[0] += rn * rn;
        s[1] += a.rt[i] * rn;
    }
    grid_reduce<2>(s, a.rc, EPI_NORM2, a.sc, a.hist, nullptr);
}

// plain dot of two level-ordered vectors restricted to owner rows (ScalarProduct::dot)
template <int B>
__global__ void __launch_bounds__(256) dot_kernel(int64_t n, int64_t n_interior, const int* __restrict__ r2n,
                                                  const double* __restrict__ x, const double* __restrict__ y,
                                                  ReduceCtx rc, int epi, Scalars* sc, double* hist, double* out)
{
    double s[1] = {0.0};
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t % n; // component-major storage
        if (n_interior == n || r2n[q] < n_interior)
            s[0] += x[t] * y[t];
    }
    grid_reduce<1>(s, rc, epi, sc, hist, out);
}

// v *= w   (ParallelOverlappingILU0_impl.hpp:415-417, after the halo copy)
__global__ void __launch_bounds__(256) scale_kernel(int64_t len, double w, double* __restrict__ v, const Scalars* sc,
                                                    int check_done)
{
    if (check_done && sc->done)
        return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        v[i] *= w;
}

// y += a x
__global__ void __launch_bounds__(256) axpy_kernel(int64_t len, double a, const double* __restrict__ x, double* __restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        y[i] += a * x[i];
}

__global__ void fill_kernel(double* p, int64_t len, double v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// halo: pack owner rows / scatter received ghost rows (positions in level order)
template <int B>
__global__ void gather_rows_kernel(int64_t n, int cnt, const int* __restrict__ rows, const double* __restrict__ v,
                                   double* __restrict__ buf)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt * B; t += gridDim.x * blockDim.x)
        buf[t] = v[VIDX(n, rows[t / B], t % B)];
}
template <int B>
__global__ void scatter_rows_kernel(int64_t n, int cnt, const int* __restrict__ rows, const double* __restrict__ buf,
                                    double* __restrict__ v)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt * B; t += gridDim.x * blockDim.x)
        v[VIDX(n, rows[t / B], t % B)] = buf[t];
}

// ---- halo exchange over peer memory: copyOwnerToAll without NCCL --------------------------------------
struct HaloDev {
    int nn;                          // neighbours
    const int* epoch_ptr;            // exchanges completed so far (device counter, same on every rank)
    int* abort;                      // see P2PDev
    int send_ptr[kMaxNeighbors + 1]; // rows sent to neighbour k: send_rows[send_ptr[k] .. send_ptr[k+1])
    int recv_ptr[kMaxNeighbors + 1];
    double* peer_recv[kMaxNeighbors]; // where neighbour k expects my rows (in ITS arena)
    int* peer_dflag[kMaxNeighbors];   // "your data has landed" flag in neighbour k's arena
    int* peer_ack[kMaxNeighbors];     // "I have consumed your data" flag in neighbour k's arena
    const double* my_recv;            // my receive buffer (neighbours write into it)
    int* my_dflag;                    // [nn] set by the neighbours
    int* my_ack;                      // [nn] set by the neighbours
};

// pack the owner rows a neighbour holds as copies straight into the neighbour's receive buffer
template <int B>
__global__ void __launch_bounds__(256) halo_push_kernel(HaloDev h, int64_t n, const int* __restrict__ send_rows,
                                                        const double* __restrict__ v, unsigned int* counter)
{
    // flow control: the neighbour must have consumed the previous exchange before we overwrite it
    const int epoch = *h.epoch_ptr + 1;
    if ((int)threadIdx.x < h.nn) {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_sys(h.my_ack + threadIdx.x) < epoch - 1)
            if ((++spins & 1023u) == 0 && clock64() - t0 > kSpinTimeoutCycles) {
                st_relaxed(h.abort, 1);
                break;
            }
    }
    __syncthreads();
    const int total = h.send_ptr[h.nn] * B;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int row = t / B, c = t - row * B;
        int k = 0;
        while (row >= h.send_ptr[k + 1])
            ++k;
        h.peer_recv[k][(size_t)(row - h.send_ptr[k]) * B + c] = v[VIDX(n, send_rows[row], c)];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            for (int k = 0; k < h.nn; ++k)
                st_sys(h.peer_dflag[k], epoch);
        }
    }
}

// wait for every neighbour's rows, scatter them into the ghost rows, acknowledge
template <int B>
__global__ void __launch_bounds__(256) halo_pull_kernel(HaloDev h, int64_t n, const int* __restrict__ recv_rows, double* __restrict__ v,
                                                        unsigned int* counter, int* epoch_rw)
{
    const int epoch = *h.epoch_ptr + 1;
    if ((int)threadIdx.x < h.nn) {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_sys(h.my_dflag + threadIdx.x) != epoch)
            if ((++spins & 1023u) == 0 && clock64() - t0 > kSpinTimeoutCycles) {
                st_relaxed(h.abort, 1);
                break;
            }
    }
    __threadfence_system();
    __syncthreads();
    const int total = h.recv_ptr[h.nn] * B;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int row = t / B, c = t - row * B;
        v[VIDX(n, recv_rows[row], c)] = __ldcg(h.my_recv + t); // written by a peer: read at the L2
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;
            for (int k = 0; k < h.nn; ++k)
                st_sys(h.peer_ack[k], epoch);
            *epoch_rw = epoch; // the exchange is complete on this rank (every CTA has read the old value long ago)
        }
    }
}

} // namespace opmb200

