// tile_kernels.cuh -- tile walkers: the triangular sweeps of schedule "tiles" (DESIGN.md section 6).
//
// The level schedule pays one L2 round trip per level (C3: 363 levels x ~1 us).  Here ONE CTA walks
// ONE CHUNK of rows -- on a box grid a tile of TJ x TK grid lines, one line per row of a step, the
// lines skewed so that step s holds cell i = s - lj - lk of line (lj, lk) -- step after step:
//   * B lanes per block row (one lane per row of the b x b blocks): a step of R = 4 warps x 32/b rows
//     is ~60 instructions per warp, its dependent chain 9 DFMA + 1 shuffle round + 3 DFMA;
//   * results travel from step to step through a shared-memory ring (one named barrier per step);
//     only dependencies that cross a chunk boundary travel through the L2, and those are polled by
//     separate POLL warps (sentinel-validated dependency records, as in sweep_kernel) which park the
//     values in the stage -- the compute warps never touch global memory for input;
//   * everything else a step needs is ONE contiguous record of a per-sweep stream (header, external
//     list, dependency codes, the rows' block values lane by lane, Dinv): a LOADER warp brings it in
//     with one TMA bulk copy (cp.async.bulk -> UBLKCP) per step into a ring of stages, plus the
//     right-hand side (cp.async from the solver vector, or a second bulk copy of the lower sweep's
//     records), all signalled on the stage's "full" mbarrier;
//   * CTAs are persistent and take chunks through the in-order ticket, so a chunk only ever waits
//     for chunks that are running or done.
// The arithmetic per row is the level kernels' (same blocks, same order of the fused multiply-adds),
// so both schedules give bit-identical preconditioner applications.
#pragma once
#include "kernels.cuh"
#include "layout.hpp"

namespace opmb200 {

constexpr int kTwPollWarps = 3;
constexpr int kTwThreads = (kTwWarps + 1 + kTwPollWarps) * 32;

template <int B, int S, bool DINV>
struct TwCfg {
    static constexpr int NW = kTwWarps;               // compute warps
    static constexpr int RPW = 32 / B;                // rows per warp
    static constexpr int R = NW * RPW;                // rows per step
    static constexpr int RP = (R + 3) & ~3;
    static constexpr int W = Rec<B>::W;               // doubles per dependency record in global memory
    static constexpr int NV = S * B + (DINV ? B : 0); // doubles per lane and step
    static constexpr int NP = (NV + 1) / 2;           // ... as 16-byte pairs
    static constexpr int RING = (4 * R <= 256) ? 256 : 512; // == Layout::tw_ring
    // record (global memory) == head of a stage (shared memory)
    static constexpr int kHdrOff = 0;                                // int4 {q0, count, n_ext, flags}
    static constexpr int kExtPosOff = 16;                            // kTwMaxExt positions
    static constexpr int kCodeOff = kExtPosOff + kTwMaxExt * 4;      // [S][RP] dependency codes
    static constexpr int kValOff = (kCodeOff + S * RP * 4 + 15) & ~15; // [NW][NP][32] double2
    static constexpr int kRecBytes = kValOff + NW * NP * 512;
    // what the loaders add to a stage
    static constexpr int kRhsOff = kRecBytes;                        // [RP][W] doubles
    static constexpr int kExtValOff = kRhsOff + RP * W * 8;          // [kTwMaxExt][4] doubles
    static constexpr int kStageBytes = (kExtValOff + kTwMaxExt * 32 + 127) & ~127;
    static constexpr int kStagesRaw = 98304 / kStageBytes;
    static constexpr int kStages = kStagesRaw < 3 ? 3 : (kStagesRaw > 8 ? 8 : kStagesRaw);
    static constexpr int kRingOff = kStages * kStageBytes;           // [RING][4] doubles
    static constexpr int kZeroOff = kRingOff + RING * 32;            // one all-zero record
    static constexpr int kBarOff = kZeroOff + 32;                    // kStages "full" mbarriers
    static constexpr int kCtlOff = kBarOff + 8 * kStages;            // int[4]: progress, first record, steps, stop
    static constexpr int kSmemBytes = kCtlOff + 16;
};

struct TwArgs {
    int nchunks;
    const int* chunk_step0;      // [nchunks+1]
    int nsteps;
    const unsigned char* stream; // this sweep's step records in WALKING order (upper: reversed)
    const double* d;             // lower: right-hand side (component-major)
    double* tmp;                 // dependency records of the lower sweep's result y
    double* vpoll;               // dependency records of the upper sweep's result
    double* v;                   // result, component-major like every solver vector
    int64_t n;
    int ghost_zero;              // ILU0 ghost rows: 1 = v enters as 0, 0 = keep v's input
    Ticket ticket;
    Scalars* sc;
    int check_done;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void lds_v2(unsigned addr, double& a, double& b)
{
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double a;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
    return a;
}
__device__ __forceinline__ int lds_s32(unsigned addr)
{
    int a;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a) : "r"(addr));
    return a;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double a)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory");
}
__device__ __forceinline__ void sts_v2(unsigned addr, double a, double b)
{
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival once all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- compute warps -------------------------------------------------------------------------------
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_compute(const TwArgs& a, unsigned char* smem, int g0, int ns, int warp, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV>;
    constexpr int NS = T::kStages;
    const int rw = lane / B, r = lane - rw * B;
    const bool lane_ok = rw < T::RPW;
    const int rho = lane_ok ? warp * T::RPW + rw : 0; // idle lanes shadow row 0 of the step: computed, never stored
    const int src0 = lane_ok ? rw * B : 0;            // first lane of this row
    const unsigned smem_s = smem_u32(smem);
    const unsigned ring_s = smem_s + T::kRingOff, zero_s = smem_s + T::kZeroOff;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    double* out = UPPER ? a.vpoll : a.tmp;
    const double sent = sentinel();

    for (int t = 0; t < ns; ++t) {
        const int g = g0 + t, st = g % NS;
        const unsigned sb = smem_s + (unsigned)st * T::kStageBytes;
        mbar_wait(full + st, (unsigned)(g / NS) & 1u);
        const int q0 = lds_s32(sb + T::kHdrOff), count = lds_s32(sb + T::kHdrOff + 4);
        const bool active = lane_ok && rho < count;
        const int q = q0 + rho;
        // all shared-memory loads of the step issue back to back ahead of the one dependent DFMA chain
        double x[S][B];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int c = lds_s32(sb + T::kCodeOff + (unsigned)(s * T::RP + rho) * 4);
            const unsigned addr = c < 0 ? zero_s
                                        : ((c & kTwRing) ? ring_s + (unsigned)(c & (T::RING - 1)) * 32
                                                         : sb + T::kExtValOff + (unsigned)(c & (kTwMaxExt - 1)) * 32);
            if constexpr (B == 1) {
                x[s][0] = lds_f64(addr);
            } else {
                lds_v2(addr, x[s][0], x[s][1]);
                if constexpr (B == 3)
                    x[s][2] = lds_f64(addr + 16);
                if constexpr (B == 4)
                    lds_v2(addr + 16, x[s][2], x[s][B - 1]);
            }
        }
        double av[2 * T::NP];
#pragma unroll
        for (int k = 0; k < T::NP; ++k)
            lds_v2(sb + T::kValOff + (unsigned)((warp * T::NP + k) * 32 + lane) * 16, av[2 * k], av[2 * k + 1]);
        const double in = lds_f64(sb + T::kRhsOff + (unsigned)(rho * T::W + r) * 8);

        // ---- the row: same blocks, same order of operations as sweep_kernel ----------------------
        double tsum = (UPPER && !ILU0) ? 0.0 : in;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int c = 0; c < B; ++c) {
                if (UPPER && !ILU0)
                    tsum += av[s * B + c] * x[s][c]; // blk_umv
                else
                    tsum -= av[s * B + c] * x[s][c]; // blk_mmv
            }
        double res;
        if (DINV) {
            double tt[B];
#pragma unroll
            for (int c = 0; c < B; ++c)
                tt[c] = __shfl_sync(0xffffffffu, tsum, src0 + c);
            if (UPPER && !ILU0) { // v_i = y_i - Dinv_i sum   (blk_mmv)
                res = in;
#pragma unroll
                for (int c = 0; c < B; ++c)
                    res -= av[S * B + c] * tt[c];
            } else { // Dinv_i (rhs - sum)   (blk_mv)
                res = 0.0;
#pragma unroll
                for (int c = 0; c < B; ++c)
                    res += av[S * B + c] * tt[c];
            }
        } else {
            res = tsum; // ILU0 lower: L_ii = I
        }
        res = guard(res);
        if (active) {
            sts_f64(ring_s + (unsigned)(q & (T::RING - 1)) * 32 + r * 8, res);
            st_relaxed(out + (size_t)q * T::W + r, res);
            if (UPPER) {
                a.v[VIDX(a.n, q, r)] = res;
                st_relaxed(a.tmp + (size_t)q * T::W + r, sent); // re-arm for the next apply
            } else {
                st_relaxed(a.vpoll + (size_t)q * T::W + r, sent); // arm the upper sweep's records
            }
        }
        named_bar_sync(1, T::NW * 32); // ring writes visible to the four warps; everybody is done with the stage
        if (threadIdx.x == 0)
            ctl[0] = g + 1; // releases the stage to the loaders
    }
}

// ---- loader warp: one TMA bulk copy per step + the right-hand side -------------------------------------
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_loader(const TwArgs& a, unsigned char* smem, int rec0, int g0, int ns, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV>;
    constexpr int NS = T::kStages;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const unsigned char* rec = a.stream + (size_t)rec0 * T::kRecBytes;
    for (int t = 0; t < ns; ++t, rec += T::kRecBytes) {
        const int g = g0 + t, st = g % NS;
        unsigned char* sb = smem + (size_t)st * T::kStageBytes;
        const int4 hdr = __ldg(reinterpret_cast<const int4*>(rec)); // q0, count, n_ext, flags
        while (ctl[0] < g + 1 - NS)
            __nanosleep(20);
        if (lane == 0) {
            mbar_expect_tx(full + st, (unsigned)T::kRecBytes + (UPPER ? (unsigned)hdr.y * T::W * 8 : 0u));
            tma_load_1d(sb, rec, (unsigned)T::kRecBytes, full + st);
            if (UPPER) // y_i of the lower sweep (complete: previous kernel)
                tma_load_1d(sb + T::kRhsOff, a.tmp + (size_t)hdr.x * T::W, (unsigned)hdr.y * T::W * 8, full + st);
        }
        if (!UPPER) {
            const bool ghost = ILU0 && (hdr.w & 1); // ParallelOverlappingILU0 never touches ghost rows
            const unsigned rhs_s = smem_u32(sb + T::kRhsOff);
            if (ghost && a.ghost_zero) {
                for (int e = lane; e < hdr.y * T::W; e += 32)
                    sts_f64(rhs_s + e * 8, 0.0);
                mbar_arrive(full + st);
            } else {
                const double* src = ghost ? a.v : a.d;
#pragma unroll
                for (int c = 0; c < B; ++c)
                    for (int rho = lane; rho < hdr.y; rho += 32)
                        cp_async8(rhs_s + (unsigned)(rho * T::W + c) * 8, src + VIDX(a.n, hdr.x + rho, c));
                cp_async_arrive_noinc(full + st);
            }
        }
    }
}

// ---- poll warps: the dependencies the chunk's ring does not serve ---------------------------------------
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_poller(const TwArgs& a, unsigned char* smem, int rec0, int g0, int ns, int pw, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV>;
    constexpr int NS = T::kStages;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const double* out = UPPER ? a.vpoll : a.tmp;
    for (int t = pw; t < ns; t += kTwPollWarps) {
        const unsigned char* rec = a.stream + (size_t)(rec0 + t) * T::kRecBytes;
        const int g = g0 + t, st = g % NS;
        const int n_ext = __ldg(reinterpret_cast<const int*>(rec) + 2);
        const int pos = lane < n_ext ? __ldg(reinterpret_cast<const int*>(rec + T::kExtPosOff) + lane) : -1;
        double x[B];
        if (pos >= 0) {
            int tries = 0;
            do { // one strong vector load per round trip; every word validates itself against the sentinel
                if (++tries > 8)
                    __nanosleep(40);
                rec_load_strong<B>(out, (size_t)pos, x);
            } while (!rec_valid<B>(x));
        }
        while (ctl[0] < g + 1 - NS)
            __nanosleep(20);
        if (pos >= 0) {
            const unsigned dst = smem_u32(smem + (size_t)st * T::kStageBytes + T::kExtValOff) + lane * 32;
#pragma unroll
            for (int c = 0; c < B; ++c)
                sts_f64(dst + c * 8, x[c]);
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(full + st); // release: the stores above are visible to whoever passes the barrier
    }
}

template <int B, int S, bool ILU0, bool UPPER>
__global__ void __launch_bounds__(kTwThreads, 2) tw_sweep_kernel(TwArgs a)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    if (threadIdx.x == 0) {
        // arrivals per stage: the loader's expect_tx, the poll warp, and (lower) the loader lanes' cp.async
        for (int i = 0; i < T::kStages; ++i)
            mbar_init(full + i, UPPER ? 2u : 34u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 4; ++i) {
            reinterpret_cast<double*>(smem + T::kZeroOff)[i] = 0.0;
            ctl[i] = 0;
        }
    }
    // idle lanes and rows beyond a step's count read ring / rhs words nobody wrote: keep them finite
    for (int i = threadIdx.x; i < (T::kZeroOff - T::kRingOff) / 8; i += blockDim.x)
        reinterpret_cast<double*>(smem + T::kRingOff)[i] = 0.0;
    for (int st = 0; st < T::kStages; ++st)
        for (int i = threadIdx.x; i < (T::kStageBytes - T::kRhsOff) / 8; i += blockDim.x)
            reinterpret_cast<double*>(smem + (size_t)st * T::kStageBytes + T::kRhsOff)[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes before the TMA writes
    __syncthreads();
    const bool skip = a.check_done && a.sc->done;
    int g = 0; // steps walked by this CTA so far: stage = g % kStages, phase = (g / kStages) & 1
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned int tk = atomicAdd(a.ticket.next, 1u);
            if (skip || tk >= (unsigned)a.nchunks) {
                ctl[3] = 1;
            } else {
                const int c = UPPER ? a.nchunks - 1 - (int)tk : (int)tk;
                const int s0 = a.chunk_step0[c], s1 = a.chunk_step0[c + 1];
                ctl[1] = UPPER ? a.nsteps - s1 : s0; // first record of the chunk in walking order
                ctl[2] = s1 - s0;
            }
        }
        __syncthreads();
        if (ctl[3])
            break;
        const int rec0 = ctl[1], ns = ctl[2];
        if (warp < T::NW)
            tw_compute<B, S, ILU0, UPPER>(a, smem, g, ns, warp, lane);
        else if (warp == T::NW)
            tw_loader<B, S, ILU0, UPPER>(a, smem, rec0, g, ns, lane);
        else
            tw_poller<B, S, ILU0, UPPER>(a, smem, rec0, g, ns, warp - T::NW - 1, lane);
        g += ns;
        __syncthreads(); // everybody has read ctl[1..2]; the chunk is finished
    }
    return_ticket(a.ticket);
}

// fills the value part of the step records after a factorisation: block values from the SELL slots (A
// for DILU, F for ILU0) lane by lane, Dinv from its row-major array.  One thread per (step, warp, lane).
template <int B, int S, bool DINV>
__global__ void __launch_bounds__(256) tw_fill_kernel(int nsteps, int upper, const int* __restrict__ step_q0,
                                                      const int* __restrict__ dep_slot /* [S][n] */, int64_t n,
                                                      const double* __restrict__ M, const double* __restrict__ dinv,
                                                      unsigned char* __restrict__ stream)
{
    using T = TwCfg<B, S, DINV>;
    constexpr int BB = B * B;
    const int64_t total = (int64_t)nsteps * T::NW * 32;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < total; tid += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(tid & 31), warp = (int)((tid >> 5) % T::NW);
        const int st = (int)(tid / (T::NW * 32));
        const int rw = lane / B, r = lane - rw * B;
        const int rho = warp * T::RPW + rw;
        const int q0 = step_q0[st], count = step_q0[st + 1] - q0;
        const bool active = rw < T::RPW && rho < count;
        const int q = q0 + rho;
        double av[2 * T::NP];
#pragma unroll
        for (int k = 0; k < 2 * T::NP; ++k)
            av[k] = 0.0;
        if (active) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int gslot = dep_slot[(size_t)s * n + q];
                if (gslot >= 0) {
#pragma unroll
                    for (int c = 0; c < B; ++c)
                        av[s * B + c] = __ldcs(M + elem_index_slot<BB>(gslot, r * B + c));
                }
            }
            if (DINV) {
#pragma unroll
                for (int c = 0; c < B; ++c)
                    av[S * B + c] = dinv[(size_t)q * BB + r * B + c];
            }
        }
        unsigned char* rec = stream + (size_t)(upper ? nsteps - 1 - st : st) * T::kRecBytes;
        double2* dst = reinterpret_cast<double2*>(rec + T::kValOff) + (size_t)(warp * T::NP) * 32 + lane;
#pragma unroll
        for (int k = 0; k < T::NP; ++k)
            dst[(size_t)k * 32] = make_double2(av[2 * k], av[2 * k + 1]);
    }
}

} // namespace opmb200
