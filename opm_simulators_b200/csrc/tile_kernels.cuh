// tile_kernels.cuh -- tile walkers: the triangular sweeps of schedule "tiles" (DESIGN.md section 6).
//
// The level schedule pays one L2 round trip per level (C3: 363 levels x ~1 us).  Here ONE CTA walks
// ONE CHUNK of rows -- on a box grid a tile of TJ x TK grid lines, one line per row of a step, the
// lines skewed so that step s holds cell i = s - lj - lk of line (lj, lk) -- step after step.
// Warp roles of a CTA (all data of a step meet in one STAGE of shared memory):
//   * 4 COMPUTE warps, B lanes per block row (one lane per row of the b x b blocks), R = 4 x 32/b rows
//     per step.  Their input comes from shared memory only.  The dependent chain of a step is: the
//     neighbours' results from the chunk's shared-memory ring (or the stage's external slots) -> 9 DFMA
//     -> one shuffle round -> 3 DFMA -> ring -> one named barrier; the block values and the dependency
//     codes of step t+1 are loaded while step t computes; the results are stored to global memory from
//     the registers BEHIND the barrier (a barrier waits for the acknowledgement of strong stores in
//     front of it: +200 cycles per step, measured), strongly only where another chunk polls them.
//   * a LOADER warp (one lane): per step ONE TMA bulk copy (cp.async.bulk -> UBLKCP) of the step's
//     record -- header, publish mask, external list, dependency codes, the rows' block values lane by
//     lane, Dinv; one contiguous piece of a per-sweep stream -- completion on the stage's "data"
//     mbarrier; the stream is pulled into the L2 a few steps ahead (UBLKPF).
//   * POLL warps, each serving every n-th step: the step's right-hand side (from the solver vector or
//     the lower sweep's records, in the L2 since the chunk started) and the dependencies that cross a
//     chunk boundary, one lane per dependency: strong loads of sentinel-armed dependency records (as in
//     sweep_kernel) until they are valid; both parked in the stage, signalled on its "ext" mbarrier.
// CTAs are persistent and take chunks through the in-order ticket, so a chunk only ever waits for
// chunks that are running or done.  The arithmetic per row is the level kernels' (same blocks, same
// order of the fused multiply-adds): both schedules give bit-identical preconditioner applications.
#pragma once
#include "kernels.cuh"
#include "layout.hpp"

namespace opmb200 {

constexpr int kTwMaxPollWarps = 4;
constexpr int kTwMaxRhsWarps = 4;
// launch bound: compute warps, loader, publisher, right-hand-side warps, poll warps; the launch picks the last two.
// ONE CTA per SM: two co-resident tile walkers slow each other down by more than they gain (C3: 250 us with two
// per SM, 205 us with one)
constexpr int kTwMaxThreads = (kTwWarps + 2 + kTwMaxRhsWarps + kTwMaxPollWarps) * 32;
#ifndef TW_NO_EARLY_TEST
#define TW_NO_EARLY_TEST 0
#endif
#ifndef TW_EXT_FLAG
#define TW_EXT_FLAG 0 // 1: "externals parked" is a pair of shared-memory words the compute warps spin on, 0: an mbarrier
#endif

template <int B, int S, bool DINV, bool UPPER>
struct TwCfg {
    static constexpr int NW = kTwWarps;               // compute warps
    static constexpr int RPW = kTwRows / NW;          // rows per warp: 8 (b lanes each; b = 3: 24 of a warp's 32 lanes work)
    static constexpr int LW = RPW * B;                // working lanes per warp
    static constexpr int R = kTwRows;                 // rows per step == one SELL slice
    static constexpr int RP = R;
    static constexpr int W = Rec<B>::W;               // doubles per dependency record in global memory
    static constexpr int NV = S * B + (DINV ? B : 0); // doubles per lane and step
    static constexpr int NP = (NV + 1) / 2;           // ... as 16-byte pairs
    static constexpr int RING = (4 * R <= 256) ? 256 : 512; // == Layout::tw_ring
    // record (global memory) == head of a stage (shared memory)
    static constexpr int kHdrOff = 0;                                // int4 {q0, count, n_ext, flags}
    static constexpr int kPubOff = 16;                               // 128 bits: rows whose result somebody polls in this sweep
    static constexpr int kArmOff = 32;                               // 128 bits: rows somebody polls in the other sweep
    static constexpr int kExtPosOff = 48;                            // kTwMaxExt positions
    static constexpr int kCodeOff = kExtPosOff + kTwMaxExt * 4;      // [S][RP] dependency codes
    static constexpr int kValOff = (kCodeOff + S * RP * 4 + 15) & ~15; // [NW][NP][LW] double2: no bytes for idle lanes
    static constexpr int kRecBytes = kValOff + NW * NP * LW * 16;
    // what the loader and the poll warps add to a stage.  Right-hand side: upper = the lower sweep's records
    // [RP][W]; lower = the step's runs of the component-major solver vector [B][RP]
    static constexpr int kRhsOff = kRecBytes;
    static constexpr int kRhsBytes = UPPER ? RP * W * 8 : B * RP * 8;
    static constexpr int kExtValOff = kRhsOff + kRhsBytes;           // [kTwMaxExt][4] doubles
    static constexpr int kStageBytes = (kExtValOff + kTwMaxExt * 32 + 127) & ~127;
    // stages: what fits in ~128 KB, and no more than the ring can back -- a stage is free once the publisher warp has
    // stored its step from the ring, and a ring slot is overwritten RING / 32 steps later at the earliest (a step
    // has at most 32 rows, csrc/analysis.cpp)
    static constexpr int kStagesRaw = 131072 / kStageBytes;
    static constexpr int kStagesCap = RING / 32 < 8 ? RING / 32 : 8;
    static constexpr int kStages = kStagesRaw < 2 ? 2 : (kStagesRaw > kStagesCap ? kStagesCap : kStagesRaw);
    static constexpr int kRingOff = kStages * kStageBytes;           // [RING][4] doubles
    static constexpr int kZeroOff = kRingOff + RING * 32;            // one all-zero record
    static constexpr int kScratchOff = kZeroOff + 32;                // where idle lanes and rows beyond a step's count store
    static constexpr int kBarOff = kScratchOff + 32;                 // kStages "data", then kStages "ext" mbarriers
    static constexpr int kCtlOff = kBarOff + 16 * kStages;           // int[8], see TwCtl
    static constexpr int kFlagOff = kCtlOff + 32;                    // [kStages][2] step tokens: right-hand side parked, externals parked
    static constexpr int kStampOff = kFlagOff + 8 * kStages;         // OPMB200_PROFILE: [kStages][2] arrival times (clock64)
    static constexpr int kSmemBytes = kStampOff + 16 * kStages + 16;
};
// control words in shared memory
enum TwCtl { kCtlReleased = 0, kCtlRec0 = 1, kCtlSteps = 2, kCtlStop = 3, kCtlQ0 = 4, kCtlQ1 = 5, kCtlPublished = 6 };

struct TwArgs {
    int nchunks;
    const int* chunk_step0;      // [nchunks+1]
    int nsteps;
    const unsigned char* stream; // this sweep's step records in WALKING order (upper: reversed)
    const double* d;             // lower: right-hand side (component-major, two doubles of slack behind it)
    double* tmp;                 // dependency records of the lower sweep's result y
    double* vpoll;               // dependency records of the upper sweep's result
    double* v;                   // result, component-major like every solver vector
    int64_t n;
    int ghost_zero;              // ILU0 ghost rows: 1 = v enters as 0, 0 = keep v's input
    const int* step_q0;          // [nsteps+1] first position of every step (NOT in walking order)
    int prefetch;                // L2 look-ahead of the loader, in steps
    int debug;                   // OPMB200_TWDBG builds only: timing experiments that switch parts off (wrong results)
    int poll_warps;
    int rhs_warps;
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
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival once all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes)
{
    // cp.async.bulk.prefetch.L2 (SASS: UBLKPF.L2); address and size rounded to 16 bytes
    const unsigned long long a = (unsigned long long)p;
    const unsigned long long a0 = a & ~15ull;
    const unsigned sz = (unsigned)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(sz) : "memory");
}

#ifdef OPMB200_TWDBG
#define TW_DBG(bit) (a.debug & (bit))
#else
#define TW_DBG(bit) false
#endif

#ifdef OPMB200_PROFILE
// in-kernel profile of the tile walkers: cycles per phase, accumulated in registers by lane 0 of one warp per
// role and flushed per chunk.  g_twp[0..7] compute phases, [8..15] loader phases, [16..23] poll phases,
// [24] steps, [25] chunks, [26] poll loads issued, [28..35] store warp phases
__device__ unsigned long long g_twp[64];
// hop anatomy: [40] polled records found valid, [41] ns from the producer's store to the consumer's valid sample
// (globaltimer stamps travel in the unused 4th word of a b = 3 record), [42] max of that, [43] steps published,
// [44] cycles from the compute warps' release to the publisher's store, [45] blocked steps, [46] cycles from the
// last arrival on the ext barrier to the compute warps' restart
__device__ __forceinline__ unsigned long long twp_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TWP_DECL long long twp_t0__ = clock64(); unsigned long long twp_acc__[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TWP_MARK(i)                                                                                                    \
    do {                                                                                                               \
        const long long now__ = clock64();                                                                             \
        twp_acc__[i] += (unsigned long long)(now__ - twp_t0__);                                                        \
        twp_t0__ = now__;                                                                                              \
    } while (0)
#define TWP_FLUSH(base, cond)                                                                                          \
    do {                                                                                                               \
        if (cond)                                                                                                      \
            for (int i__ = 0; i__ < 8; ++i__)                                                                          \
                atomicAdd(&g_twp[(base) + i__], twp_acc__[i__]);                                                       \
    } while (0)
#define TWP_COUNT(i, cond, v)                                                                                          \
    do {                                                                                                               \
        if (cond)                                                                                                      \
            atomicAdd(&g_twp[i], (unsigned long long)(v));                                                             \
    } while (0)
#else
#define TWP_DECL
#define TWP_MARK(i)
#define TWP_FLUSH(base, cond)
#define TWP_COUNT(i, cond, v)
#endif

// what a compute lane holds about a step before its dependencies are there
template <int B, int S, int NP>
struct TwStep {
    int q0, count, flags;
    unsigned xaddr[S]; // where dependency s is found (ring, the stage's external slots, the zero record)
    double av[2 * NP]; // this lane's block values (+ Dinv): in registers a step early, off the dependent chain
};

// ---- compute warps -------------------------------------------------------------------------------
// A lone warp per scheduler issues in order: a dependent DFMA chain overlaps with other work only if that
// work sits BETWEEN the DFMAs of the same basic block.  So the step proper is one branch-free block -- the
// neighbours' values, the next step's loads (block values, dependency codes -> addresses by arithmetic,
// no branches), the chain, the ring store -- and everything that may loop (the barrier waits) is in front.
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_compute(const TwArgs& a, unsigned char* smem, int g0, int ns, int warp, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    using Step = TwStep<B, S, T::NP>;
    constexpr int NS = T::kStages;
    const int rw = lane / B, r = lane - rw * B;
    const bool lane_ok = rw < T::RPW;
    const int rho = lane_ok ? warp * T::RPW + rw : 0; // idle lanes shadow row 0 of the step: computed, never stored
    const int src0 = lane_ok ? rw * B : 0;            // first lane of this row
    const unsigned smem_s = smem_u32(smem);
    const unsigned ring_s = smem_s + T::kRingOff;
    unsigned long long* data_bar = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    unsigned long long* ext_bar = data_bar + NS;
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const bool zero_ghosts = !UPPER && ILU0 && a.ghost_zero;
    // per-lane constants of the look-ahead
    const unsigned code_o = T::kCodeOff + (unsigned)rho * 4;
    const unsigned val_o = T::kValOff + (unsigned)(warp * T::NP * T::LW + (lane_ok ? lane : 0)) * 16;
    const unsigned in_o = T::kRhsOff + (unsigned)(UPPER ? rho * T::W + r : r * T::RP + rho) * 8;
    TWP_DECL;

    // everything about the step in stage `st` that does not depend on the steps before it: loads and
    // arithmetic only.  Harmless on a stage that holds no step yet (the addresses are masked into range).
    auto look_ahead = [&](int st, Step& N) {
        const unsigned sb = smem_s + (unsigned)st * T::kStageBytes;
        N.q0 = lds_s32(sb + T::kHdrOff);
        N.count = lds_s32(sb + T::kHdrOff + 4);
        N.flags = lds_s32(sb + T::kHdrOff + 12);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            // kTwRing + index: the chunk's ring (index RING = the zero record behind it: "no dependency");
            // else: slot of the stage's external values
            const int c = lds_s32(sb + code_o + (unsigned)(s * T::RP) * 4);
            N.xaddr[s] = ((c & kTwRing) ? ring_s : sb + T::kExtValOff) + (unsigned)(c & (2 * T::RING - 1)) * 32;
        }
#pragma unroll
        for (int k = 0; k < T::NP; ++k)
            lds_v2(sb + val_o + (unsigned)k * T::LW * 16, N.av[2 * k], N.av[2 * k + 1]);
    };

    int st = g0 % NS;
    unsigned par = (unsigned)(g0 / NS) & 1u;
    Step SA, SB;
    mbar_wait(data_bar + st, par);
    look_ahead(st, SA);
    TWP_MARK(7);

    bool ext_ready = false;
    auto step = [&](int t, const Step& C, Step& N) {
        const int g = g0 + t;
        int st1 = st + 1;
        unsigned par1 = par;
        if (st1 == NS) {
            st1 = 0;
            par1 ^= 1u;
        }
        // ---- what may loop first: ONE barrier ---------------------------------------------------------
        // this step's right-hand side and externals are parked -- and the poll warp that says so has seen the NEXT
        // step's record land (a second test of a completed phase here cost ~100 cycles per step)
#if TW_EXT_FLAG
        {
            const unsigned fa = smem_s + T::kFlagOff + (unsigned)st * 8;
            int f0, f1;
            do {
                asm volatile("ld.volatile.shared.v2.s32 {%0,%1}, [%2];" : "=r"(f0), "=r"(f1) : "r"(fa) : "memory");
            } while (f0 != g + 1 || f1 != g + 1);
        }
#else
#ifdef OPMB200_PROFILE
        const long long tw0__ = clock64();
#endif
        if (!ext_ready) // (tested while the step before ran: a test of a completed phase costs 130-200 cycles)
            mbar_wait(ext_bar + st, par);
#endif
#ifdef OPMB200_PROFILE
        if (warp == 0 && lane == 0 && t >= 8) { // who was the step waiting for?  [36] rhs late, [37] externals late: counts; [38], [39]: cycles
            const long long tr = reinterpret_cast<volatile long long*>(smem + T::kStampOff)[st * 2];
            const long long te = reinterpret_cast<volatile long long*>(smem + T::kStampOff)[st * 2 + 1];
            const long long last = tr > te ? tr : te;
            if (last > tw0__) {
                atomicAdd(&g_twp[tr > te ? 36 : 37], 1ull);
                atomicAdd(&g_twp[tr > te ? 38 : 39], (unsigned long long)(last - tw0__));
                atomicAdd(&g_twp[45], 1ull);
                atomicAdd(&g_twp[46], (unsigned long long)(clock64() - last));
            }
        }
#endif
        TWP_MARK(t < 8 ? 6 : 0);
        // ---- the step proper: one basic block ------------------------------------------------------
        double x[S][B];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if constexpr (B == 1) {
                x[s][0] = lds_f64(C.xaddr[s]);
            } else {
                lds_v2(C.xaddr[s], x[s][0], x[s][1]);
                if constexpr (B == 3)
                    x[s][2] = lds_f64(C.xaddr[s] + 16);
                if constexpr (B == 4)
                    lds_v2(C.xaddr[s] + 16, x[s][2], x[s][B - 1]);
            }
        }
        // this lane's right-hand side (its block values came with the look-ahead of the step before)
        const unsigned sb0 = smem_s + (unsigned)st * T::kStageBytes;
        const double(&av)[2 * T::NP] = C.av;
        double in = lds_f64(sb0 + in_o);
        in = (zero_ghosts && (C.flags & 1)) ? 0.0 : in; // ParallelOverlappingILU0 never touches ghost rows
        look_ahead(st1, N);
#if !TW_EXT_FLAG
        // are the next step's right-hand side and externals parked already?  The answer travels while the chain runs
        ext_ready = (t + 1 < ns) && !TW_NO_EARLY_TEST && mbar_test_wait(ext_bar + st1, par1);
#endif
        // the row: same blocks, same order of operations as sweep_kernel
#ifdef TW_EXPERIMENT_PRESCALED // timing experiment: pre-scaled blocks, three short chains, no shuffle round (wrong results)
        double acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            acc[s] = s == 0 ? in : 0.0;
#pragma unroll
            for (int c = 0; c < B; ++c)
                acc[s] -= av[s * B + c] * x[s][c];
        }
        double res = acc[0];
#pragma unroll
        for (int s = 1; s < S; ++s)
            res += acc[s];
#else
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
#endif
        res = guard(res);
        // (no branch: a divergent store costs a reconvergence per step; lanes without a row store into a scratch record)
        const bool active = lane_ok && rho < C.count;
        const int q = C.q0 + rho;
        sts_f64((active ? ring_s + (unsigned)(q & (T::RING - 1)) * 32 : smem_s + T::kScratchOff) + r * 8, res);
        TWP_MARK(1);
        named_bar_sync(1, T::NW * 32); // the ring writes are visible to the four warps; everybody is done with the stage
#ifdef OPMB200_PROFILE
        if (threadIdx.x == 0)
            reinterpret_cast<volatile long long*>(smem + T::kStampOff)[2 * NS] = clock64();
#endif
        if (threadIdx.x == 0)
            ctl[kCtlReleased] = g + 1; // the publisher warp stores the step's results from the ring and frees the stage
        TWP_MARK(2);
        TWP_MARK(3);
        st = st1;
        par = par1;
    };

    for (int t = 0; t < ns; t += 2) {
        step(t, SA, SB);
        if (t + 1 < ns)
            step(t + 1, SB, SA);
    }
    TWP_FLUSH(0, warp == 0 && lane == 0);
    TWP_COUNT(24, warp == 0 && lane == 0, ns);
    TWP_COUNT(25, warp == 0 && lane == 0, 1);
}

// ---- loader warp: ONE TMA bulk copy of the step's record per step, completion on the stage's "data" mbarrier ----
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_loader(const TwArgs& a, unsigned char* smem, int rec0, int g0, int ns, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    constexpr int NS = T::kStages;
    unsigned long long* data_bar = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const int PF = a.prefetch;
    // the chunk's right-hand side is a contiguous range of positions: pull all of it into the L2 now (a
    // few tens of KB), the step records in batches a few steps ahead of the copies into shared memory --
    // a stage filled from HBM takes ~1.5 us, kStages of them in flight would cap a step at ~500 cycles
    if (!TW_DBG(16)) {
        const int qa = ctl[kCtlQ0], qb = ctl[kCtlQ1];
        if (UPPER) {
            if (lane == 0)
                l2_prefetch_bulk(a.tmp + (size_t)qa * T::W, (unsigned)(qb - qa) * T::W * 8);
        } else if (lane < B) {
            l2_prefetch_bulk(a.d + VIDX(a.n, qa, lane), (unsigned)(qb - qa) * 8);
        }
        if (lane == 0 && PF > 0)
            l2_prefetch_bulk(a.stream + (size_t)rec0 * T::kRecBytes, (unsigned)(min(PF + NS, ns) * T::kRecBytes));
    }
    if (lane != 0)
        return;
    TWP_DECL;
    const unsigned char* rec = a.stream + (size_t)rec0 * T::kRecBytes;
    for (int t = 0; t < ns; ++t, rec += T::kRecBytes) {
        const int g = g0 + t, st = g % NS;
        if (PF > 0 && (t & 3) == 0 && t + NS + PF < ns && !TW_DBG(16)) // four records per prefetch
            l2_prefetch_bulk(rec + (size_t)(NS + PF) * T::kRecBytes, (unsigned)(min(4, ns - t - NS - PF) * T::kRecBytes));
        TWP_MARK(0);
        while (ctl[kCtlPublished] < g + 1 - NS) {} // plain spin (a __nanosleep would cost a microsecond)
        TWP_MARK(1);
        const unsigned bytes = TW_DBG(4) ? 48u : (unsigned)T::kRecBytes; // timing experiment: header only
        mbar_expect_tx(data_bar + st, bytes);
        tma_load_1d(smem + (size_t)st * T::kStageBytes, rec, bytes, data_bar + st);
        TWP_MARK(2);
    }
    TWP_FLUSH(8, true);
}

// "step g's right-hand side / externals are parked in stage st": one arrival on the stage's "ext" mbarrier, or a
// shared-memory word (the lanes' stores ordered in front of it)
template <class T>
__device__ __forceinline__ void tw_signal(unsigned char* smem, int st, int g, int which, int lane)
{
    __syncwarp();
    if (lane == 0) {
#ifdef OPMB200_PROFILE
        reinterpret_cast<volatile long long*>(smem + T::kStampOff)[st * 2 + which] = clock64();
#endif
#if TW_EXT_FLAG
        __threadfence_block();
        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(smem) + T::kFlagOff + (unsigned)(st * 8 + which * 4)), "r"(g + 1) : "memory");
#else
        unsigned long long* ext_bar = reinterpret_cast<unsigned long long*>(smem + T::kBarOff) + T::kStages;
        mbar_arrive(ext_bar + st); // release: the stores above are visible to whoever passes the barrier
#endif
    }
}

// ---- right-hand-side warps, each serving every n-th step: the step's right-hand side (from the solver vector or the
// lower sweep's records, in the L2 since the chunk started), loaded one own-step ahead, parked in the stage ----------
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_rhs(const TwArgs& a, unsigned char* smem, int rec0, int g0, int ns, int pw, int npw, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    constexpr int NS = T::kStages;
    unsigned long long* data_bar = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const unsigned char* rec = a.stream + (size_t)(rec0 + pw) * T::kRecBytes;
    constexpr int NR = (T::R + 31) / 32;      // rows per lane
    constexpr int NC = UPPER ? T::W : B;      // right-hand side words per row
    auto load_head = [&](const unsigned char* r, int4& hdr) {
        hdr = __ldg(reinterpret_cast<const int4*>(r)); // q0, count, n_ext, flags
    };
    auto load_rhs = [&](const int4& hdr, double (&rhs)[NR][NC]) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int rho = lane + 32 * i;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                rhs[i][c] = 0.0;
            if (rho < hdr.y && !TW_DBG(8)) {
                if constexpr (UPPER) { // y_i of the lower sweep (complete: previous kernel)
                    const double2* p = reinterpret_cast<const double2*>(a.tmp + (size_t)(hdr.x + rho) * T::W);
                    const double2 u0 = __ldcs(p);
                    rhs[i][0] = u0.x;
                    rhs[i][1] = u0.y;
                    if constexpr (T::W == 4) {
                        const double2 u1 = __ldcs(p + 1);
                        rhs[i][2] = u1.x;
                        rhs[i][NC - 1] = u1.y;
                    }
                } else { // ParallelOverlappingILU0 never touches ghost rows: they keep v's input (or enter as 0)
                    const double* src = (ILU0 && (hdr.w & 1)) ? a.v : a.d;
#pragma unroll
                    for (int c = 0; c < B; ++c)
                        rhs[i][c] = __ldcs(src + VIDX(a.n, hdr.x + rho, c));
                }
            }
        }
    };
    int4 hdr = make_int4(0, 0, 0, 0), hdr1 = hdr, hdr2 = hdr;
    double rhs[NR][NC], rhs1[NR][NC];
    if (pw < ns) {
        load_head(rec, hdr);
        load_rhs(hdr, rhs);
    }
    if (pw + npw < ns)
        load_head(rec + (size_t)npw * T::kRecBytes, hdr1);
    TWP_DECL;
    for (int t = pw; t < ns; t += npw, rec += (size_t)npw * T::kRecBytes) {
        const int g = g0 + t, st = g % NS;
        if (t + 2 * npw < ns)
            load_head(rec + (size_t)2 * npw * T::kRecBytes, hdr2);
        if (t + npw < ns) // the next step this warp serves
            load_rhs(hdr1, rhs1);
        TWP_MARK(0);
        while (ctl[kCtlPublished] < g + 1 - NS) {}
        if (t + 1 < ns) { // the compute warps decode the next step's record while they run this one
            const int g1 = g + 1;
            // A parity wait is only meaningful against the phase right before: the previous step of THAT stage
            // (g1 - NS) must be done with, so that its record has landed.  The gate above says so for this step's
            // stage only -- for the kStages-th step of a chunk the next stage still holds the chunk's FIRST record,
            // possibly in flight, the wait for the phase after it returned at once and, once in ~1e5 chunks, the
            // compute warps decoded the stale record (a wrong preconditioner application in ~1 % of the calls on C3,
            // scripts/determinism_probe2.py).  The loader cannot issue record g1 before this condition either.
            while (ctl[kCtlPublished] < g1 + 1 - NS) {}
            mbar_wait(data_bar + g1 % NS, (unsigned)(g1 / NS) & 1u);
        }
        TWP_MARK(1);
        const unsigned sb = smem_u32(smem + (size_t)st * T::kStageBytes);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int rho = lane + 32 * i;
            if (rho < T::RP) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    sts_f64(sb + T::kRhsOff + (unsigned)(UPPER ? rho * T::W + c : c * T::RP + rho) * 8, rhs[i][c]);
            }
        }
        tw_signal<T>(smem, st, g, 0, lane);
        TWP_MARK(2);
        hdr1 = hdr2;
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c)
                rhs[i][c] = rhs1[i][c];
    }
    TWP_FLUSH(32, lane == 0 && pw == 0);
}

// ---- poll warps, each serving every n-th step: the dependencies the chunk's ring does not serve, one lane per
// dependency: strong loads of the sentinel-armed records (as in sweep_kernel) until they are valid.  Nothing else is in
// flight in these warps -- a sample that waits behind other loads is a longer hop from chunk to chunk -- and the list
// of positions comes from the step's record in shared memory.
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_poller(const TwArgs& a, unsigned char* smem, int g0, int ns, int pw, int npw, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    constexpr int NS = T::kStages;
    unsigned long long* data_bar = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    const double* out = UPPER ? a.vpoll : a.tmp;
    TWP_DECL;
    for (int t = pw; t < ns; t += npw) {
        const int g = g0 + t, st = g % NS;
        const unsigned sb = smem_u32(smem + (size_t)st * T::kStageBytes);
        // the stage's previous step is done with (so its record HAS landed: a parity wait is only meaningful
        // against the phase right before -- a warp that serves every n-th step does not see every phase), then
        // this step's record: its list of positions is there
        while (ctl[kCtlPublished] < g + 1 - NS) {}
        mbar_wait(data_bar + st, (unsigned)(g / NS) & 1u);
        // one dependency per lane; a step with more than 32 of them (kTwMaxExt = 64: the steps of a rank's first plane
        // also read the ghost plane below) takes a second round -- rare, and the ghost values are there already
        const int n_ext = lds_s32(sb + T::kHdrOff + 8);
        TWP_MARK(0);
        for (int e0 = 0; e0 < n_ext && !TW_DBG(2); e0 += 32) {
            const int pos = lds_s32(sb + T::kExtPosOff + (e0 + lane) * 4);
            if (pos < 0)
                continue;
            double x[B];
            int tries = 0;
            // every word validates itself against the sentinel.  No nap between the samples: the wait IS the hop from
            // chunk to chunk, and a __nanosleep costs about a microsecond
#ifdef OPMB200_PROFILE
            if constexpr (B == 3) {
                double w3;
                do {
                    ++tries;
                    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                 : "=d"(x[0]), "=d"(x[1]), "=d"(x[B - 1]), "=d"(w3)
                                 : "l"(out + (size_t)pos * 4)
                                 : "memory");
                } while (!rec_valid<B>(x));
                if (tries > 1) { // the hops that were waited for
                    const unsigned long long dt = twp_globaltimer() - (unsigned long long)__double_as_longlong(w3);
                    if (dt < 1000000ull) {
                        atomicAdd(&g_twp[40], 1ull);
                        atomicAdd(&g_twp[41], dt);
                        atomicMax(&g_twp[42], dt);
                    }
                }
            } else
#endif
            do {
                ++tries;
                rec_load_strong<B>(out, (size_t)pos, x);
            } while (!rec_valid<B>(x));
            TWP_COUNT(26, lane == 0 && pw == 0 && e0 == 0, tries);
#pragma unroll
            for (int c = 0; c < B; ++c)
                sts_f64(sb + T::kExtValOff + (e0 + lane) * 32 + c * 8, x[c]);
        }
        TWP_MARK(1);
        tw_signal<T>(smem, st, g, 1, lane);
        TWP_MARK(2);
    }
    TWP_FLUSH(16, lane == 0 && pw == 0);
}

// ---- publisher warp: stores a step's results from the ring as soon as the compute warps are done with it ------
// One lane per row (a step has at most 32 rows).  Rows another chunk polls first, with a strong store -- that store
// is the hop from chunk to chunk; the rest of y (read by the upper sweep, a later kernel), the result vector and the
// sentinels of the rows polled in the other sweep follow as plain stores.  Then the stage is free.  Inside the
// compute warps the same stores cost ~140 cycles per step of a chain that is bound by its instruction count.
template <int B, int S, bool ILU0, bool UPPER>
__device__ __forceinline__ void tw_publisher(const TwArgs& a, unsigned char* smem, int g0, int ns, int lane)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    constexpr int NS = T::kStages;
    const unsigned smem_s = smem_u32(smem);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    double* out = UPPER ? a.vpoll : a.tmp;
    double* arm = UPPER ? a.tmp : a.vpoll; // lower: arm the upper sweep's records; upper: re-arm for the next apply
    TWP_DECL;
    for (int t = 0; t < ns; ++t) {
        const int g = g0 + t, st = g % NS;
        const unsigned sb = smem_s + (unsigned)st * T::kStageBytes;
        // the step's header before the wait for its results (the record landed while the compute warps were busy
        // with the step before)
        mbar_wait(reinterpret_cast<unsigned long long*>(smem + T::kBarOff) + st, (unsigned)(g / NS) & 1u);
        const int q0 = lds_s32(sb + T::kHdrOff), count = lds_s32(sb + T::kHdrOff + 4);
        const bool polled = (lds_s32(sb + T::kPubOff) >> lane) & 1, armed = (lds_s32(sb + T::kArmOff) >> lane) & 1;
        while (ctl[kCtlReleased] < g + 1) {}
#ifdef OPMB200_PROFILE
        if (lane == 0) {
            atomicAdd(&g_twp[43], 1ull);
            atomicAdd(&g_twp[44], (unsigned long long)(clock64() - reinterpret_cast<volatile long long*>(smem + T::kStampOff)[2 * NS]));
        }
#endif
        TWP_MARK(0);
        if (lane < count && !TW_DBG(1)) {
            const int q = q0 + lane;
            const unsigned ra = smem_s + T::kRingOff + (unsigned)(q & (T::RING - 1)) * 32;
            double x[B];
            if constexpr (B == 1) {
                x[0] = lds_f64(ra);
            } else {
                lds_v2(ra, x[0], x[1]);
                if constexpr (B == 3)
                    x[2] = lds_f64(ra + 16);
                if constexpr (B == 4)
                    lds_v2(ra + 16, x[2], x[B - 1]);
            }
#ifdef OPMB200_PROFILE
            if (polled && B == 3) {
                double* pr = out + (size_t)q * 4;
                asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(pr), "d"(x[0]), "d"(x[1]), "d"(x[B - 1]),
                             "d"(__longlong_as_double((long long)twp_globaltimer()))
                             : "memory");
            } else
#endif
            if (polled)
                rec_store_strong<B>(out, (size_t)q, x);
            else if (!UPPER)
                rec_store_weak<B>(out, (size_t)q, x);
            if (UPPER) {
#pragma unroll
                for (int c = 0; c < B; ++c)
                    a.v[VIDX(a.n, q, c)] = x[c];
            }
            if (armed) {
                double sx[B];
#pragma unroll
                for (int c = 0; c < B; ++c)
                    sx[c] = sentinel();
                rec_store_weak<B>(arm, (size_t)q, sx);
            }
        }
        __syncwarp();
        if (lane == 0)
            ctl[kCtlPublished] = g + 1; // the stage may be refilled, the ring slots overwritten (8 steps from now at the earliest)
        TWP_MARK(1);
    }
    TWP_FLUSH(28, lane == 0);
}

template <int B, int S, bool ILU0, bool UPPER>
__global__ void __launch_bounds__(kTwMaxThreads, 1) tw_sweep_kernel(TwArgs a)
{
    constexpr bool DINV = !(ILU0 && !UPPER);
    using T = TwCfg<B, S, DINV, UPPER>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + T::kBarOff);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem + T::kCtlOff);
    if (threadIdx.x == 0) {
        for (int i = 0; i < T::kStages; ++i) {
            mbar_init(bars + i, 1u);                // data: the loader's expect_tx
            mbar_init(bars + T::kStages + i, 2u);   // ext: the right-hand-side warp and the poll warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 8 + 2 * T::kStages; ++i) // control words and step tokens
            ctl[i] = 0;
    }
    // idle lanes, rows beyond a step's count and the look-ahead behind a chunk's last step read words nobody
    // wrote: all-zero stages decode to in-range addresses and finite values
    for (int i = threadIdx.x; i < T::kBarOff / 16; i += blockDim.x)
        reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes before the TMA writes
    __syncthreads();
    const bool skip = a.check_done && a.sc->done;
    const int npw = a.poll_warps, nrw = a.rhs_warps;
    int g = 0; // steps walked by this CTA so far: stage = g % kStages, phase = (g / kStages) & 1
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned int tk = atomicAdd(a.ticket.next, 1u);
            if (skip || tk >= (unsigned)a.nchunks) {
                ctl[kCtlStop] = 1;
            } else {
                const int c = UPPER ? a.nchunks - 1 - (int)tk : (int)tk;
                const int s0 = a.chunk_step0[c], s1 = a.chunk_step0[c + 1];
                ctl[kCtlRec0] = UPPER ? a.nsteps - s1 : s0; // first record of the chunk in walking order
                ctl[kCtlSteps] = s1 - s0;
                ctl[kCtlQ0] = a.step_q0[s0]; // the chunk's positions
                ctl[kCtlQ1] = a.step_q0[s1];
            }
        }
        __syncthreads();
        if (ctl[kCtlStop])
            break;
        const int rec0 = ctl[kCtlRec0], ns = ctl[kCtlSteps];
        if (warp < T::NW)
            tw_compute<B, S, ILU0, UPPER>(a, smem, g, ns, warp, lane);
        else if (warp == T::NW)
            tw_loader<B, S, ILU0, UPPER>(a, smem, rec0, g, ns, lane);
        else if (warp == T::NW + 1)
            tw_publisher<B, S, ILU0, UPPER>(a, smem, g, ns, lane);
        else if (warp < T::NW + 2 + nrw)
            tw_rhs<B, S, ILU0, UPPER>(a, smem, rec0, g, ns, warp - T::NW - 2, nrw, lane);
        else
            tw_poller<B, S, ILU0, UPPER>(a, smem, g, ns, warp - T::NW - 2 - nrw, npw, lane);
        g += ns;
        __syncthreads(); // everybody has read the control words; the chunk is finished and published
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
    using T = TwCfg<B, S, DINV, true>; // the record layout does not depend on the direction
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
        if (rw < T::RPW) { // the record holds the working lanes only
            double2* dst = reinterpret_cast<double2*>(rec + T::kValOff) + (size_t)(warp * T::NP) * T::LW + lane;
#pragma unroll
            for (int k = 0; k < T::NP; ++k)
                dst[(size_t)k * T::LW] = make_double2(av[2 * k], av[2 * k + 1]);
        }
    }
}

} // namespace opmb200
