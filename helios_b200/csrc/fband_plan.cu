// Planned two-stream sweeps (round 2): one warp owns a column (or a pair of columns), everything a CTA needs for its NEXT
// tile is staged into shared memory while it sweeps the current one, and the passes run from registers.
//
// Why a plan.  Between two opacity refreshes (10 RT iterations, C:860) only the Planck terms change.  The sweep of
// K:1366-1799 is, per (half-)layer, the affine step
//      F_out = a F_in - b F_opp + s,        a = P/M,  b = N/M,
//      isothermal     (K:1443-1451):  s_down = k0d + k1 B_layer,            s_up = k0u + k1 B_layer
//      non-isothermal (K:1640-1795):  s_down = k0d + k1 B_layer + k2 B_int, s_up = k0u + k2 B_layer + k1 B_int
// with k0 = beam source / M and k1, k2 = products of 1/M, the source factor 2 pi eps (1-w0)/(E-w0) and the Planck
// weights of the half-layer.  (The up-going weights ARE the down-going ones swapped, bit for bit: (M-P)-N == -(N+(P-M)),
// so the six constants [a, b, k0d, k0u, k1, k2] carry what round 1 stored as eight.)  The plan-build kernels evaluate
// them once per refresh with the rounding-exact building blocks of sweep_math.cuh; with the beam known to be zero the two
// k0 rows are not stored at all.
//
// Plan layout = the shared-memory image of a warp's tile.  A warp tile is CPW = 32/LPC columns; lane = (column, chunk of
// CH layers).  Its block is NR*CH rows of RL = CPW*RS doubles (RS = active lanes per column, rounded to even; a lane's
// element sits at column*RS + chunk), row (k*NR + j) = constant j of the lane's k-th layer, followed by 4 doubles of
// per-column surface constants.  The block is contiguous in HBM, so ONE cp.async.bulk (TMA, SASS UBLKCP) per warp and tile
// brings it in, completion on a warp-private mbarrier.  The flux arrays keep the reference's [interface][column] layout
// (K:1076) -- one 8-byte piece per cache line for a lane that owns a column -- so the CTA (4 warps = 4 or 8 adjacent
// columns) moves them cooperatively, lanes along the columns, full 32 / 64-byte sectors, through a staged block in shared
// memory (see CtaShape).  The copies for tile n+1 are issued as soon as tile n has been lifted into registers, so the memory
// round trip overlaps the dependent sweep arithmetic.
//
// The passes (phase B of k_fband_wp): every lane composes the affine map of its chunk, a Kogge-Stone shuffle scan over
// the lanes of the column hands it the flux entering the chunk, the lane walks its layers.  Cells outside the column are
// identity steps BY DATA (a = 1, b = 0, s = 0: 1*F + 0 == F exactly), so there is no "inside" select.  The A parts of the
// scans are pass-invariant and hoisted (scan_setup); the reference's tiny-value clean-up is replaced by a sign-bit vote
// per tile with an exact redo (clean<>).  Measured on B200: the first build of this file, one warp per
// column with per-lane 8-byte copies, spent 16 of its 46 us per C2 solve in the flux stores and 8 us in the flux loads
// (25 wavefronts per LSU instruction); the cooperative staging brought the solve to 44 us.  The global side of the staging
// then took it to 38 us: 16-byte cp.async.cg column pairs instead of 8-byte copies through L1 (flux_pairs), Planck values
// staged once per CTA when its columns share a bin (one_bin), tile-invariant index arithmetic formed at its use instead
// of living in registers across the passes (FluxMap / PairMap), the batch's converged flag requested before the waits.
#include "common.cuh"
#include "sweep_math.cuh"
#include "fband_plan.cuh"
#include <cstdlib>
#include <type_traits>

namespace {

struct PlanScalars {
    double Rstar, a, f_factor;
    int nint, nbin, ny, dir_beam, npass, nch, rs, nbatch;
    const int* done;  // batch: converged atmospheres are skipped (their fluxes stay as they are)
    int cp;             // column pitch of the staged flux rows (CtaShape::cpitch), formed by the launcher
    unsigned ny_magic;  // floor(2^32 / ny) + 1: column -> bin by one multiply-high (exact below 2^32 / ny columns)
    unsigned nct_magic;  // the same for CTA tile -> atmosphere (0: divide)
#ifdef HELIOS_ABLATE
    int ablate;  // experiment builds only (scripts/exp_ablate.sh): bit mask of parts of the sweep to leave out
    int skew_ns;  // experiment: co-resident CTA r (blockIdx.x / SMs) starts r * skew_ns late
    int nsm;
#endif
};
#ifdef HELIOS_ABLATE
#define ABL(bit) ((s.ablate & (bit)) != 0)
#else
#define ABL(bit) false
#endif

// CTA tile -> atmosphere of a batch: nothing for one atmosphere, one multiply-high when the launcher found it exact
__device__ __forceinline__ unsigned atm_of(unsigned ct, unsigned nct, const PlanScalars& s) {
    if (s.nbatch == 1) return 0u;
    return s.nct_magic != 0u ? __umulhi(ct, s.nct_magic) : ct / nct;
}
__device__ __forceinline__ int bin_of(int col, const PlanScalars& s) {
    return s.ny == 1 ? col : (int)__umulhi((unsigned)col, s.ny_magic);
}

// ---------------------------------------------------------------- async-copy plumbing (PTX) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {  // .cg: L2 -> shared, no L1 line allocated
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- the passes -------------------
// Scans.  The A parts of the affine maps (products of the a's of a chunk) do not depend on the fluxes, so everything the
// Kogge-Stone scans need from them is formed ONCE per tile: the multiplier every step applies to the incoming B part
// (0 where the step does not apply to this lane -- that replaces the guard and its selects) and the prefix product that
// multiplies the boundary flux.  A pass then shuffles and multiply-adds the B parts only: 2 SHFL + 1 DFMA per step.
// The multipliers live in warp-private shared memory, [step][down | up][lane].
template <int LPC>
struct Log2;
template <>
struct Log2<16> {
    static constexpr int v = 4;
};
template <>
struct Log2<32> {
    static constexpr int v = 5;
};

template <int LPC>
__device__ __forceinline__ void scan_setup(double Adn, double Aup, int sl, int nch, int lane, volatile double* mult,
                                           double& scA_dn, double& scA_up) {
#pragma unroll
    for (int r = 0, d = 1; d < LPC; r++, d <<= 1) {
        const double odn = __shfl_down_sync(0xffffffffu, Adn, d, LPC);
        const double oup = __shfl_up_sync(0xffffffffu, Aup, d, LPC);
        const bool okd = sl + d < nch, oku = sl >= d;
        mult[(2 * r) * 32 + lane] = okd ? Adn : 0.0;
        mult[(2 * r + 1) * 32 + lane] = oku ? Aup : 0.0;
        if (okd) Adn = Adn * odn;
        if (oku) Aup = Aup * oup;
    }
    scA_dn = Adn;
    scA_up = Aup;
}
template <int LPC>
__device__ __forceinline__ double scan_dn(double mB, int lane, const volatile double* mult) {
#pragma unroll
    for (int r = 0, d = 1; d < LPC; r++, d <<= 1) mB = __fma_rn(mult[(2 * r) * 32 + lane], __shfl_down_sync(0xffffffffu, mB, d, LPC), mB);
    return mB;
}
template <int LPC>
__device__ __forceinline__ double scan_up(double mB, int lane, const volatile double* mult) {
#pragma unroll
    for (int r = 0, d = 1; d < LPC; r++, d <<= 1) mB = __fma_rn(mult[(2 * r + 1) * 32 + lane], __shfl_up_sync(0xffffffffu, mB, d, LPC), mB);
    return mB;
}

// The reference cleans every new flux, fabs(f) < 1e-100 ? fabs(f) : f (K:1453 ...): 18 of the 26 cycles of a dependent
// walk step, although it changes a value only if its sign bit is set (and |f| < 1e-100).  EXACT = false: no clean-up,
// the sign bits of all results are OR-ed into `acc` (one LOP3, off the critical path); a tile whose `acc` stays positive
// in every lane -- practically all of them -- is bit-identical to the cleaned evaluation.  Otherwise the tile is redone
// with EXACT = true.
template <bool EXACT>
__device__ __forceinline__ double clean(double f, unsigned& acc) {
    if (EXACT) return tiny_to_abs(f);
    acc |= (unsigned)__double2hiint(f);
    return f;
}

// Shared-memory map of a CTA (doubles).  A CTA of WARPS warps owns NC = WARPS * CPW consecutive columns (a CTA tile).
//   per warp : [mbarrier 2][plan block PB][Planck values NB * rl][column constants 8]          rl = CPW * rs
//   per CTA  : [scan multipliers WARPS * 2 * NST * 32 | results of the downward sweep]  [previous fluxes]  [results up]
// The flux arrays are [interface][column] in HBM (K:1076): a warp that owns one column touches one 8-byte piece of a
// different cache line with every lane -- 25 wavefronts per load/store instruction, and the LSU serialises them (that,
// not HBM, bound the first build of this kernel: 16 of its 46 us were the stores, 8 us the flux loads).  So the fluxes
// move COOPERATIVELY: all threads of the CTA copy the [layer][NC columns] block of the tile with lanes along the columns
// (full 32 / 64-byte sectors), staged in shared memory in the order the lanes consume them, [array][k][column * (rs + 1) +
// chunk].
template <int CH, int NR, int NF, int NB, int NST, int CPW, int WARPS>
struct CtaShape {
    __host__ __device__ static int rl(int rs) { return rs * CPW; }
    __host__ __device__ static int pb(int rs) { return CH * NR * rl(rs) + 4; }
    __host__ __device__ static int warp_doubles(int rs) { return 2 + pb(rs) + NB * rl(rs) + 8; }
    // staged flux rows: the cooperative copies run with lanes along the NC columns (32 / NC chunk slots each), the
    // owners with lanes along the chunk slots of one column.  8-byte accesses are served per half-warp over 16 bank
    // pairs: the NC column starts must land 16 / NC pairs apart, i.e. pitch = (16 / NC) * odd  (NC = 4: 28 for 26
    // slots, NC = 8: 18 for 16) -- both patterns are then conflict-free
    __host__ __device__ static int cpitch(int rs) {
        constexpr int NCOL = WARPS * CPW;
        constexpr int st = NCOL >= 16 ? 1 : 16 / NCOL;
        int v = rs;
        while (v % (2 * st) != st) v++;
        return v;
    }
    __host__ __device__ static int frow(int rs) { return WARPS * CPW * cpitch(rs); }
    __host__ __device__ static int flux_doubles(int rs) { return NF * CH * frow(rs); }
    __host__ __device__ static int mult_doubles(int rs) {
        const int m = WARPS * 2 * NST * 32, f = flux_doubles(rs);
        return m > f ? m : f;
    }
    __host__ __device__ static int cta_doubles(int rs) {
        return WARPS * warp_doubles(rs) + mult_doubles(rs) + 2 * flux_doubles(rs);
    }
};

// cooperative copy of one flux array block: global [nlay rows][NC columns at col0] <-> staged [k][c * cp + lane].
// THREADS / NC equals the lanes per column, so thread (lane l, column c) moves exactly the CH interfaces lane l owns
// (rows l * CH + k): the staged offsets are o0 + k * frow and the global rows are consecutive -- no index arithmetic per
// element, and nothing about the map depends on the tile but the base pointer.  A warp request still covers NC
// adjacent columns (NC * 8 contiguous bytes) of 32 / NC rows.
template <int CH, int NC, int THREADS>
struct FluxMap {
    int o0;  // staged offset of the thread's first row
    int n;   // rows of this thread that exist (0..CH)
    int c;   // column within the CTA tile
    int r0;  // first global row
    __device__ __forceinline__ void init(int nlay, int cp) {
        int t = threadIdx.x;
        asm volatile("" : "+r"(t));  // formed where it is used: keeps the map out of the registers the passes need
        c = t % NC;
        const int l = t / NC;
        r0 = l * CH;
        o0 = c * cp + l;
        n = min(max(nlay - r0, 0), CH);
    }
};

template <int CH, int NC, int THREADS, bool LOAD>
__device__ __forceinline__ void flux_block(const FluxMap<CH, NC, THREADS>& m, double* __restrict__ g, unsigned stage_s,
                                           const double* stage, int ncol, int col0, int frow) {
    if (col0 + m.c >= ncol) return;
    double* __restrict__ p = g + col0 + m.c + (size_t)m.r0 * ncol;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        if (k < m.n) {
            if (LOAD) cp_async8(stage_s + (unsigned)(m.o0 + k * frow) * 8u, p + (size_t)k * ncol);
            else p[(size_t)k * ncol] = stage[m.o0 + k * frow];
        }
    }
}

// With an even column count the fluxes move as 16-byte pieces (two adjacent columns of one interface), staged side by
// side, [k][column pair][lane][2]; thread (kh, lane l, pair pc) moves the rows l * CH + k, k = kh, kh + 2, ...
//   in : cp.async.cg -- an 8-byte cp.async has to go through L1 (.ca), and a tile touches one 32-byte piece of 100+
//        different 128-byte lines per array, more lines than the L1 left beside the shared-memory carve-out holds: the
//        LSU stalled on line allocation (the flux loads cost 7 of 44 us, profiles/r2_sweep_ablation.txt);
//   out: LDS.128 + STG.128, half the instructions of the dependent load-store pairs of the 8-byte form.
// The owners' 8-byte accesses to this layout stride 16 bytes (2 wavefronts more per access).
template <int CH, int NC, int THREADS>
struct PairMap {
    int o0, r0, pc, kh;
    __device__ __forceinline__ void init(int cp) {
        constexpr int LPCX = THREADS / NC;
        int t = threadIdx.x;
        asm volatile("" : "+r"(t));  // (as in FluxMap)
        pc = t % (NC / 2);
        const int l = (t / (NC / 2)) % LPCX;
        kh = t / (THREADS / 2);
        r0 = l * CH;
        o0 = pc * 2 * cp + 2 * l;
    }
    // owner's view: element (column c of the CTA tile, lane slot sl) of a staged row
    __device__ static __forceinline__ int owner(int c, int sl, int cp) { return (c >> 1) * 2 * cp + 2 * sl + (c & 1); }
};

template <int CH, int NC, int THREADS, bool LOAD>
__device__ __forceinline__ void flux_pairs(const PairMap<CH, NC, THREADS>& m, double* __restrict__ g, unsigned stage_s,
                                           const double* stage, int nlay, int ncol, int col0, int frow) {
    const int col = col0 + 2 * m.pc;
    if (col >= ncol) return;
    double* __restrict__ p = g + col + (size_t)m.r0 * ncol;
#pragma unroll
    for (int k2 = 0; k2 < (CH + 1) / 2; k2++) {
        const int k = 2 * k2 + m.kh;
        if (k < CH && m.r0 + k < nlay) {
            if (LOAD) cp_async16(stage_s + (unsigned)(m.o0 + k * frow) * 8u, p + (size_t)k * ncol);
            else *reinterpret_cast<double2*>(p + (size_t)k * ncol) = *reinterpret_cast<const double2*>(stage + m.o0 + k * frow);
        }
    }
}

// =================================================================================================
// Isothermal layers: NR = 3 rows per layer [a, b, k1] (+ [k0d, k0u] with a beam), one previous flux, CH Planck values.
// =================================================================================================
template <int CH, int LPC, bool NOBEAM, int WARPS, int MINB, bool PAIRS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_sweep_iso(double* __restrict__ F_down, double* __restrict__ F_up, const double* __restrict__ planck_lay,
            const double* __restrict__ plan, const double* __restrict__ albedo, PlanScalars s) {
    constexpr int NR = NOBEAM ? 3 : 5;
    constexpr int CPW = 32 / LPC;
    constexpr int NC = WARPS * CPW;
    constexpr int NST = Log2<LPC>::v;
    constexpr int THREADS = WARPS * 32;
    using CS = CtaShape<CH, NR, 1, CH, NST, CPW, WARPS>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nint = s.nint, nlay = nint - 1, nch = s.nch, rs = s.rs, rl = rs * CPW, cp = s.cp, frow = NC * cp;
    const int ncol = s.nbin * s.ny;
    const unsigned ntw = (unsigned)(ncol + CPW - 1) / CPW;  // warp tiles per atmosphere (plan blocks)
    const unsigned nct = (unsigned)(ncol + NC - 1) / NC;    // CTA tiles per atmosphere
    const unsigned total = nct * (unsigned)s.nbatch;
    const int PB = CS::pb(rs);
    double* wst = smem + (size_t)warp * CS::warp_doubles(rs);
    double* pbuf = wst + 2;
    double* bbuf = pbuf + PB;
    double* cbuf = bbuf + CH * rl;
    double* cta = smem + (size_t)WARPS * CS::warp_doubles(rs);
    volatile double* mult = cta + warp * (2 * NST * 32);
    double* fout_dn = cta;  // aliases the multipliers: written only after every warp has finished its passes
    double* fin = cta + CS::mult_doubles(rs);
    double* fout_up = fin + CS::flux_doubles(rs);
    const unsigned bar = smem_u32(wst), pbuf_s = smem_u32(pbuf), bbuf_s = smem_u32(bbuf), cbuf_s = smem_u32(cbuf),
                   fin_s = smem_u32(fin);
    const int cw = lane / LPC, sl = lane % LPC;
    const bool act = sl < nch;
    const int lo = sl * CH;
    const int me = cw * rs + sl;                        // my element of a staged row of my warp
    // slots never written by the copies (layers beyond the column) must read as zeros
    for (int k = lane; k < CH * rl + 8; k += 32) bbuf[k] = 0.0;
    for (int k = threadIdx.x; k < CS::flux_doubles(rs); k += THREADS) fin[k] = 0.0;
    // PAIRS (the host checks: even column count, 16-byte aligned arrays): the fluxes move as 16-byte pieces
    const int fme = PAIRS ? PairMap<CH, NC, THREADS>::owner(warp * CPW + cw, act ? sl : 0, cp)
                          : (warp * CPW + cw) * cp + (act ? sl : 0);  // my element of a staged flux row of the CTA
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    const double* bbuf0 = smem + 2 + PB;  // warp 0's Planck block
    const unsigned bbuf0_s = smem_u32(bbuf0);
    auto one_bin = [&](int ctile) { return bin_of(ctile * NC, s) == bin_of(min(ctile * NC + NC - 1, ncol - 1), s); };

    auto issue = [&](unsigned ct) {
        const unsigned atm = atm_of(ct, nct, s);
        const int ctile = (int)(ct - atm * nct);
        const unsigned tile = min((unsigned)ctile * WARPS + warp, ntw - 1u);  // warps beyond the last column mirror it
        const int colc = min((int)tile * CPW + cw, ncol - 1);
        const int x = bin_of(colc, s);
        if (lane == 0) {
            fence_proxy_async();  // the generic-proxy reads of the previous tile are ordered before the bulk write
            mbar_expect_tx(bar, (unsigned)PB * 8u);
            bulk_g2s(pbuf_s, plan + ((size_t)atm * ntw + tile) * PB, (unsigned)PB * 8u, bar);
        }
        // The Planck values belong to the bin: when the CTA's columns all lie in one bin they are staged ONCE, in column
        // slot 0 of warp 0's block, each warp fetching every WARPS-th row (as in the non-isothermal sweep)
        const double* __restrict__ BL = planck_lay + (size_t)atm * (nlay + 2) * s.nbin + (size_t)x * (nlay + 2);
        const bool shx = one_bin(ctile);
        if (act && (!shx || cw == 0)) {
            const unsigned bdst = shx ? bbuf0_s + (unsigned)sl * 8u : bbuf_s + (unsigned)me * 8u;
#pragma unroll
            for (int k = 0; k < CH; k++)
                if (lo + k < nlay && (!shx || k % WARPS == warp)) cp_async8(bdst + (unsigned)(k * rl) * 8u, BL + lo + k);
        }
        if (sl < 3) cp_async8(cbuf_s + (unsigned)(cw * 4 + sl) * 8u, sl == 0 ? BL + nlay : (sl == 1 ? BL + nlay + 1 : albedo + x));
        if (PAIRS) {
            PairMap<CH, NC, THREADS> pmap;
            pmap.init(cp);
            flux_pairs<CH, NC, THREADS, true>(pmap, F_up + (size_t)atm * ncol * nint, fin_s, nullptr, nlay, ncol, ctile * NC, frow);
        } else {
            FluxMap<CH, NC, THREADS> fmap;
            fmap.init(nlay, cp);
            flux_block<CH, NC, THREADS, true>(fmap, F_up + (size_t)atm * ncol * nint, fin_s, nullptr, ncol, ctile * NC, frow);
        }
        cp_async_commit();
    };

    if (blockIdx.x < total) issue(blockIdx.x);
    unsigned phase = 0;
    const double toa_scale = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI;
    for (unsigned ct = blockIdx.x; ct < total; ct += gridDim.x) {
        const unsigned atm = atm_of(ct, nct, s);
        const int ctile = (int)(ct - atm * nct);
        const int col = (ctile * WARPS + warp) * CPW + cw;
        const bool live = col < ncol;  // uniform per segment; dead segments still shuffle
        const int colc = live ? col : ncol - 1;
        // ---- lift the staged tile into registers
        // (batch) the atmosphere's converged flag, requested before the waits so that its latency hides behind them
        const int done_flag = s.done != nullptr ? __ldg(s.done + atm) : 0;
        cp_async_wait_all();
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();  // everybody's flux copies have landed; the previous tile's cooperative stores are done
        double a[CH], b[CH], sd[CH], su[NOBEAM ? 1 : CH], Fu_reg[CH], Fd_reg[CH], cc[CH];
        const int mec = act ? me : cw * rs;  // idle lanes mirror chunk 0 of their column; nothing of theirs is consumed
        const double* __restrict__ bb = one_bin(ctile) ? bbuf0 + (act ? sl : 0) : bbuf + mec;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const double B = bb[k * rl];
            a[k] = pbuf[(k * NR + 0) * rl + mec];
            b[k] = pbuf[(k * NR + 1) * rl + mec];
            const double k1 = pbuf[(k * NR + 2) * rl + mec];
            Fu_reg[k] = fin[k * frow + fme];
            Fd_reg[k] = 0.0;
            if (NOBEAM) {
                sd[k] = __dmul_rn(k1, B);
            } else {
                sd[k] = __fma_rn(k1, B, pbuf[(k * NR + 3) * rl + mec]);
                su[k] = __fma_rn(k1, B, pbuf[(k * NR + 4) * rl + mec]);
            }
        }
        const double emis = __dmul_rn(pbuf[CH * NR * rl + cw * 2], cbuf[cw * 4 + 1]);
        const double Fdir0 = pbuf[CH * NR * rl + cw * 2 + 1];
        const double toa = toa_scale * cbuf[cw * 4 + 0];
        const double A_s = cbuf[cw * 4 + 2];
        __syncthreads();  // every thread has read the staging areas: the next tile's copies may overwrite them
        if (ct + gridDim.x < total) issue(ct + gridDim.x);
        const bool skip = done_flag != 0;  // uniform per CTA

        double F_out = 0.0;
        if (!skip) {
            double scA_dn, scA_up;
            {
                double Adn = 1.0, Aup = 1.0;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) Adn = a[k] * Adn;
#pragma unroll
                for (int k = 0; k < CH; k++) Aup = a[k] * Aup;
                scan_setup<LPC>(Adn, Aup, sl, nch, lane, mult, scA_dn, scA_up);
            }
            const bool top = sl >= nch - 1, bottom = sl == 0;
            auto passes = [&](auto exact_tag) -> unsigned {
                constexpr bool EXACT = decltype(exact_tag)::value;
                unsigned acc = 0u;
                for (int pass = 0; pass < s.npass; pass++) {
                    // ---------------- downward sweep ----------------
                    double mB = 0.0;
#pragma unroll
                    for (int k = CH - 1; k >= 0; k--) {
                        cc[k] = sd[k] - b[k] * Fu_reg[k];
                        mB = a[k] * mB + cc[k];
                    }
                    mB = scan_dn<LPC>(mB, lane, mult);
                    const double Fbot = scA_dn * toa + mB;                   // flux leaving my chunk (interface lo)
                    double F = __shfl_down_sync(0xffffffffu, Fbot, 1, LPC);  // = flux entering it
                    F = top ? toa : F;
#pragma unroll
                    for (int k = CH - 1; k >= 0; k--) {
                        F = clean<EXACT>(__fma_rn(a[k], F, cc[k]), acc);
                        Fd_reg[k] = F;
                    }
                    // the flux at my top interface as WALKED (and stored) by the lane above
                    double Fd_hi = __shfl_down_sync(0xffffffffu, Fd_reg[0], 1, LPC);
                    Fd_hi = top ? toa : Fd_hi;
                    // ---------------- upward sweep ----------------
                    double fu0 = __fma_rn(A_s, __dadd_rn(Fdir0, Fd_reg[0]), emis);  // surface, valid in lane 0 (K:1469-1474)
                    fu0 = __shfl_sync(0xffffffffu, fu0, 0, LPC);
                    mB = 0.0;
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const double Fd_top = (k + 1 < CH) ? Fd_reg[(k + 1) % CH] : Fd_hi;
                        cc[k] = (NOBEAM ? sd[k] : su[NOBEAM ? 0 : k]) - b[k] * Fd_top;
                        mB = a[k] * mB + cc[k];
                    }
                    mB = scan_up<LPC>(mB, lane, mult);
                    const double Ftop = scA_up * fu0 + mB;          // flux leaving my chunk (interface hi)
                    F = __shfl_up_sync(0xffffffffu, Ftop, 1, LPC);  // = flux entering it (interface lo)
                    F = bottom ? fu0 : F;
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        Fu_reg[k] = F;  // interface lo+k: what the next pass's downward sweep reads
                        F = clean<EXACT>(__fma_rn(a[k], F, cc[k]), acc);
                    }
                    const double Fu_lo = __shfl_up_sync(0xffffffffu, F, 1, LPC);
                    Fu_reg[0] = bottom ? fu0 : Fu_lo;  // as walked by the lane below: bit-identical to what F_up holds there
                    F_out = F;
                }
                return acc;
            };
            const unsigned acc = passes(std::false_type{});
            if (__any_sync(0xffffffffu, (int)acc < 0)) {
                // a flux with its sign bit set appeared somewhere in the tile: redo it with the reference's clean-up, from
                // the previous upward fluxes (still unmodified in HBM)
                const double* __restrict__ Fu = F_up + (size_t)atm * ncol * nint + colc;
#pragma unroll
                for (int k = 0; k < CH; k++) Fu_reg[k] = (act && lo + k < nlay) ? Fu[(size_t)ncol * (lo + k)] : 0.0;
                passes(std::true_type{});
            }
        }
        // ---- the fluxes of the last pass: registers -> staged block -> cooperative full-sector stores
        __syncthreads();  // every warp is done with its scan multipliers (fout_dn aliases them)
        if (act) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                fout_dn[k * frow + fme] = Fd_reg[k];
                fout_up[k * frow + fme] = Fu_reg[k];
            }
        }
        __syncthreads();
        if (!skip) {
            double* __restrict__ gd = F_down + (size_t)atm * ncol * nint;
            double* __restrict__ gu = F_up + (size_t)atm * ncol * nint;
            if (PAIRS) {
                PairMap<CH, NC, THREADS> pmap;
                pmap.init(cp);
                flux_pairs<CH, NC, THREADS, false>(pmap, gd, 0u, fout_dn, nlay, ncol, ctile * NC, frow);
                flux_pairs<CH, NC, THREADS, false>(pmap, gu, 0u, fout_up, nlay, ncol, ctile * NC, frow);
            } else {
                FluxMap<CH, NC, THREADS> fmap;
                fmap.init(nlay, cp);
                flux_block<CH, NC, THREADS, false>(fmap, gd, 0u, fout_dn, ncol, ctile * NC, frow);
                flux_block<CH, NC, THREADS, false>(fmap, gu, 0u, fout_up, ncol, ctile * NC, frow);
            }
            if (live && sl == nch - 1) {  // interface nlayer (TOA)
                gd[(size_t)ncol * nlay + colc] = toa;
                gu[(size_t)ncol * nlay + colc] = F_out;
            }
        }
    }
}

// =================================================================================================
// Non-isothermal layers: two steps per layer (upper half, lower half).  NR = 8 rows per layer
// [a_u, b_u, k1_u, k2_u, a_l, b_l, k1_l, k2_l] (+ [k0d_u, k0u_u, k0d_l, k0u_l] with a beam), two previous fluxes
// (F_up, Fc_up), CH layer Planck values + CH+1 interface Planck values.  One column per warp (LPC = 32).
// =================================================================================================
template <int CH, bool NOBEAM, int WARPS, int MINB, bool PAIRS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_sweep_noniso(double* __restrict__ F_down, double* __restrict__ F_up, double* __restrict__ Fc_down,
               double* __restrict__ Fc_up, const double* __restrict__ planck_lay, const double* __restrict__ planck_int,
               const double* __restrict__ plan, const double* __restrict__ albedo, PlanScalars s) {
    constexpr int NR = NOBEAM ? 8 : 12;
    constexpr int LPC = 32;
    constexpr int NC = WARPS;
    constexpr int NBV = 2 * CH + 1;
    constexpr int THREADS = WARPS * 32;
    using CS = CtaShape<CH, NR, 2, NBV, 5, 1, WARPS>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nint = s.nint, nlay = nint - 1, nch = s.nch, rl = s.rs, cp = s.cp, frow = NC * cp;
    const int ncol = s.nbin * s.ny;
    const unsigned nct = (unsigned)(ncol + NC - 1) / NC;  // CTA tiles per atmosphere
    const unsigned total = nct * (unsigned)s.nbatch;
    const int PB = CS::pb(rl);
    double* wst = smem + (size_t)warp * CS::warp_doubles(rl);
    double* pbuf = wst + 2;
    double* bbuf = pbuf + PB;  // [2*CH+1][rl]: B_layer (k), B_interface (CH + k), k = 0..CH
    double* cbuf = bbuf + NBV * rl;
    double* cta = smem + (size_t)WARPS * CS::warp_doubles(rl);
    volatile double* mult = cta + warp * (2 * 5 * 32);
    double* fout_dn = cta;  // [F_down | Fc_down]; aliases the multipliers
    double* fin = cta + CS::mult_doubles(rl);  // [F_up | Fc_up] of the previous solve
    double* fout_up = fin + CS::flux_doubles(rl);
    const int fone = CH * frow;  // one staged flux array
    const unsigned bar = smem_u32(wst), pbuf_s = smem_u32(pbuf), bbuf_s = smem_u32(bbuf), cbuf_s = smem_u32(cbuf),
                   fin_s = smem_u32(fin);
    const int sl = lane;
    const bool act = sl < nch;
    const int lo = sl * CH;
    for (int k = lane; k < NBV * rl + 8; k += 32) bbuf[k] = 0.0;
    for (int k = threadIdx.x; k < CS::flux_doubles(rl); k += THREADS) fin[k] = 0.0;
    // PAIRS (the host checks: even column count, 16-byte aligned arrays): the fluxes move as 16-byte pieces
    const int fme = PAIRS ? PairMap<CH, NC, THREADS>::owner(warp, act ? sl : 0, cp)
                          : warp * cp + (act ? sl : 0);  // my element of a staged flux row of the CTA
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    const double* bbuf0 = smem + 2 + PB;  // warp 0's Planck block
    const unsigned bbuf0_s = smem_u32(bbuf0);
    auto one_bin = [&](int ctile) { return bin_of(ctile * NC, s) == bin_of(min(ctile * NC + NC - 1, ncol - 1), s); };

    auto issue = [&](unsigned ct) {
        const unsigned atm = atm_of(ct, nct, s);
        const int ctile = (int)(ct - atm * nct);
        const int col = min(ctile * NC + warp, ncol - 1);  // warps beyond the last column mirror it
        const int x = bin_of(col, s);
        if (lane == 0 && !ABL(8)) {
            fence_proxy_async();
            mbar_expect_tx(bar, (unsigned)PB * 8u);
            bulk_g2s(pbuf_s, plan + ((size_t)atm * ncol + (ABL(256) ? (col & 255) : col)) * PB, (unsigned)PB * 8u, bar);
        }
        const double* __restrict__ BL = planck_lay + (size_t)atm * (nlay + 2) * s.nbin + (size_t)x * (nlay + 2);
        const double* __restrict__ BI = planck_int + (size_t)atm * s.nbin * nint + (size_t)x * nint;
        // The Planck values belong to the bin: when the CTA's columns all lie in one bin (always, if the tile width
        // divides ny) they are staged ONCE, in warp 0's block, each warp fetching every WARPS-th row -- a quarter of the
        // requests (each touches a sector per lane) and of the L1 lines of the per-warp form.
        const bool shx = one_bin(ctile);
        const unsigned bdst = shx ? bbuf0_s : bbuf_s;
        if (act && !ABL(16)) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const int i = lo + k;
                if (i < nlay && (!shx || k % WARPS == warp)) {
                    cp_async8(bdst + (unsigned)(k * rl + sl) * 8u, BL + i);
                    cp_async8(bdst + (unsigned)((CH + k) * rl + sl) * 8u, BI + i);
                    if (k == CH - 1 || i == nlay - 1) cp_async8(bdst + (unsigned)((CH + k + 1) * rl + sl) * 8u, BI + i + 1);
                }
            }
        }
        if (sl < 3) cp_async8(cbuf_s + (unsigned)sl * 8u, sl == 0 ? BL + nlay : (sl == 1 ? BL + nlay + 1 : albedo + x));
        const size_t ao = (size_t)atm * ncol * nint;
        if (ABL(4)) {
        } else if (PAIRS) {
            PairMap<CH, NC, THREADS> pmap;
            pmap.init(cp);
            flux_pairs<CH, NC, THREADS, true>(pmap, F_up + ao, fin_s, nullptr, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
            flux_pairs<CH, NC, THREADS, true>(pmap, Fc_up + ao, fin_s + (unsigned)fone * 8u, nullptr, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
        } else {
            FluxMap<CH, NC, THREADS> fmap;
            fmap.init(nlay, cp);
            flux_block<CH, NC, THREADS, true>(fmap, F_up + ao, fin_s, nullptr, ncol, ctile * NC, frow);
            flux_block<CH, NC, THREADS, true>(fmap, Fc_up + ao, fin_s + (unsigned)fone * 8u, nullptr, ncol, ctile * NC, frow);
        }
        cp_async_commit();
    };

    if (blockIdx.x < total) issue(blockIdx.x);
#ifdef HELIOS_ABLATE
    if (s.skew_ns > 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const unsigned long long wait = (unsigned long long)(blockIdx.x / s.nsm) * s.skew_ns;
        do {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < wait);
    }
#endif
    unsigned phase = 0;
    const double toa_scale = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI;
    for (unsigned ct = blockIdx.x; ct < total; ct += gridDim.x) {
        const unsigned atm = atm_of(ct, nct, s);
        const int ctile = (int)(ct - atm * nct);
        const int col = ctile * NC + warp;
        const bool live = col < ncol;  // uniform per warp
        const int colc = live ? col : ncol - 1;
        // (batch) the atmosphere's converged flag, requested before the waits so that its latency hides behind them
        const int done_flag = s.done != nullptr ? __ldg(s.done + atm) : 0;
        cp_async_wait_all();
        if (!ABL(8)) mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();  // everybody's flux copies have landed; the previous tile's cooperative stores are done
        // step constants: [0] = upper half, [1] = lower half of the lane's k-th layer
        double a[2][CH], b[2][CH], sd[2][CH], su[2][CH], Fu_reg[CH], Fcu_reg[CH], Fd_reg[CH], Fcd_reg[CH], cc[2][CH];
        const int mec = act ? sl : 0;
        const double* __restrict__ bb = one_bin(ctile) ? bbuf0 : bbuf;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const double Blay = bb[k * rl + mec], Bint_lo = bb[(CH + k) * rl + mec];
            const double Bint_hi = bb[(CH + k + 1) * rl + mec];  // the interface above my k-th layer
            const double* __restrict__ p = pbuf + (size_t)(k * NR) * rl + mec;
#ifdef HELIOS_ABLATE
            if (ABL(128)) {  // no shared-memory reads of the constants
                a[0][k] = a[1][k] = s.f_factor * 0.9; b[0][k] = b[1][k] = s.f_factor * 0.01;
                sd[0][k] = sd[1][k] = su[0][k] = su[1][k] = s.Rstar * 1e-12 + k;
                Fu_reg[k] = Fcu_reg[k] = s.a * 1e-13; Fd_reg[k] = Fcd_reg[k] = 0.0;
                continue;
            }
#endif
            a[0][k] = p[0];
            b[0][k] = p[rl];
            const double k1u = p[2 * rl], k2u = p[3 * rl];
            a[1][k] = p[4 * rl];
            b[1][k] = p[5 * rl];
            const double k1l = p[6 * rl], k2l = p[7 * rl];
            double k0d_u = 0.0, k0u_u = 0.0, k0d_l = 0.0, k0u_l = 0.0;
            if (!NOBEAM) {
                k0d_u = p[8 * rl];
                k0u_u = p[9 * rl];
                k0d_l = p[10 * rl];
                k0u_l = p[11 * rl];
            }
            sd[0][k] = __fma_rn(k1u, Blay, __fma_rn(k2u, Bint_hi, k0d_u));
            su[0][k] = __fma_rn(k2u, Blay, __fma_rn(k1u, Bint_hi, k0u_u));
            sd[1][k] = __fma_rn(k1l, Blay, __fma_rn(k2l, Bint_lo, k0d_l));
            su[1][k] = __fma_rn(k2l, Blay, __fma_rn(k1l, Bint_lo, k0u_l));
            Fu_reg[k] = fin[k * frow + fme];
            Fcu_reg[k] = fin[fone + k * frow + fme];
            Fd_reg[k] = Fcd_reg[k] = 0.0;
        }
        const double emis = __dmul_rn(pbuf[CH * NR * rl], cbuf[1]);
        const double Fdir0 = pbuf[CH * NR * rl + 1];
        const double toa = toa_scale * cbuf[0];
        const double A_s = cbuf[2];
        __syncthreads();  // every thread has read the staging areas: the next tile's copies may overwrite them
        if (ct + gridDim.x < total) issue(ct + gridDim.x);
        const bool skip = done_flag != 0;  // uniform per CTA

        double F_out = 0.0;
        if (!skip) {
            double scA_dn, scA_up;
            {
                double Adn = 1.0, Aup = 1.0;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) Adn = a[1][k] * (a[0][k] * Adn);
#pragma unroll
                for (int k = 0; k < CH; k++) Aup = a[0][k] * (a[1][k] * Aup);
                if (!ABL(64)) scan_setup<LPC>(Adn, Aup, sl, nch, lane, mult, scA_dn, scA_up);
                else scA_dn = Adn, scA_up = Aup;
            }
            const bool top = sl >= nch - 1, bottom = sl == 0;
            auto passes = [&](auto exact_tag) -> unsigned {
                constexpr bool EXACT = decltype(exact_tag)::value;
                unsigned acc = 0u;
                for (int pass = 0; pass < (ABL(1) ? 0 : s.npass); pass++) {
                    // ---------------- downward sweep (per layer: upper half, then lower half) ----------------
                    double mB = 0.0;
#pragma unroll
                    for (int k = CH - 1; k >= 0; k--) {
                        cc[0][k] = sd[0][k] - b[0][k] * Fcu_reg[k];
                        mB = a[0][k] * mB + cc[0][k];
                        cc[1][k] = sd[1][k] - b[1][k] * Fu_reg[k];
                        mB = a[1][k] * mB + cc[1][k];
                    }
                    mB = scan_dn<LPC>(mB, lane, mult);
                    const double Fbot = scA_dn * toa + mB;
                    double F = __shfl_down_sync(0xffffffffu, Fbot, 1, LPC);
                    F = top ? toa : F;
#pragma unroll
                    for (int k = CH - 1; k >= 0; k--) {
                        F = clean<EXACT>(__fma_rn(a[0][k], F, cc[0][k]), acc);
                        Fcd_reg[k] = F;
                        F = clean<EXACT>(__fma_rn(a[1][k], F, cc[1][k]), acc);
                        Fd_reg[k] = F;
                    }
                    double Fd_hi = __shfl_down_sync(0xffffffffu, Fd_reg[0], 1, LPC);
                    Fd_hi = top ? toa : Fd_hi;
                    // ---------------- upward sweep (per layer: lower half, then upper half) ----------------
                    double fu0 = __fma_rn(A_s, __dadd_rn(Fdir0, Fd_reg[0]), emis);  // surface, valid in lane 0 (K:1469-1474)
                    fu0 = __shfl_sync(0xffffffffu, fu0, 0, LPC);
                    mB = 0.0;
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const double Fd_top = (k + 1 < CH) ? Fd_reg[(k + 1) % CH] : Fd_hi;
                        cc[1][k] = su[1][k] - b[1][k] * Fcd_reg[k];
                        mB = a[1][k] * mB + cc[1][k];
                        cc[0][k] = su[0][k] - b[0][k] * Fd_top;
                        mB = a[0][k] * mB + cc[0][k];
                    }
                    mB = scan_up<LPC>(mB, lane, mult);
                    const double Ftop = scA_up * fu0 + mB;
                    F = __shfl_up_sync(0xffffffffu, Ftop, 1, LPC);
                    F = bottom ? fu0 : F;
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        Fu_reg[k] = F;
                        // no tiny-value clean-up on Fc_up: the reference applies it to index i, not i-1 (K:1763)
                        F = __fma_rn(a[1][k], F, cc[1][k]);
                        Fcu_reg[k] = F;
                        F = clean<EXACT>(__fma_rn(a[0][k], F, cc[0][k]), acc);
                    }
                    const double Fu_lo = __shfl_up_sync(0xffffffffu, F, 1, LPC);
                    Fu_reg[0] = bottom ? fu0 : Fu_lo;
                    F_out = F;
                }
                return acc;
            };
            const unsigned acc = passes(std::false_type{});
            if (__any_sync(0xffffffffu, (int)acc < 0)) {
                const size_t wgo = (size_t)atm * ncol * nint + colc;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const bool in = act && lo + k < nlay;
                    Fu_reg[k] = in ? F_up[wgo + (size_t)ncol * (lo + k)] : 0.0;
                    Fcu_reg[k] = in ? Fc_up[wgo + (size_t)ncol * (lo + k)] : 0.0;
                }
                passes(std::true_type{});
            }
        }
        // ---- the fluxes of the last pass: registers -> staged block -> cooperative full-sector stores
        __syncthreads();  // every warp is done with its scan multipliers (fout_dn aliases them)
        if (act && !ABL(32)) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const int o = k * frow + fme;
                fout_dn[o] = Fd_reg[k];
                fout_dn[fone + o] = Fcd_reg[k];
                fout_up[o] = Fu_reg[k];
                fout_up[fone + o] = Fcu_reg[k];
            }
        }
        __syncthreads();
        if (ABL(32)) {  // keep the results alive without the staging round trip
            double v = F_out;
#pragma unroll
            for (int k = 0; k < CH; k++) v += Fd_reg[k] + Fcd_reg[k] + Fu_reg[k] + Fcu_reg[k];
            if (v == 1.2345e-300) F_down[0] = v;
        }
        if (!skip && !ABL(2) && !ABL(32)) {
            const size_t ao = (size_t)atm * ncol * nint;
            if (PAIRS) {
                PairMap<CH, NC, THREADS> pmap;
                pmap.init(cp);
                flux_pairs<CH, NC, THREADS, false>(pmap, F_down + ao, 0u, fout_dn, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
                flux_pairs<CH, NC, THREADS, false>(pmap, Fc_down + ao, 0u, fout_dn + fone, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
                flux_pairs<CH, NC, THREADS, false>(pmap, F_up + ao, 0u, fout_up, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
                flux_pairs<CH, NC, THREADS, false>(pmap, Fc_up + ao, 0u, fout_up + fone, nlay, ncol, (ABL(512) ? (ctile & 63) : ctile) * NC, frow);
            } else {
                FluxMap<CH, NC, THREADS> fmap;
                fmap.init(nlay, cp);
                flux_block<CH, NC, THREADS, false>(fmap, F_down + ao, 0u, fout_dn, ncol, ctile * NC, frow);
                flux_block<CH, NC, THREADS, false>(fmap, Fc_down + ao, 0u, fout_dn + fone, ncol, ctile * NC, frow);
                flux_block<CH, NC, THREADS, false>(fmap, F_up + ao, 0u, fout_up, ncol, ctile * NC, frow);
                flux_block<CH, NC, THREADS, false>(fmap, Fc_up + ao, 0u, fout_up + fone, ncol, ctile * NC, frow);
            }
            if (live && sl == nch - 1) {  // interface nlayer (TOA)
                F_down[ao + (size_t)ncol * nlay + colc] = toa;
                F_up[ao + (size_t)ncol * nlay + colc] = F_out;
            }
        }
    }
}

// =================================================================================================
// Plan builders.  A CTA owns TC consecutive columns: its threads read the coefficient arrays with lanes along the
// columns (coalesced row segments), form the constants, place them at their position of the tile blocks in shared
// memory (the transposition), and the finished blocks -- contiguous in the plan -- go out as full 16-byte vector stores.
// =================================================================================================
struct BuildScalars {
    double g_0, mu_star, epsi, delta_tau_limit, i2s_transition;
    int nint, nbin, ny, clouds, scat_corr, no_beam, nch, rs, nbatch;
};

__device__ __forceinline__ void beam_pair(double Fa, double Fb, double neg_mu, double M, double N, double P, double G_pl,
                                          double G_min, double m1d, double m2d, double& Dd, double& Du) {
    Dd = 0.0;
    Du = 0.0;
    if (Fa != 0.0 || Fb != 0.0) {  // min(0, +-0) contributes nothing; also saves the four fp64 divisions
        Dd = beam_source(Fa, Fb, neg_mu, M, G_min, N, G_pl, m1d, m2d);
        Du = beam_source(Fb, Fa, neg_mu, N, G_min, M, G_pl, P, G_pl);
    }
}

template <int CH, int LPC, int TC>
__global__ void __launch_bounds__(256)
k_plan_build_iso(double* __restrict__ plan, const double* __restrict__ F_dir, const double* __restrict__ w_0,
                 const double* __restrict__ Mt, const double* __restrict__ Nt, const double* __restrict__ Pt,
                 const double* __restrict__ Gp, const double* __restrict__ Gm, const double* __restrict__ albedo,
                 const double* __restrict__ g0_lay, BuildScalars s) {
    constexpr int CPW = 32 / LPC;
    constexpr int TPB = TC / CPW;  // warp tiles per CTA
    extern __shared__ __align__(16) double sm[];
    const int NR = s.no_beam ? 3 : 5;
    const int nint = s.nint, nlay = nint - 1, ncol = s.nbin * s.ny, rs = s.rs, rl = rs * CPW;
    const int ntw = (ncol + CPW - 1) / CPW;
    const int PB = CH * NR * rl + 4;
    const int nblk = (ntw + TPB - 1) / TPB;  // CTAs' worth of tiles per atmosphere
    const double neg_mu = -s.mu_star;
    const int c = threadIdx.x % TC, r0 = threadIdx.x / TC;
    constexpr int RSTEP = 256 / TC;
    for (int blk = blockIdx.x; blk < nblk * s.nbatch; blk += gridDim.x) {
        const int atm = blk / nblk;
        const int t0 = (blk - atm * nblk) * TPB;  // first warp tile of this CTA
        const int ntl = min(TPB, ntw - t0);
        // identity steps (a = 1, everything else 0) in the slots no layer maps to: chunk * CH + k >= nlay.  Every other slot
        // is written below, and so are the surface constants of every column.
        {
            const int first = nlay;                 // slots are numbered chunk * CH + k
            const int nfree = rs * CH - first;      // per column
            for (int e = threadIdx.x; e < ntl * CPW * nfree; e += 256) {
                const int cc = e / nfree, slot = first + e % nfree;
                double* __restrict__ p = sm + (size_t)(cc / CPW) * PB + (size_t)((slot % CH) * NR) * rl + (cc % CPW) * rs + slot / CH;
                p[0] = 1.0;
                for (int j = 1; j < NR; j++) p[j * rl] = 0.0;
            }
            // (the trailing doubles behind the 2 * CPW surface constants are padding)
            if constexpr (CPW == 1) {
                for (int e = threadIdx.x; e < ntl * 2; e += 256) sm[(size_t)(e / 2) * PB + CH * NR * rl + 2 + e % 2] = 0.0;
            }
        }
        const int col = min(t0 * CPW + c, ncol - 1);
        const int x = col / s.ny;
        const int tl = c / CPW, cw = c % CPW;  // tile within the CTA, column within the tile
        if (tl < ntl) {
            for (int i = r0; i < nlay; i += RSTEP) {
                const size_t e = (size_t)atm * ncol * nint + col + (size_t)ncol * i;
                const double w0 = w_0[e], M = Mt[e], N = Nt[e], P = Pt[e];
                const double g0 = s.clouds ? g0_lay[(size_t)atm * s.nbin * nlay + (size_t)x + (size_t)s.nbin * i] : s.g_0;
                const double E = s.scat_corr ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
                double Dd = 0.0, Du = 0.0, Fdir_i = -0.0;
                if (!s.no_beam) {
                    Fdir_i = F_dir[e];
                    const double gm = Gm[e];
                    beam_pair(Fdir_i, F_dir[e + ncol], neg_mu, M, N, P, Gp[e], gm, P, gm, Dd, Du);
                }
                const double invM = 1.0 / M;
                const double f = invM * source_factor(s.epsi, w0, E);
                double* __restrict__ p = sm + (size_t)tl * PB + (size_t)((i % CH) * NR) * rl + cw * rs + i / CH;
                p[0] = invM * P;
                p[rl] = invM * N;
                p[2 * rl] = f * ((M + N) - P);
                if (!s.no_beam) {
                    p[3 * rl] = invM * Dd;
                    p[4 * rl] = invM * Du;
                }
                if (i == 0) {  // surface constants (K:1472: w0 and E of layer 0)
                    const double A_s = albedo[x];
                    double* __restrict__ ex = sm + (size_t)tl * PB + CH * NR * rl + cw * 2;
                    ex[0] = __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, A_s), 3.141592653589793), __dsub_rn(1.0, w0)),
                                      __dsub_rn(E, w0));
                    ex[1] = Fdir_i;
                }
            }
        }
        __syncthreads();
        double2* __restrict__ dst = reinterpret_cast<double2*>(plan + ((size_t)atm * ntw + t0) * PB);
        const double2* __restrict__ src = reinterpret_cast<const double2*>(sm);
        for (int k = threadIdx.x; k < ntl * PB / 2; k += 256) dst[k] = src[k];
        __syncthreads();
    }
}

struct NonisoCoefP {
    const double *w0_u, *w0_l, *dtau_u, *dtau_l, *dtc_u, *dtc_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u, *Gp_l, *Gm_u,
        *Gm_l;
};

struct Half {
    double a, b, k0d, k0u, k1, k2;
};

// one half-layer (K:1640-1691 down, K:1744-1795 up).  `upper`: B1 of the downward form is the layer value.
__device__ __forceinline__ Half plan_half(double w0, double M, double N, double P, double Gp, double Gm, double dt, double g0,
                                          double Fa_d, double Fb_d, double m1d, double m2d, bool upper, double neg_mu,
                                          const BuildScalars& s, double& E_out) {
    const double E = s.scat_corr ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
    E_out = E;
    const double invM = 1.0 / M;
    const double fac = source_factor(s.epsi, w0, E);
    double Dd, Du;
    beam_pair(Fa_d, Fb_d, neg_mu, M, N, P, Gp, Gm, m1d, m2d, Dd, Du);
    double lay_d, int_d;  // weights of B_layer / B_interface in the downward Planck term; upward: swapped (see header)
    if (dt < s.delta_tau_limit) {
        lay_d = int_d = 0.5 * ((M + N) - P);
    } else {
        const double pre = gradient_factor(s.epsi, w0, g0, E);
        const double cd = (N + (P - M)) * pre / dt;
        if (upper) {  // down: B1 = layer, B2 = interface above
            lay_d = (M + N) + cd;
            int_d = -P - cd;
        } else {      // down: B1 = interface below, B2 = layer
            lay_d = -P - cd;
            int_d = (M + N) + cd;
        }
    }
    const double f = invM * fac;
    return Half{invM * P, invM * N, invM * Dd, invM * Du, f * lay_d, f * int_d};
}

template <int CH, int TC>
__global__ void __launch_bounds__(256)
k_plan_build_noniso(double* __restrict__ plan, const double* __restrict__ F_dir, const double* __restrict__ Fc_dir,
                    NonisoCoefP cfg, const double* __restrict__ albedo, const double* __restrict__ g0_lay,
                    const double* __restrict__ g0_int, BuildScalars s) {
    extern __shared__ __align__(16) double sm[];
    const int NR = s.no_beam ? 8 : 12;
    const int nint = s.nint, nlay = nint - 1, ncol = s.nbin * s.ny, rl = s.rs;
    const int PB = CH * NR * rl + 4;
    const int nblk = (ncol + TC - 1) / TC;
    const double neg_mu = -s.mu_star;
    const int c = threadIdx.x % TC, r0 = threadIdx.x / TC;
    constexpr int RSTEP = 256 / TC;
    for (int blk = blockIdx.x; blk < nblk * s.nbatch; blk += gridDim.x) {
        const int atm = blk / nblk;
        const int t0 = (blk - atm * nblk) * TC;
        const int ntl = min(TC, ncol - t0);
        {   // identity steps (a_u = a_l = 1, everything else 0) in the slots no layer maps to; trailing doubles cleared
            const int first = nlay, nfree = rl * CH - first;
            for (int e = threadIdx.x; e < ntl * nfree; e += 256) {
                const int cc = e / nfree, slot = first + e % nfree;
                double* __restrict__ p = sm + (size_t)cc * PB + (size_t)((slot % CH) * NR) * rl + slot / CH;
                for (int j = 0; j < NR; j++) p[j * rl] = (j == 0 || j == 4) ? 1.0 : 0.0;
            }
            for (int e = threadIdx.x; e < ntl * 2; e += 256) sm[(size_t)(e / 2) * PB + CH * NR * rl + 2 + e % 2] = 0.0;  // padding
        }
        const int col = min(t0 + c, ncol - 1);
        const int x = col / s.ny;
        if (c < ntl) {
            for (int i = r0; i < nlay; i += RSTEP) {
                const size_t e = (size_t)atm * ncol * nint + col + (size_t)ncol * i;
                const size_t bl = (size_t)atm * s.nbin * nlay + (size_t)x + (size_t)s.nbin * i;
                const size_t bi = (size_t)atm * s.nbin * nint + (size_t)x + (size_t)s.nbin * i;
                double g0_up = s.g_0, g0_low = s.g_0;
                if (s.clouds) {
                    const double gl = g0_lay[bl];
                    g0_up = (gl + g0_int[bi + s.nbin]) / 2.0;
                    g0_low = (g0_int[bi] + gl) / 2.0;
                }
                double Fdir_i = -0.0, Fdir_ip1 = -0.0, Fcdir = -0.0;
                double Gp_u = 0.0, Gm_u = 0.0, Gp_l = 0.0, Gm_l = 0.0;
                if (!s.no_beam) {
                    Fdir_i = F_dir[e];
                    Fdir_ip1 = F_dir[e + ncol];
                    Fcdir = Fc_dir[e];
                    Gp_u = cfg.Gp_u[e];
                    Gm_u = cfg.Gm_u[e];
                    Gp_l = cfg.Gp_l[e];
                    Gm_l = cfg.Gm_l[e];
                }
                double E_u, E_l;
                const double P_u = cfg.P_u[e];
                const Half u = plan_half(cfg.w0_u[e], cfg.M_u[e], cfg.N_u[e], P_u, Gp_u, Gm_u, cfg.dtau_u[e] + cfg.dtc_u[bl],
                                         g0_up, Fcdir, Fdir_ip1, Gm_u, P_u, true, neg_mu, s, E_u);
                const double w0_l = cfg.w0_l[e], P_l = cfg.P_l[e];
                const Half l = plan_half(w0_l, cfg.M_l[e], cfg.N_l[e], P_l, Gp_l, Gm_l, cfg.dtau_l[e] + cfg.dtc_l[bl], g0_low,
                                         Fdir_i, Fcdir, P_l, Gm_l, false, neg_mu, s, E_l);
                double* __restrict__ p = sm + (size_t)c * PB + (size_t)((i % CH) * NR) * rl + i / CH;
                p[0] = u.a;
                p[rl] = u.b;
                p[2 * rl] = u.k1;
                p[3 * rl] = u.k2;
                p[4 * rl] = l.a;
                p[5 * rl] = l.b;
                p[6 * rl] = l.k1;
                p[7 * rl] = l.k2;
                if (!s.no_beam) {
                    p[8 * rl] = u.k0d;
                    p[9 * rl] = u.k0u;
                    p[10 * rl] = l.k0d;
                    p[11 * rl] = l.k0u;
                }
                if (i == 0) {  // surface constants (K:1704: w0 and E of layer 0's lower half)
                    const double A_s = albedo[x];
                    double* __restrict__ ex = sm + (size_t)c * PB + CH * NR * rl;
                    ex[0] = __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(1.0, A_s), 3.141592653589793), __dsub_rn(1.0, w0_l)),
                                      __dsub_rn(E_l, w0_l));
                    ex[1] = Fdir_i;
                }
            }
        }
        __syncthreads();
        double2* __restrict__ dst = reinterpret_cast<double2*>(plan + ((size_t)atm * ncol + t0) * PB);
        const double2* __restrict__ src = reinterpret_cast<const double2*>(sm);
        for (int k = threadIdx.x; k < ntl * PB / 2; k += 256) dst[k] = src[k];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host side ----------------------
inline int even_up(int v) { return (v + 1) & ~1; }

// tile shapes of the isothermal plan: CH layers per lane
#define ISO_SHAPES(X)                     \
    if (nlay <= 16) { X(1, 16); }         \
    else if (nlay <= 32) { X(2, 16); }    \
    else if (nlay <= 48) { X(3, 16); }    \
    else if (nlay <= 80) { X(5, 16); }    \
    else if (nlay <= 112) { X(7, 16); }   \
    else if (nlay <= 128) { X(8, 16); }   \
    else if (nlay <= 192) { X(6, 32); }   \
    else if (nlay <= 256) { X(8, 32); }

#define NONISO_SHAPES(X)               \
    if (nlay <= 32) { X(1); }          \
    else if (nlay <= 64) { X(2); }     \
    else if (nlay <= 96) { X(3); }     \
    else if (nlay <= 128) { X(4); }

// columns per CTA of the isothermal plan build: 32 (256-byte row segments) for two-column warp tiles, 8 where a warp
// tile is one column of up to 256 layers (the CTA's tile blocks have to fit into shared memory)
constexpr int iso_tc(int lpc) { return lpc == 16 ? 32 : 8; }
constexpr int NONISO_TC = 8;  // non-isothermal: a column's block is 4x larger, 8 columns fill the shared memory

struct IsoGeom {
    int CH, LPC, nch, rs, cpw, ntw;
    size_t pb;
};
bool iso_geom(int nlay, int ncol, bool nobeam, IsoGeom& g) {
    g.CH = 0;
#define X(CH_, LPC_) { g.CH = CH_; g.LPC = LPC_; }
    ISO_SHAPES(X)
#undef X
    if (g.CH == 0) return false;
    g.nch = (nlay + g.CH - 1) / g.CH;
    g.rs = even_up(g.nch);
    g.cpw = 32 / g.LPC;
    g.ntw = (ncol + g.cpw - 1) / g.cpw;
    g.pb = (size_t)g.CH * (nobeam ? 3 : 5) * g.rs * g.cpw + 4;
    return true;
}
struct NonisoGeom {
    int CH, nch, rs;
    size_t pb;
};
bool noniso_geom(int nlay, bool nobeam, NonisoGeom& g) {
    g.CH = 0;
#define X(CH_) { g.CH = CH_; }
    NONISO_SHAPES(X)
#undef X
    if (g.CH == 0) return false;
    g.nch = (nlay + g.CH - 1) / g.CH;
    g.rs = even_up(g.nch);
    g.pb = (size_t)g.CH * (nobeam ? 8 : 12) * g.rs + 4;
    return true;
}

// floor(2^32 / d) + 1 if n -> n / d by multiply-high is exact for every n < total (n * d < 2^32 suffices), else 0
static unsigned tile_magic(unsigned long long total, unsigned d) {
    if (d <= 1u || total * d >= (1ull << 32)) return 0u;
    return (unsigned)((1ull << 32) / d + 1ull);
}

// persistent grid: as many CTAs as fit per SM (registers, shared memory), capped by the work.  The occupancy query is
// cached per (kernel, shared-memory size): launches may sit inside a CUDA-graph capture.
template <typename K>
int resident_grid(helios_ctx* ctx, K kern, int threads, size_t smem, long long work_ctas) {
    static std::mutex mu;
    static std::unordered_map<size_t, int> cache;
    const size_t key = reinterpret_cast<size_t>(kern) ^ (smem * 0x9E3779B97F4A7C15ull) ^ ((size_t)ctx->device << 56);
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) per_sm = it->second;
    }
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        std::lock_guard<std::mutex> lk(mu);
        cache[key] = per_sm;
    }
    const long long cap = (long long)ctx->num_sms * per_sm;
    return (int)(work_ctas < cap ? work_ctas : cap);
}

template <int CH, int LPC, bool NOBEAM>
int launch_sweep_iso(helios_ctx* ctx, double* F_down, double* F_up, const double* planck_lay, const double* plan,
                     const double* albedo, PlanScalars s, const IsoGeom& g) {
    constexpr int WARPS = 4;
    constexpr int MINB = 4;
    constexpr int NC = WARPS * (32 / LPC);
    // 16-byte pieces of the flux rows (two columns of one interface) need an even column count and aligned arrays
    const bool pairs = ((s.nbin * s.ny) & 1) == 0 && ((reinterpret_cast<size_t>(F_up) | reinterpret_cast<size_t>(F_down)) & 15) == 0;
    auto kern = pairs ? k_sweep_iso<CH, LPC, NOBEAM, WARPS, MINB, true> : k_sweep_iso<CH, LPC, NOBEAM, WARPS, MINB, false>;
    using CS = CtaShape<CH, NOBEAM ? 3 : 5, 1, CH, Log2<LPC>::v, 32 / LPC, WARPS>;
    const size_t smem = (size_t)CS::cta_doubles(g.rs) * sizeof(double);
    HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long total = (long long)((s.nbin * s.ny + NC - 1) / NC) * s.nbatch;
    const int grid = resident_grid(ctx, kern, WARPS * 32, smem, total);
    s.cp = CS::cpitch(g.rs);
    if ((unsigned long long)s.nbin * s.ny * s.ny >= (1ull << 32)) {
        helios_set_error("planned sweep: %d x %d columns exceed the range of the column -> bin multiply", s.nbin, s.ny);
        return HELIOS_ERR_ARG;
    }
    s.ny_magic = (unsigned)((1ull << 32) / (unsigned)s.ny + 1ull);
    s.nct_magic = tile_magic((unsigned long long)total, (unsigned)((s.nbin * s.ny + NC - 1) / NC));
    kern<<<grid, WARPS * 32, smem, ctx->stream>>>(F_down, F_up, planck_lay, plan, albedo, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

template <int CH, bool NOBEAM>
int launch_sweep_noniso(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                        const double* planck_lay, const double* planck_int, const double* plan, const double* albedo,
                        PlanScalars s, const NonisoGeom& g, int ncol) {
    // All step constants live in registers (~156 per lane at CH = 4).  The register file is split four ways (16 K
    // registers per scheduler), so the occupancy points are 12 warps per SM at <= 168 registers and 16 at <= 128; 128
    // spills inside the passes and was measured slower.  CTAs of 4 warps (4 columns = one 32-byte sector per flux row).
    constexpr int WARPS = 4;
    constexpr int MINB = (CH >= 4) ? 3 : 4;
    const bool pairs = (ncol & 1) == 0 && ((reinterpret_cast<size_t>(F_up) | reinterpret_cast<size_t>(Fc_up) |
                                            reinterpret_cast<size_t>(F_down) | reinterpret_cast<size_t>(Fc_down)) & 15) == 0;
    auto kern = pairs ? k_sweep_noniso<CH, NOBEAM, WARPS, MINB, true> : k_sweep_noniso<CH, NOBEAM, WARPS, MINB, false>;
    using CS = CtaShape<CH, NOBEAM ? 8 : 12, 2, 2 * CH + 1, 5, 1, WARPS>;
    const size_t smem = (size_t)CS::cta_doubles(g.rs) * sizeof(double);
    HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long total = (long long)((ncol + WARPS - 1) / WARPS) * s.nbatch;
    int grid = resident_grid(ctx, kern, WARPS * 32, smem, total);
    s.cp = CS::cpitch(g.rs);
    if ((unsigned long long)s.nbin * s.ny * s.ny >= (1ull << 32)) {
        helios_set_error("planned sweep: %d x %d columns exceed the range of the column -> bin multiply", s.nbin, s.ny);
        return HELIOS_ERR_ARG;
    }
    s.ny_magic = (unsigned)((1ull << 32) / (unsigned)s.ny + 1ull);
    s.nct_magic = tile_magic((unsigned long long)total, (unsigned)((ncol + WARPS - 1) / WARPS));
#ifdef HELIOS_ABLATE
    if (const char* e = getenv("HELIOS_SWEEP_ABLATE")) s.ablate = atoi(e);
    if (const char* e = getenv("HELIOS_SWEEP_SKEW_NS")) s.skew_ns = atoi(e);
    s.nsm = ctx->num_sms;
    if (const char* e = getenv("HELIOS_SWEEP_CTAS_PER_SM")) grid = std::min(grid, ctx->num_sms * atoi(e));
#endif
    kern<<<grid, WARPS * 32, smem, ctx->stream>>>(F_down, F_up, Fc_down, Fc_up, planck_lay, planck_int, plan, albedo, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

}  // namespace

// ---------------------------------------------------------------- entry points (called from fband.cu)
size_t plan2_iso_size(int nint, int ncol, int nbatch) {
    IsoGeom g;
    if (!iso_geom(nint - 1, ncol, false, g)) return 0;
    return g.pb * (size_t)g.ntw * nbatch;
}

size_t plan2_noniso_size(int nint, int ncol, int nbatch) {
    NonisoGeom g;
    if (!noniso_geom(nint - 1, false, g)) return 0;
    return g.pb * (size_t)ncol * nbatch;
}

int plan2_iso_build(helios_ctx* ctx, double* plan, const double* F_dir, const double* w_0, const double* M,
                    const double* N, const double* P, const double* Gp, const double* Gm, const double* albedo,
                    const double* g0tot, double g_0, double mu_star, double epsi, int nint, int nbin, int ny,
                    int dir_beam, int clouds, int scat_corr, double i2s) {
    const int nlay = nint - 1, ncol = nbin * ny;
    const bool nobeam = dir_beam == 0 && ctx->zero_beam[0] == F_dir;
    IsoGeom g;
    if (!iso_geom(nlay, ncol, nobeam, g)) return -1;
    BuildScalars s{g_0, mu_star, epsi, 0.0, i2s, nint, nbin, ny, clouds, scat_corr, nobeam ? 1 : 0, g.nch, g.rs,
                   ctx->batch.nbatch};
    const int tpb = iso_tc(g.LPC) / g.cpw;
    const size_t smem = (size_t)tpb * g.pb * sizeof(double);
    const long long nblk = (long long)((g.ntw + tpb - 1) / tpb) * s.nbatch;
#define X(CH_, LPC_)                                                                                              \
    {                                                                                                             \
        auto kern = k_plan_build_iso<CH_, LPC_, iso_tc(LPC_)>;                                                       \
        HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        const int grid = resident_grid(ctx, kern, 256, smem, nblk);                                               \
        kern<<<grid, 256, smem, ctx->stream>>>(plan, F_dir, w_0, M, N, P, Gp, Gm, albedo, g0tot, s);               \
    }
    ISO_SHAPES(X)
#undef X
    HLAUNCHED(ctx);
    helios_plan_record(ctx, plan, g.pb * (size_t)g.ntw * s.nbatch * sizeof(double), 1, nobeam ? 1 : 0, nint, ncol);
    return HELIOS_OK;
}

int plan2_iso_sweep(helios_ctx* ctx, double* F_down, double* F_up, const double* plan, const double* planck_lay,
                    const double* albedo, double Rstar, double a, int nint, int nbin, double f_factor, int ny,
                    int dir_beam, int npass) {
    const int nlay = nint - 1, ncol = nbin * ny;
    const helios_plan_info* info = helios_plan_lookup(ctx, plan, 1, nint, ncol);
    if (info == nullptr) return -2;
    const bool nobeam = info->nobeam != 0;
    IsoGeom g;
    if (!iso_geom(nlay, ncol, nobeam, g)) return -1;
    PlanScalars s{Rstar, a, f_factor, nint, nbin, ny, dir_beam, npass, g.nch, g.rs, ctx->batch.nbatch,
                  ctx->batch.active ? ctx->batch.done : nullptr};
#define X(CH_, LPC_)                                                                                             \
    {                                                                                                            \
        if (nobeam) return launch_sweep_iso<CH_, LPC_, true>(ctx, F_down, F_up, planck_lay, plan, albedo, s, g);  \
        return launch_sweep_iso<CH_, LPC_, false>(ctx, F_down, F_up, planck_lay, plan, albedo, s, g);             \
    }
    ISO_SHAPES(X)
#undef X
    return -1;
}

int plan2_noniso_build(helios_ctx* ctx, double* plan, const double* F_dir, const double* Fc_dir, const double* const* coef,
                       const double* albedo, const double* g0_lay, const double* g0_int, double g_0, double mu_star,
                       double epsi, double delta_tau_limit, int nint, int nbin, int ny, int dir_beam, int clouds,
                       int scat_corr, double i2s) {
    const int nlay = nint - 1, ncol = nbin * ny;
    const bool nobeam = dir_beam == 0 && ctx->zero_beam[0] == F_dir && ctx->zero_beam[1] == Fc_dir;
    NonisoGeom g;
    if (!noniso_geom(nlay, nobeam, g)) return -1;
    BuildScalars s{g_0, mu_star, epsi, delta_tau_limit, i2s, nint, nbin, ny, clouds, scat_corr, nobeam ? 1 : 0, g.nch, g.rs,
                   ctx->batch.nbatch};
    NonisoCoefP c{coef[0], coef[1], coef[2], coef[3], coef[4], coef[5], coef[6], coef[7], coef[8], coef[9], coef[10],
                  coef[11], coef[12], coef[13], coef[14], coef[15]};
    const size_t smem = (size_t)NONISO_TC * g.pb * sizeof(double);
    const long long nblk = (long long)((ncol + NONISO_TC - 1) / NONISO_TC) * s.nbatch;
#define X(CH_)                                                                                               \
    {                                                                                                        \
        auto kern = k_plan_build_noniso<CH_, NONISO_TC>;                                                     \
        HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        const int grid = resident_grid(ctx, kern, 256, smem, nblk);                                          \
        kern<<<grid, 256, smem, ctx->stream>>>(plan, F_dir, Fc_dir, c, albedo, g0_lay, g0_int, s);           \
    }
    NONISO_SHAPES(X)
#undef X
    HLAUNCHED(ctx);
    helios_plan_record(ctx, plan, g.pb * (size_t)ncol * s.nbatch * sizeof(double), 2, nobeam ? 1 : 0, nint, ncol);
    return HELIOS_OK;
}

int plan2_noniso_sweep(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up, const double* plan,
                       const double* planck_lay, const double* planck_int, const double* albedo, double Rstar, double a,
                       int nint, int nbin, double f_factor, int ny, int dir_beam, int npass) {
    const int nlay = nint - 1, ncol = nbin * ny;
    const helios_plan_info* info = helios_plan_lookup(ctx, plan, 2, nint, ncol);
    if (info == nullptr) return -2;
    const bool nobeam = info->nobeam != 0;
    NonisoGeom g;
    if (!noniso_geom(nlay, nobeam, g)) return -1;
    PlanScalars s{Rstar, a, f_factor, nint, nbin, ny, dir_beam, npass, g.nch, g.rs, ctx->batch.nbatch,
                  ctx->batch.active ? ctx->batch.done : nullptr};
#define X(CH_)                                                                                                          \
    {                                                                                                                   \
        if (nobeam) return launch_sweep_noniso<CH_, true>(ctx, F_down, F_up, Fc_down, Fc_up, planck_lay, planck_int, plan, albedo, s, g, ncol); \
        return launch_sweep_noniso<CH_, false>(ctx, F_down, F_up, Fc_down, Fc_up, planck_lay, planck_int, plan, albedo, s, g, ncol);            \
    }
    NONISO_SHAPES(X)
#undef X
    return -1;
}
