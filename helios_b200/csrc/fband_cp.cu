// Layer-parallel two-stream sweeps -- the default flux solver.
//
// Why.  The reference's fband_* (K:1366-1799) walks the ~100 layers of a column serially, twice per
// pass and 3*scat+1 passes per RT iteration: ~800 dependent steps, each a handful of fp64 divides and
// ~20 loads.  With one thread per column that chain is pure latency (an atmosphere of 385 x 20 columns
// occupies 60 warps on 148 SMs).  But each sweep is a first-order AFFINE recurrence in the flux,
//     F[i] = a_i F[i +- 1] - b_i F_opp[i] + s_i,        a = P/M, b = N/M, s = (fac * planck + beam)/M,
// because the opposite-direction flux F_opp that couples in is the other sweep's finished result.  So:
//
//   phase A (block-wide, lanes along the flat column index y + ny*x => coalesced): the coefficient arrays
//     stream in ONCE per flux solve; the four constants a, b, s_down, s_up of every (half-)layer are
//     formed with the rounding-exact building blocks of sweep_math.cuh and parked, transposed, in shared
//     memory.  HBM sees every input once and every output once per solve: the 80 B (iso) / 176 B (non-iso)
//     per cell of DESIGN.md instead of that figure times 2 sweeps times npass.
//   phase B (16 lanes per column, i.e. two columns per warp; lane = chunk of CH consecutive layers): every
//     lane composes the affine map of its chunk, a Kogge-Stone scan over the lanes (4 shuffle steps) gives each lane the flux
//     entering its chunk, and the lane then walks its own CH layers.  A sweep is 2*CH + 5 short steps
//     instead of nlayer long ones; all passes run back to back inside the warp with no block barrier, the
//     fluxes needed from the other direction / the previous pass stay in registers (chunk-edge values are
//     handed over by shuffle).  The passes are branch-free (cells outside the column are identity steps) and
//     store nothing; the fluxes of the last pass are stored from the registers afterwards.
//
//   Between two opacity refreshes the PLANNED sweeps of fband_plan.cu take over: the Planck-independent part of the
//   step constants is formed once per refresh and stored in the order the sweep lanes consume it, so the sweep has no
//   phase A at all.  This file remains the unplanned form (any call without a plan, > 128 / 256 layers).
//
// Arithmetic: a F - b F_opp + s is the reference's 1/M (P F - N F_opp + ...) with the division by M
// distributed, and the chunk-entry flux comes from composed maps: both reorder a few multiply-adds.
// Measured deviation from the reference's kernels: <= ~1e-13 relative (tests/test_gpu_parity.py; the bar
// is 1e-10).  The bit-faithful evaluation order lives in the column-serial kernels of fband.cu
// (helios_ctx_set_fband_mode(ctx, 1)); consecutive launches and one fused launch agree bit for bit.
#include <type_traits>
#include "common.cuh"
#include "sweep_math.cuh"
#include <cstdlib>

#ifndef ISO_NCOLS
#define ISO_NCOLS 8  // columns per tile of the isothermal sweep at 81..112 layers: 128-thread CTAs, 4 per SM (measured
                     // 1-2 % faster than 16 columns / 256 threads / 2 per SM: more CTAs in different phases overlap better)
#endif

struct CpScalars {
    double g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s_transition;
    int nint, nbin, ny, dir_beam, clouds, scat_corr, npass, nchunk, colpitch;
    long long reserved;     // (kept so that the positional initialisers below stay aligned)
    int no_beam;      // F_dir / Fc_dir are known to be all -0.0: do not load them, nor G+/-
    int nbatch;       // atmospheres per launch (helios_ctx_set_batch), 1 otherwise
    const int* done;  // batch: converged atmospheres are skipped (their fluxes stay as they are)
};

struct CpNonisoCoef {
    const double *w0_u, *w0_l, *dtau_u, *dtau_l, *dtc_u, *dtc_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u,
        *Gp_l, *Gm_u, *Gm_l;
};

struct Step {
    double a, b, sd, su;
};

__device__ __forceinline__ Step make_step(double M, double N, double P, double fac, double pt_d, double pt_u,
                                          double D_d, double D_u) {
    const double invM = 1.0 / M;
    Step s;
    s.a = invM * P;
    s.b = invM * N;
    s.sd = invM * (fac * pt_d + D_d);
    s.su = invM * (fac * pt_u + D_u);
    return s;
}

// affine maps x -> A x + B; compose(f, g) = f after g
struct Aff {
    double A, B;
};
__device__ __forceinline__ Aff aff_after(const Aff& f, const Aff& g) { return Aff{f.A * g.A, f.A * g.B + f.B}; }

// Kogge-Stone scans inside segments of LPC lanes (one segment = one column).
// from_top:    lane j receives the composition of the maps of lanes j, j+1, ..., n-1 (highest applied first)
// from_bottom: lane j receives the composition of the maps of lanes j, j-1, ..., 0   (lane 0 applied first)
template <int LPC>
__device__ __forceinline__ Aff scan_from_top(Aff m, int sl, int n) {
#pragma unroll
    for (int d = 1; d < LPC; d <<= 1) {
        Aff o;
        o.A = __shfl_down_sync(0xffffffffu, m.A, d, LPC);
        o.B = __shfl_down_sync(0xffffffffu, m.B, d, LPC);
        if (sl + d < n) m = aff_after(m, o);
    }
    return m;
}
template <int LPC>
__device__ __forceinline__ Aff scan_from_bottom(Aff m, int sl) {
#pragma unroll
    for (int d = 1; d < LPC; d <<= 1) {
        Aff o;
        o.A = __shfl_up_sync(0xffffffffu, m.A, d, LPC);
        o.B = __shfl_up_sync(0xffffffffu, m.B, d, LPC);
        if (sl >= d) m = aff_after(m, o);
    }
    return m;
}

// direct-beam sources of one (half-)layer; skipped when no beam reaches the layer (all-zero F_dir), which
// also saves the four fp64 divisions: min(0, +-0) contributes nothing to the sums it enters
__device__ __forceinline__ void beam_pair(double Fa, double Fb, double neg_mu, double M, double N, double P,
                                          double G_pl, double G_min, double m1d, double m2d, double& Dd,
                                          double& Du) {
    Dd = 0.0;
    Du = 0.0;
    if (Fa != 0.0 || Fb != 0.0) {
        Dd = beam_source(Fa, Fb, neg_mu, M, G_min, N, G_pl, m1d, m2d);
        Du = beam_source(Fb, Fa, neg_mu, N, G_min, M, G_pl, P, G_pl);
    }
}

// ------------------------------------------------------------------------------------------------
// The kernel.  NONISO = false: one step per layer, shared planes a, b, sd, su, F_up(prev)            (5)
//              NONISO = true : two steps per layer (upper half, lower half): [a, b, sd, su] x 2,
//                              F_up(prev), Fc_up(prev)                                               (10)
// Block: NCOLS columns, LPC lanes per column (NCOLS * LPC threads), CH = ceil(nlay / LPC) layers per lane.
// ------------------------------------------------------------------------------------------------
template <bool NONISO, int CH, int LPC, int NCOLS, bool FULL, bool HOIST>
__global__ void __launch_bounds__(NCOLS * LPC, (NCOLS * LPC <= 128) ? 4 : 2)
k_fband_wp(double* __restrict__ F_down, double* __restrict__ F_up, double* __restrict__ Fc_down,
           double* __restrict__ Fc_up, const double* __restrict__ F_dir, const double* __restrict__ Fc_dir,
           const double* __restrict__ planck_lay, const double* __restrict__ planck_int,
           // iso: w0_u = w_0, M_u = M_term, ... ; the *_l members are unused
           CpNonisoCoef cfg, const double* __restrict__ albedo, const double* __restrict__ g0_lay,
           const double* __restrict__ g0_int, CpScalars s) {
    extern __shared__ double sm[];
    constexpr int NPL = NONISO ? 10 : 5;
    constexpr int STRIDE = CH + ((CH % 2 == 0) ? 1 : 0);
    const int nint = s.nint, nlay = nint - 1, nch = s.nchunk, pitch = s.colpitch;
    const int ncol = s.nbin * s.ny;
    const size_t plane = (size_t)NCOLS * pitch;
    double* c_toa = sm + NPL * plane;  // per-column scalars, [NCOLS] each
    double* c_alb = c_toa + NCOLS;
    double* c_fdir0 = c_alb + NCOLS;
    double* c_emis = c_fdir0 + NCOLS;
    const double neg_mu = -s.mu_star;
    const int ntile = (ncol + NCOLS - 1) / NCOLS;
    for (int gtile = blockIdx.x; gtile < ntile * s.nbatch; gtile += gridDim.x) {
        // batch: tiles enumerate (atmosphere, column tile); per-atmosphere arrays are offset by the
        // reference's allocation sizes (every [i][x][y] array holds ninterface rows, Q:407)
        const int atm = gtile / ntile;
        const int tile = gtile - atm * ntile;
        if (s.done != nullptr && s.done[atm] != 0) continue;  // uniform per block
        const size_t wgo = (size_t)atm * ncol * nint;          // [i][x][y] arrays
        const size_t blo = (size_t)atm * s.nbin * nlay;        // [layer][x] arrays
        const size_t bio = (size_t)atm * s.nbin * nint;        // [interface][x] arrays
        // ================= phase A: coalesced streaming, lanes along columns =================
        // Every thread owns at most CH layer rows (nlay <= LPC*CH).
        {
            const int c = threadIdx.x % NCOLS;
            const int r = threadIdx.x / NCOLS;  // 0 .. LPC-1
            const int col = min(tile * NCOLS + c, ncol - 1);
            const int x = col / s.ny;
            const double* __restrict__ BL = planck_lay + (size_t)atm * (nlay + 2) * s.nbin + (size_t)x * (nlay + 2);
            const double* __restrict__ BI = NONISO ? planck_int + bio + (size_t)x * nint : nullptr;
            constexpr int NRAW = NONISO ? 24 : 11;
            // cells whose loads are in flight together per thread (2 or 4 were measured SLOWER for the isothermal
            // kernel on B200: register spills outweigh the extra memory-level parallelism)
            constexpr int UA = 1;
            double raw[UA][NRAW];
            auto load_cell = [&](int m, double* q) {
                const int i = r + LPC * m;
                if (i < nlay) {
                    const size_t e = wgo + col + (size_t)ncol * i;
                    const size_t bb = (size_t)x + (size_t)s.nbin * i;
                    if (s.no_beam) {
                        q[0] = q[1] = -0.0;  // what fdir_* wrote (trans.cu); G+/- only ever multiply the beam
                        q[6] = q[7] = 0.0;
                    } else {
                        q[0] = F_dir[e];
                        q[1] = F_dir[e + ncol];
                        q[6] = cfg.Gp_u[e]; q[7] = cfg.Gm_u[e];
                    }
                    q[2] = cfg.w0_u[e]; q[3] = cfg.M_u[e]; q[4] = cfg.N_u[e]; q[5] = cfg.P_u[e];
                    q[8] = BL[i];
                    q[9] = F_up[e];
                    q[10] = s.clouds ? g0_lay[blo + bb] : s.g_0;
                    if (NONISO) {
                        q[11] = cfg.w0_l[e]; q[12] = cfg.M_l[e]; q[13] = cfg.N_l[e]; q[14] = cfg.P_l[e];
                        if (s.no_beam) {
                            q[15] = q[16] = 0.0;
                            q[19] = -0.0;
                        } else {
                            q[15] = cfg.Gp_l[e]; q[16] = cfg.Gm_l[e];
                            q[19] = Fc_dir[e];
                        }
                        q[17] = cfg.dtau_u[e] + cfg.dtc_u[blo + bb];
                        q[18] = cfg.dtau_l[e] + cfg.dtc_l[blo + bb];
                        q[20] = BI[i];
                        q[21] = BI[i + 1];
                        q[22] = Fc_up[e];
                        q[23] = s.clouds ? g0_int[bio + bb] : s.g_0;
                        if (s.clouds) q[10] = (q[10] + g0_int[bio + bb + s.nbin]) / 2.0;  // g0 of the upper half
                    }
                }
            };
            auto emit_cell = [&](int m, const double* q) {
                {
                    const int i = r + LPC * m;
                    if (i < nlay) {
                        const int o = c * pitch + (i / CH) * STRIDE + (i % CH);
                        const double Fdir_i = q[0], Fdir_ip1 = q[1];
                        if (!NONISO) {
                            const double w0 = q[2], M = q[3], N = q[4], P = q[5];
                            const double E = s.scat_corr ? E_parameter(w0, q[10], s.i2s_transition) : 1.0;
                            double Dd, Du;
                            beam_pair(Fdir_i, Fdir_ip1, neg_mu, M, N, P, q[6], q[7], P, q[7], Dd, Du);
                            const double pt = planck_iso(q[8], M, N, P);
                            const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt, pt, Dd, Du);
                            sm[o] = st.a;
                            sm[plane + o] = st.b;
                            sm[2 * plane + o] = st.sd;
                            sm[3 * plane + o] = st.su;
                            sm[4 * plane + o] = q[9];  // upward flux of the previous flux solve at interface i
                            if (i == 0) {  // the BOA emission uses layer 0's w0 and E (K:1472)
                                const double A_s = albedo[x];
                                c_alb[c] = A_s;
                                c_fdir0[c] = Fdir_i;
                                c_emis[c] = boa_emission(A_s, w0, E, BL[nlay + 1]);
                                c_toa[c] = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * BL[nlay];
                            }
                        } else {
                            const double Blay = q[8], Bint_lo = q[20], Bint_hi = q[21], Fcdir = q[19];
                            double g0_up = s.g_0, g0_low = s.g_0;
                            if (s.clouds) {
                                g0_up = q[10];
                                // the layer value is recovered exactly: q[10] held (gl + gi_hi)/2 only for the upper
                                // half, so the lower half re-reads gl (cheap, L1-resident)
                                const double gl = g0_lay[blo + (size_t)x + (size_t)s.nbin * i];
                                g0_low = (q[23] + gl) / 2.0;
                            }
                            {   // ---- upper half: layer centre <-> interface i+1 (K:1640-1664, 1771-1795)
                                const double w0 = q[2], M = q[3], N = q[4], P = q[5], dt = q[17];
                                const double E = s.scat_corr ? E_parameter(w0, g0_up, s.i2s_transition) : 1.0;
                                double pt_d, pt_u;
                                if (dt < s.delta_tau_limit) {
                                    pt_d = planck_thin(Bint_hi, Blay, M, N, P);
                                    pt_u = pt_d;
                                } else {
                                    const double pre = gradient_factor(s.epsi, w0, g0_up, E);
                                    const double pgrad = __ddiv_rn(__dsub_rn(Blay, Bint_hi), dt);
                                    pt_d = planck_grad_down(Blay, Bint_hi, M, N, P, pre, pgrad);
                                    pt_u = planck_grad_up(Bint_hi, Blay, M, N, P, pre, pgrad);
                                }
                                double Dd, Du;
                                beam_pair(Fcdir, Fdir_ip1, neg_mu, M, N, P, q[6], q[7], q[7], P, Dd, Du);
                                const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt_d, pt_u, Dd, Du);
                                sm[o] = st.a;
                                sm[plane + o] = st.b;
                                sm[2 * plane + o] = st.sd;
                                sm[3 * plane + o] = st.su;
                            }
                            {   // ---- lower half: interface i <-> layer centre (K:1667-1691, 1744-1768)
                                const double w0 = q[11], M = q[12], N = q[13], P = q[14], dt = q[18];
                                const double E = s.scat_corr ? E_parameter(w0, g0_low, s.i2s_transition) : 1.0;
                                double pt_d, pt_u;
                                if (dt < s.delta_tau_limit) {
                                    pt_d = planck_thin(Bint_lo, Blay, M, N, P);
                                    pt_u = pt_d;
                                } else {
                                    const double pre = gradient_factor(s.epsi, w0, g0_low, E);
                                    const double pgrad = __ddiv_rn(__dsub_rn(Bint_lo, Blay), dt);
                                    pt_d = planck_grad_down(Bint_lo, Blay, M, N, P, pre, pgrad);
                                    pt_u = planck_grad_up(Blay, Bint_lo, M, N, P, pre, pgrad);
                                }
                                double Dd, Du;
                                beam_pair(Fdir_i, Fcdir, neg_mu, M, N, P, q[15], q[16], P, q[16], Dd, Du);
                                const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt_d, pt_u, Dd, Du);
                                sm[4 * plane + o] = st.a;
                                sm[5 * plane + o] = st.b;
                                sm[6 * plane + o] = st.sd;
                                sm[7 * plane + o] = st.su;
                                if (i == 0) {  // lower half of layer 0 feeds the BOA emission (K:1704)
                                    const double A_s = albedo[x];
                                    c_alb[c] = A_s;
                                    c_fdir0[c] = Fdir_i;
                                    c_emis[c] = boa_emission(A_s, w0, E, BL[nlay + 1]);
                                    c_toa[c] = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * BL[nlay];
                                }
                            }
                            sm[8 * plane + o] = q[9];
                            sm[9 * plane + o] = q[22];
                        }
                    }
                }
            };
#pragma unroll
            for (int m0 = 0; m0 < CH; m0 += UA) {
#pragma unroll
                for (int u = 0; u < UA; u++)  // all loads of UA cells are issued back to back, then consumed
                    if (m0 + u < CH) load_cell(m0 + u, raw[u]);
#pragma unroll
                for (int u = 0; u < UA; u++)
                    if (m0 + u < CH) emit_cell(m0 + u, raw[u]);
            }
        }
        __syncthreads();
        // ================= phase B: LPC lanes per column, lane = chunk of CH layers =================
        // Branch-free passes: cells outside the column are identity steps (a = 1, b = 0, s = 0), the arithmetic
        // runs unconditionally and the stores follow the last pass.  A lane without a chunk
        // (sl >= nch) recomputes chunk 0 -- nothing of it is ever consumed: the scans, the edge shuffles and the
        // stores all look at sl < nch.  FULL (nlay == nch * CH): no partial chunk, the "inside" selects vanish.
        {
            const int cw = threadIdx.x / LPC;  // column of this lane segment
            const int sl = threadIdx.x % LPC;  // chunk index
            const int col = tile * NCOLS + cw;
            const bool live = col < ncol;      // uniform per segment; dead segments still shuffle
            const int colc = live ? col : ncol - 1;
            const bool act = sl < nch;
            const int lo = sl * CH;
            const int hi = act ? min(lo + CH, nlay) : CH;  // idle lanes: chunk 0 is complete (nlay >= CH or nch == 1)
            const int base = cw * pitch + (act ? sl : 0) * STRIDE;
            const double toa = c_toa[cw], A_s = c_alb[cw], Fdir0 = c_fdir0[cw], emis = c_emis[cw];
            // step constants kept in registers: a, b of every step; sd / su are re-read from shared memory
            constexpr int NS = NONISO ? 2 : 1;  // steps per layer
            double a[NS][CH], b[NS][CH], Fu_reg[CH], Fd_reg[CH], Fcu_reg[CH], Fcd_reg[CH], cc[NS][CH];
            const int lo_c = act ? lo : 0;
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const bool in = FULL || lo_c + k < hi;
                a[0][k] = in ? sm[base + k] : 1.0;  // identity step outside the column
                b[0][k] = in ? sm[plane + base + k] : 0.0;
                if (NONISO) {
                    a[1][k] = in ? sm[4 * plane + base + k] : 1.0;
                    b[1][k] = in ? sm[5 * plane + base + k] : 0.0;
                }
                Fu_reg[k] = in ? sm[(NONISO ? 8 : 4) * plane + base + k] : 0.0;
                Fcu_reg[k] = (NONISO && in) ? sm[9 * plane + base + k] : 0.0;
                Fd_reg[k] = Fcd_reg[k] = 0.0;
            }
            const bool st_ok = live && act;
            const size_t off = wgo + colc + (size_t)ncol * lo;
            // Pass-invariant half of the scans (HOIST: isothermal kernel, long pass sequences): the A parts of the affine maps are
            // products of the a's only, so the chunk products, the multiplier each Kogge-Stone step applies to the
            // incoming B (0 where the step does not apply to this lane -- that replaces the select) and the final
            // prefix products are formed once per tile; a pass shuffles and multiply-adds the B parts only.
            double scA_dn = 1.0, scA_up = 1.0;
            double2* __restrict__ as2 = reinterpret_cast<double2*>(c_emis + NCOLS) + threadIdx.x;
            if constexpr (HOIST) {
                double Adn = 1.0, Aup = 1.0;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) Adn = a[0][k] * Adn;
#pragma unroll
                for (int k = 0; k < CH; k++) Aup = a[0][k] * Aup;
#pragma unroll
                for (int r = 0, d = 1; d < LPC; r++, d <<= 1) {
                    const double odn = __shfl_down_sync(0xffffffffu, Adn, d, LPC);
                    const double oup = __shfl_up_sync(0xffffffffu, Aup, d, LPC);
                    const bool okd = sl + d < nch, oku = sl >= d;
                    as2[r * (NCOLS * LPC)] = make_double2(okd ? Adn : 0.0, oku ? Aup : 0.0);
                    if (okd) Adn = Adn * odn;
                    if (oku) Aup = Aup * oup;
                }
                scA_dn = Adn;
                scA_up = Aup;
            }
            auto one_pass = [&]() -> double {
                // ---------------- downward sweep (per layer: upper half, then lower half) ----------------
                Aff m{1.0, 0.0};
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const bool in = FULL || lo_c + k < hi;
                    if (NONISO) {
                        cc[0][k] = (in ? sm[2 * plane + base + k] : 0.0) - b[0][k] * Fcu_reg[k];
                        m = Aff{a[0][k] * m.A, a[0][k] * m.B + cc[0][k]};
                        cc[1][k] = (in ? sm[6 * plane + base + k] : 0.0) - b[1][k] * Fu_reg[k];
                        m = Aff{a[1][k] * m.A, a[1][k] * m.B + cc[1][k]};
                    } else {
                        cc[0][k] = (in ? sm[2 * plane + base + k] : 0.0) - b[0][k] * Fu_reg[k];
                        m = Aff{HOIST ? 1.0 : a[0][k] * m.A, a[0][k] * m.B + cc[0][k]};
                    }
                }
                Aff sc;
                if constexpr (HOIST) {
                    double mB = m.B;
#pragma unroll
                    for (int r = 0, d = 1; d < LPC; r++, d <<= 1)
                        mB = __fma_rn(as2[r * (NCOLS * LPC)].x, __shfl_down_sync(0xffffffffu, mB, d, LPC), mB);
                    sc = Aff{scA_dn, mB};
                } else {
                    sc = scan_from_top<LPC>(m, sl, nch);
                }
                const double Fbot = sc.A * toa + sc.B;                   // flux leaving my chunk (interface lo)
                double F = __shfl_down_sync(0xffffffffu, Fbot, 1, LPC);  // = flux entering it (interface hi)
                if (sl >= nch - 1) F = toa;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const bool in = FULL || lo_c + k < hi;
                    if (NONISO) {
                        double Fn = tiny_to_abs(a[0][k] * F + cc[0][k]);
                        if (!FULL) Fn = in ? Fn : F;  // identity cells pass the entering flux through unchanged
                        Fcd_reg[k] = Fn;
                        double Fm = tiny_to_abs(a[1][k] * Fn + cc[1][k]);
                        if (!FULL) Fm = in ? Fm : F;
                        F = Fm;
                    } else {
                        double Fm = tiny_to_abs(a[0][k] * F + cc[0][k]);
                        if (!FULL) Fm = in ? Fm : F;
                        F = Fm;
                    }
                    Fd_reg[k] = F;
                }
                // the flux at my top interface as WALKED (and stored) by the lane above: every flux consumed
                // later is bit-identical to what the output arrays hold
                double Fd_hi = __shfl_down_sync(0xffffffffu, Fd_reg[0], 1, LPC);
                if (sl >= nch - 1) Fd_hi = toa;
                // ---------------- upward sweep (per layer: lower half, then upper half) ----------------
                double fu0 = __fma_rn(A_s, __dadd_rn(Fdir0, Fd_reg[0]), emis);  // surface, valid in lane 0 (K:1469-1474)
                fu0 = __shfl_sync(0xffffffffu, fu0, 0, LPC);
                m = Aff{1.0, 0.0};
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const bool in = FULL || lo_c + k < hi;
                    // identity cells above the column hold the entering flux (= Fd_hi), so no "inside" test here
                    const double Fd_top = (k + 1 < CH) ? Fd_reg[(k + 1) % CH] : Fd_hi;
                    if (NONISO) {
                        cc[1][k] = (in ? sm[7 * plane + base + k] : 0.0) - b[1][k] * Fcd_reg[k];
                        m = Aff{a[1][k] * m.A, a[1][k] * m.B + cc[1][k]};
                        cc[0][k] = (in ? sm[3 * plane + base + k] : 0.0) - b[0][k] * Fd_top;
                        m = Aff{a[0][k] * m.A, a[0][k] * m.B + cc[0][k]};
                    } else {
                        cc[0][k] = (in ? sm[3 * plane + base + k] : 0.0) - b[0][k] * Fd_top;
                        m = Aff{HOIST ? 1.0 : a[0][k] * m.A, a[0][k] * m.B + cc[0][k]};
                    }
                }
                if constexpr (HOIST) {
                    double mB = m.B;
#pragma unroll
                    for (int r = 0, d = 1; d < LPC; r++, d <<= 1)
                        mB = __fma_rn(as2[r * (NCOLS * LPC)].y, __shfl_up_sync(0xffffffffu, mB, d, LPC), mB);
                    sc = Aff{scA_up, mB};
                } else {
                    sc = scan_from_bottom<LPC>(m, sl);
                }
                const double Ftop = sc.A * fu0 + sc.B;          // flux leaving my chunk (interface hi)
                F = __shfl_up_sync(0xffffffffu, Ftop, 1, LPC);  // = flux entering it (interface lo)
                if (sl == 0) F = fu0;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const bool in = FULL || lo_c + k < hi;
                    Fu_reg[k] = F;  // interface lo+k: what the next pass's downward sweep reads
                    if (NONISO) {
                        // no tiny-value clean-up on Fc_up: the reference applies it to index i, not i-1 (K:1763)
                        double Fn = a[1][k] * F + cc[1][k];
                        if (!FULL) Fn = in ? Fn : F;
                        Fcu_reg[k] = Fn;
                        double Fm = tiny_to_abs(a[0][k] * Fn + cc[0][k]);
                        if (!FULL) Fm = in ? Fm : F;
                        F = Fm;
                    } else {
                        double Fm = tiny_to_abs(a[0][k] * F + cc[0][k]);
                        if (!FULL) Fm = in ? Fm : F;
                        F = Fm;
                    }
                }
                // the flux at my bottom interface as walked by the lane below (next pass, and what F_up holds there)
                const double Fu_lo = __shfl_up_sync(0xffffffffu, F, 1, LPC);
                Fu_reg[0] = (sl == 0) ? fu0 : Fu_lo;
                return F;  // the flux leaving my chunk upwards (interface hi)
            };
            double F_out = 0.0;
            for (int pass = 0; pass < s.npass; pass++) F_out = one_pass();
            // Only the fluxes of the last pass are stored, and every one of them is still in a register: the lane's
            // downward fluxes at interfaces lo..lo+CH-1, its upward fluxes at the same interfaces (Fu_reg[0] is the
            // value the lane below walked to, bit for bit) and, in the top lane, interface nlay.
            if (st_ok) {
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    if (FULL || lo + k < hi) {
                        const size_t e = off + (size_t)(k * ncol);
                        F_down[e] = Fd_reg[k];
                        F_up[e] = Fu_reg[k];
                        if (NONISO) {
                            Fc_down[e] = Fcd_reg[k];
                            Fc_up[e] = Fcu_reg[k];
                        }
                    }
                }
                if (sl == nch - 1) {
                    const size_t e = wgo + colc + (size_t)ncol * nlay;
                    F_down[e] = toa;
                    F_up[e] = F_out;
                }
            }
        }
        __syncthreads();  // the next tile overwrites the shared planes
    }
}

// ------------------------------------------------------------------------------------------------
// launch planning: LPC = 16 lanes per column (two columns per warp) while nlay <= 128, else 32
// ------------------------------------------------------------------------------------------------
template <bool NONISO, int CH, int LPC, int NCOLS>
static int launch_wp(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                     const double* F_dir, const double* Fc_dir, const double* planck_lay, const double* planck_int,
                     CpNonisoCoef c, const double* albedo, const double* g0_lay, const double* g0_int, CpScalars s,
                     int ncol) {
    const int nlay = s.nint - 1;
    const int nchunk = (nlay + CH - 1) / CH;
    int pitch = nchunk * (CH + ((CH % 2 == 0) ? 1 : 0));
    if (pitch % 2 == 0) pitch += 1;  // odd column pitch: the phase-A stores of 16 lanes hit distinct banks
    s.nchunk = nchunk;
    s.colpitch = pitch;
    // Hoisting the pass-invariant scan multipliers (HOIST) costs one extra scan per tile and log2(LPC) double2 of
    // shared memory per thread: measured on B200 it loses 10 % at 1 and 4 passes (C4 single pass, C5) and wins 16 % at
    // 1001 passes (C4 with scattering), so it is used for long pass sequences of the isothermal kernel only.
    const bool hoist = !NONISO && s.npass >= 8;
    const size_t smem = ((size_t)(NONISO ? 10 : 5) * NCOLS * pitch + 4 * NCOLS) * sizeof(double) +
                        (hoist ? (size_t)(LPC == 32 ? 5 : 4) * NCOLS * LPC * sizeof(double2) : 0);
    if (nchunk > LPC || smem > 227 * 1024) return -1;
    const int ntile = (ncol + NCOLS - 1) / NCOLS * ctx->batch.nbatch;
    s.nbatch = ctx->batch.nbatch;
    s.done = ctx->batch.active ? ctx->batch.done : nullptr;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    const int max_per_sm = (NCOLS * LPC <= 128) ? 4 : 2;  // __launch_bounds__
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    if (per_sm < 1) per_sm = 1;
    const int grid = ntile < ctx->num_sms * per_sm ? ntile : ctx->num_sms * per_sm;
    const bool full = nlay == nchunk * CH;
    auto kern = hoist ? (full ? k_fband_wp<NONISO, CH, LPC, NCOLS, true, !NONISO> : k_fband_wp<NONISO, CH, LPC, NCOLS, false, !NONISO>)
                      : (full ? k_fband_wp<NONISO, CH, LPC, NCOLS, true, false> : k_fband_wp<NONISO, CH, LPC, NCOLS, false, false>);
    HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NCOLS * LPC, smem, ctx->stream>>>(F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planck_lay,
                                                   planck_int, c, albedo, g0_lay, g0_int, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

template <bool NONISO>
static int dispatch_wp(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                       const double* F_dir, const double* Fc_dir, const double* planck_lay,
                       const double* planck_int, CpNonisoCoef c, const double* albedo, const double* g0_lay,
                       const double* g0_int, CpScalars s, int ncol) {
    const int nlay = s.nint - 1;
#define WP_ARGS ctx, F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planck_lay, planck_int, c, albedo, g0_lay, g0_int, s, ncol
    if constexpr (NONISO) {
        // twice the constants per layer: 32 lanes per column keep the per-lane register arrays short
        if (nlay <= 32) return launch_wp<NONISO, 1, 32, 8>(WP_ARGS);
        if (nlay <= 64) return launch_wp<NONISO, 2, 32, 8>(WP_ARGS);
        if (nlay <= 96) return launch_wp<NONISO, 3, 32, 8>(WP_ARGS);
        if (nlay <= 128) return launch_wp<NONISO, 4, 32, 8>(WP_ARGS);
        if (nlay <= 256) return launch_wp<NONISO, 8, 32, 4>(WP_ARGS);
    } else {
        if (nlay <= 16) return launch_wp<NONISO, 1, 16, 16>(WP_ARGS);
        if (nlay <= 32) return launch_wp<NONISO, 2, 16, 16>(WP_ARGS);
        if (nlay <= 48) return launch_wp<NONISO, 3, 16, 16>(WP_ARGS);
        if (nlay <= 80) return launch_wp<NONISO, 5, 16, 16>(WP_ARGS);
        if (nlay <= 112) return launch_wp<NONISO, 7, 16, ISO_NCOLS>(WP_ARGS);
        if (nlay <= 128) return launch_wp<NONISO, 8, 16, 16>(WP_ARGS);
        if (nlay <= 256) return launch_wp<NONISO, 8, 32, 8>(WP_ARGS);
    }
#undef WP_ARGS
    return -1;
}

// return HELIOS_OK when launched, -1 when the shape does not fit this scheme (the caller then falls back to
// the one-thread-per-column kernel of fband.cu)
int fband_iso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, const double* F_dir, const double* planck,
                     const double* w_0, const double* M, const double* N, const double* P, const double* Gp,
                     const double* Gm, const double* albedo, const double* g0tot, double g_0, double Rstar,
                     double a, int nint, int nbin, double f_factor, double mu_star, int ny, double epsi,
                     int dir_beam, int clouds, int scat_corr, double i2s, int npass) {
    CpScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, 0.0, i2s, nint, nbin, ny, dir_beam, clouds, scat_corr, npass, 0, 0, 0,
                (dir_beam == 0 && ctx->zero_beam[0] == F_dir) ? 1 : 0, 1, nullptr};
    CpNonisoCoef c{w_0, nullptr, nullptr, nullptr, nullptr, nullptr, M, nullptr, N, nullptr, P, nullptr, Gp, nullptr, Gm, nullptr};
    return dispatch_wp<false>(ctx, F_down, F_up, nullptr, nullptr, F_dir, nullptr, planck, nullptr, c, albedo, g0tot,
                              nullptr, s, nbin * ny);
}

int fband_noniso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                        const double* F_dir, const double* Fc_dir, const double* planck_lay,
                        const double* planck_int, CpNonisoCoef c, const double* albedo, const double* g0_lay,
                        const double* g0_int, double g_0, double Rstar, double a, int nint, int nbin,
                        double f_factor, double mu_star, int ny, double epsi, double delta_tau_limit,
                        int dir_beam, int clouds, int scat_corr, double i2s, int npass) {
    CpScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s, nint, nbin, ny, dir_beam, clouds,
                scat_corr, npass, 0, 0, 0,
                (dir_beam == 0 && ctx->zero_beam[0] == F_dir && ctx->zero_beam[1] == Fc_dir) ? 1 : 0, 1, nullptr};
    return dispatch_wp<true>(ctx, F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planck_lay, planck_int, c, albedo,
                             g0_lay, g0_int, s, nbin * ny);
}
