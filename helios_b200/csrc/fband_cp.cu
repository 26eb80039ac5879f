// Layer-parallel two-stream sweeps ("chunk-parallel" fband kernels) -- the default flux solver.
//
// Why.  The reference's fband_* (K:1366-1799) walks the ~100 layers of a column serially, twice per
// pass and 3*scat+1 passes per RT iteration: ~800 dependent steps, each a handful of fp64 divides and
// ~20 loads.  With one thread per column that chain is pure latency (an atmosphere of 385 x 20 columns
// occupies 60 warps on 148 SMs).  But each sweep is a first-order AFFINE recurrence in the flux,
//     F[i] = a_i F[i +- 1] - b_i F_opp[i] + s_i,        a = P/M, b = N/M, s = (fac * planck + beam)/M,
// because the opposite-direction flux F_opp that couples in is the other sweep's finished result.  So:
//
//   * a block owns COLS consecutive columns and cuts the layers into chunks of CH; thread (col, chunk)
//     composes its chunk's affine map locally, the per-chunk maps are exchanged through shared memory,
//     every thread folds the maps in front of it to get the flux entering its chunk, and then walks its
//     own CH layers.  A sweep is CH + (#chunks) two-FMA steps instead of nlayer long ones, and ~25x more
//     threads hide the fp64 latency.
//   * the four constants of every (half-)layer -- a, b, s_down, s_up -- are computed ONCE per flux solve
//     while the coefficient arrays stream in (coalesced: lanes run along the flat column index y + ny*x),
//     with the rounding-exact building blocks of sweep_math.cuh, and are parked in shared memory for all
//     passes.  HBM sees every input once and every output once per solve: the 80 B (iso) / 176 B
//     (non-iso) per cell of DESIGN.md instead of that figure times 2 sweeps times npass.
//   * the fluxes a thread needs from the other direction / the previous pass are the ones it produced
//     itself (or its chunk neighbour's edge value, handed over through shared memory), so they stay in
//     registers; only the last pass writes the flux arrays.
//
// Arithmetic: a F - b F_opp + s is the reference's 1/M (P F - N F_opp + ...) with the division by M
// distributed, and the chunk-entry flux comes from composed maps: both reorder a few multiply-adds.
// Measured deviation from the reference's kernels: <= ~1e-13 relative (tests/test_gpu_parity.py; the bar
// is 1e-10).  The bit-faithful evaluation order lives in the column-serial kernels of fband.cu
// (helios_ctx_set_fband_mode(ctx, 1)); consecutive launches and one fused launch agree bit for bit.
#include "common.cuh"
#include "sweep_math.cuh"
#include <cstdlib>

struct CpScalars {
    double g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s_transition;
    int nint, nbin, ny, dir_beam, clouds, scat_corr, npass, nchunk;
};

struct CpNonisoCoef {
    const double *w0_u, *w0_l, *dtau_u, *dtau_l, *dtc_u, *dtc_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u,
        *Gp_l, *Gm_u, *Gm_l;
};

// constants of one (half-)layer step
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

// shared memory carve-up (doubles): planes[NPL][nlay][COLS] | mDA mDB mUA mUB edgeD edgeU [nch][COLS] | fu0[COLS]
template <int NPL, int COLS>
struct CpSmem {
    double *cf, *mDA, *mDB, *mUA, *mUB, *edgeD, *edgeU, *fu0;
    size_t plane;
    __device__ CpSmem(double* sm, int nlay, int nch) {
        plane = (size_t)nlay * COLS;
        cf = sm;
        mDA = sm + NPL * plane;
        mDB = mDA + (size_t)nch * COLS;
        mUA = mDB + (size_t)nch * COLS;
        mUB = mUA + (size_t)nch * COLS;
        edgeD = mUB + (size_t)nch * COLS;
        edgeU = edgeD + (size_t)nch * COLS;
        fu0 = edgeU + (size_t)nch * COLS;
    }
};

// ------------------------------------------------------------------------------------------------
// isothermal layers: one step per layer, planes a, b, sd, su
// ------------------------------------------------------------------------------------------------
template <int COLS, int CH, int MINB>
__global__ void __launch_bounds__(32 * COLS, MINB)
k_fband_iso_cp(double* __restrict__ F_down, double* __restrict__ F_up, const double* __restrict__ F_dir,
               const double* __restrict__ planck, const double* __restrict__ w_0, const double* __restrict__ Mt,
               const double* __restrict__ Nt, const double* __restrict__ Pt, const double* __restrict__ Gp,
               const double* __restrict__ Gm, const double* __restrict__ albedo,
               const double* __restrict__ g0tot, CpScalars s) {
    extern __shared__ double sm[];
    const int nint = s.nint, nlay = nint - 1, nch = s.nchunk;
    const int ncol = s.nbin * s.ny;
    const int c = threadIdx.x % COLS;
    const int w = threadIdx.x / COLS;
    const int lo = w * CH;
    const int hi = min(lo + CH, nlay);  // layers lo .. hi-1, interfaces lo .. hi
    CpSmem<4, COLS> S(sm, nlay, nch);
    const size_t plane = S.plane;
    double* cf = S.cf;
    const double neg_mu = -s.mu_star;
    const int ntile = (ncol + COLS - 1) / COLS;

    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int col = tile * COLS + c;
        const bool live = col < ncol;
        const int colc = live ? col : ncol - 1;  // dead lanes shadow the last column and never store
        const int x = colc / s.ny;
        const double* __restrict__ B = planck + (size_t)x * (nlay + 2);
        const double A_s = albedo[x];
        const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * B[nlay];

        // ---- stream the coefficients in once; park a, b, s_down, s_up of every layer in shared memory
        double Fu_reg[CH], Fd_reg[CH];
        double w0_0 = 0.0, E_0 = 1.0, Fdir0 = 0.0;  // the BOA emission uses layer 0's w0 and E (K:1472)
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int i = lo + k;
            Fu_reg[k] = 0.0;
            Fd_reg[k] = 0.0;
            if (i < hi) {
                const size_t e = colc + (size_t)ncol * i;
                const double w0 = w_0[e], M = Mt[e], N = Nt[e], P = Pt[e], G_pl = Gp[e], G_min = Gm[e];
                const double Fdir_i = F_dir[e], Fdir_ip1 = F_dir[e + ncol];
                const double g0 = s.clouds ? g0tot[x + (size_t)s.nbin * i] : s.g_0;
                const double E = s.scat_corr ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
                const double pt = planck_iso(B[i], M, N, P);
                const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt, pt,
                                          beam_source(Fdir_i, Fdir_ip1, neg_mu, M, G_min, N, G_pl, P, G_min),
                                          beam_source(Fdir_ip1, Fdir_i, neg_mu, N, G_min, M, G_pl, P, G_pl));
                const size_t o = (size_t)i * COLS + c;
                cf[o] = st.a;
                cf[plane + o] = st.b;
                cf[2 * plane + o] = st.sd;
                cf[3 * plane + o] = st.su;
                Fu_reg[k] = F_up[e];  // upward flux of the previous flux solve at interface i
                w0_0 = i == 0 ? w0 : w0_0;
                E_0 = i == 0 ? E : E_0;
                Fdir0 = i == 0 ? Fdir_i : Fdir0;
            }
        }
        const double B_surf = B[nlay + 1];

        for (int pass = 0; pass < s.npass; pass++) {
            const bool last = pass == s.npass - 1;
            double cc[CH];
            // ================= downward sweep =================
            {
                double A = 1.0, Bm = 0.0;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const int i = lo + k;
                    cc[k] = 0.0;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        const double a = cf[o];
                        cc[k] = cf[2 * plane + o] - cf[plane + o] * Fu_reg[k];
                        A = a * A;
                        Bm = a * Bm + cc[k];
                    }
                }
                S.mDA[w * COLS + c] = A;
                S.mDB[w * COLS + c] = Bm;
            }
            __syncthreads();
            {
                double F = toa;
                for (int v = nch - 1; v > w; v--) F = S.mDA[v * COLS + c] * F + S.mDB[v * COLS + c];
                if (last && live && w == nch - 1) F_down[col + (size_t)ncol * nlay] = toa;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const int i = lo + k;
                    if (i < hi) {
                        F = tiny_to_abs(cf[(size_t)i * COLS + c] * F + cc[k]);
                        Fd_reg[k] = F;
                        if (last && live) F_down[col + (size_t)ncol * i] = F;
                    }
                }
                S.edgeD[w * COLS + c] = Fd_reg[0];
            }
            // ================= upward sweep =================
            if (w == 0) S.fu0[c] = boa_flux(A_s, Fdir0, Fd_reg[0], w0_0, E_0, B_surf);
            __syncthreads();
            // the flux at the chunk's top interface is the value the chunk above WALKED (and stored), not the
            // composed one: every flux consumed later is bit-identical to what the output arrays hold
            const double Fd_hi = (w == nch - 1) ? toa : S.edgeD[(w + 1) * COLS + c];
            {
                double A = 1.0, Bm = 0.0;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const int i = lo + k;
                    cc[k] = 0.0;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        const double Fd_top = (k + 1 < CH && i + 1 < hi) ? Fd_reg[(k + 1) % CH] : Fd_hi;
                        const double a = cf[o];
                        cc[k] = cf[3 * plane + o] - cf[plane + o] * Fd_top;
                        A = a * A;
                        Bm = a * Bm + cc[k];
                    }
                }
                S.mUA[w * COLS + c] = A;
                S.mUB[w * COLS + c] = Bm;
            }
            __syncthreads();
            {
                double F = S.fu0[c];
                for (int v = 0; v < w; v++) F = S.mUA[v * COLS + c] * F + S.mUB[v * COLS + c];
                if (last && live && w == 0) F_up[col] = F;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const int i = lo + k;
                    if (i < hi) {
                        Fu_reg[k] = F;  // interface i: what the next pass's downward sweep reads
                        F = tiny_to_abs(cf[(size_t)i * COLS + c] * F + cc[k]);
                        if (last && live) F_up[col + (size_t)ncol * (i + 1)] = F;
                    }
                }
                S.edgeU[w * COLS + c] = F;
            }
            __syncthreads();
            Fu_reg[0] = (w == 0) ? S.fu0[c] : S.edgeU[(w - 1) * COLS + c];
        }
        __syncthreads();  // the next tile overwrites the shared planes
    }
}

// ------------------------------------------------------------------------------------------------
// non-isothermal layers: two steps per layer (upper / lower half), planes [half][a, b, sd, su]
// ------------------------------------------------------------------------------------------------
#define CF(q) cf[(size_t)(q) * plane + o]

template <int COLS, int CH, int MINB>
__global__ void __launch_bounds__(32 * COLS, MINB)
k_fband_noniso_cp(double* __restrict__ F_down, double* __restrict__ F_up, double* __restrict__ Fc_down,
                  double* __restrict__ Fc_up, const double* __restrict__ F_dir, const double* __restrict__ Fc_dir,
                  const double* __restrict__ planck_lay, const double* __restrict__ planck_int, CpNonisoCoef cfg,
                  const double* __restrict__ albedo, const double* __restrict__ g0_lay,
                  const double* __restrict__ g0_int, CpScalars s) {
    extern __shared__ double sm[];
    const int nint = s.nint, nlay = nint - 1, nch = s.nchunk;
    const int ncol = s.nbin * s.ny;
    const int c = threadIdx.x % COLS;
    const int w = threadIdx.x / COLS;
    const int lo = w * CH;
    const int hi = min(lo + CH, nlay);
    CpSmem<8, COLS> S(sm, nlay, nch);
    const size_t plane = S.plane;
    double* cf = S.cf;
    const double neg_mu = -s.mu_star;
    const int ntile = (ncol + COLS - 1) / COLS;

    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int col = tile * COLS + c;
        const bool live = col < ncol;
        const int colc = live ? col : ncol - 1;
        const int x = colc / s.ny;
        const double* __restrict__ BL = planck_lay + (size_t)x * (nlay + 2);
        const double* __restrict__ BI = planck_int + (size_t)x * nint;
        const double A_s = albedo[x];
        const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * BL[nlay];

        double Fu_reg[CH], Fcu_reg[CH], Fd_reg[CH], Fcd_reg[CH];
        double w0_0 = 0.0, E_0 = 1.0, Fdir0 = 0.0;  // lower half of layer 0 feeds the BOA emission (K:1704)
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int i = lo + k;
            Fu_reg[k] = Fcu_reg[k] = Fd_reg[k] = Fcd_reg[k] = 0.0;
            if (i < hi) {
                const size_t e = colc + (size_t)ncol * i;
                const size_t b = (size_t)x + (size_t)s.nbin * i;
                const double Blay = BL[i], Bint_lo = BI[i], Bint_hi = BI[i + 1];
                const double Fdir_i = F_dir[e], Fdir_ip1 = F_dir[e + ncol], Fcdir = Fc_dir[e];
                double g0_up = s.g_0, g0_low = s.g_0;
                if (s.clouds) {
                    const double gl = g0_lay[b];
                    g0_up = (gl + g0_int[b + s.nbin]) / 2.0;
                    g0_low = (g0_int[b] + gl) / 2.0;
                }
                const size_t o = (size_t)i * COLS + c;
                {   // ---- upper half: layer centre <-> interface i+1 (K:1640-1664, 1771-1795)
                    const double w0 = cfg.w0_u[e], M = cfg.M_u[e], N = cfg.N_u[e], P = cfg.P_u[e];
                    const double G_pl = cfg.Gp_u[e], G_min = cfg.Gm_u[e];
                    const double dt = cfg.dtau_u[e] + cfg.dtc_u[b];
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
                    const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt_d, pt_u,
                                              beam_source(Fcdir, Fdir_ip1, neg_mu, M, G_min, N, G_pl, G_min, P),
                                              beam_source(Fdir_ip1, Fcdir, neg_mu, N, G_min, M, G_pl, P, G_pl));
                    CF(0) = st.a; CF(1) = st.b; CF(2) = st.sd; CF(3) = st.su;
                }
                {   // ---- lower half: interface i <-> layer centre (K:1667-1691, 1744-1768)
                    const double w0 = cfg.w0_l[e], M = cfg.M_l[e], N = cfg.N_l[e], P = cfg.P_l[e];
                    const double G_pl = cfg.Gp_l[e], G_min = cfg.Gm_l[e];
                    const double dt = cfg.dtau_l[e] + cfg.dtc_l[b];
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
                    const Step st = make_step(M, N, P, source_factor(s.epsi, w0, E), pt_d, pt_u,
                                              beam_source(Fdir_i, Fcdir, neg_mu, M, G_min, N, G_pl, P, G_min),
                                              beam_source(Fcdir, Fdir_i, neg_mu, N, G_min, M, G_pl, P, G_pl));
                    CF(4) = st.a; CF(5) = st.b; CF(6) = st.sd; CF(7) = st.su;
                    w0_0 = i == 0 ? w0 : w0_0;
                    E_0 = i == 0 ? E : E_0;
                    Fdir0 = i == 0 ? Fdir_i : Fdir0;
                }
                Fu_reg[k] = F_up[e];
                Fcu_reg[k] = Fc_up[e];
            }
        }
        const double B_surf = BL[nlay + 1];

        for (int pass = 0; pass < s.npass; pass++) {
            const bool last = pass == s.npass - 1;
            double ccu[CH], ccl[CH];
            // ================= downward sweep: upper half, then lower half of every layer =================
            {
                double A = 1.0, Bm = 0.0;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const int i = lo + k;
                    ccu[k] = ccl[k] = 0.0;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        double a = CF(0);
                        ccu[k] = CF(2) - CF(1) * Fcu_reg[k];
                        A = a * A;
                        Bm = a * Bm + ccu[k];
                        a = CF(4);
                        ccl[k] = CF(6) - CF(5) * Fu_reg[k];
                        A = a * A;
                        Bm = a * Bm + ccl[k];
                    }
                }
                S.mDA[w * COLS + c] = A;
                S.mDB[w * COLS + c] = Bm;
            }
            __syncthreads();
            {
                double F = toa;
                for (int v = nch - 1; v > w; v--) F = S.mDA[v * COLS + c] * F + S.mDB[v * COLS + c];
                if (last && live && w == nch - 1) F_down[col + (size_t)ncol * nlay] = toa;
#pragma unroll
                for (int k = CH - 1; k >= 0; k--) {
                    const int i = lo + k;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        F = tiny_to_abs(CF(0) * F + ccu[k]);
                        Fcd_reg[k] = F;
                        if (last && live) Fc_down[col + (size_t)ncol * i] = F;
                        F = tiny_to_abs(CF(4) * F + ccl[k]);
                        Fd_reg[k] = F;
                        if (last && live) F_down[col + (size_t)ncol * i] = F;
                    }
                }
                S.edgeD[w * COLS + c] = Fd_reg[0];
            }
            // ================= upward sweep: lower half, then upper half =================
            if (w == 0) S.fu0[c] = boa_flux(A_s, Fdir0, Fd_reg[0], w0_0, E_0, B_surf);
            __syncthreads();
            const double Fd_hi = (w == nch - 1) ? toa : S.edgeD[(w + 1) * COLS + c];  // walked value (see iso)
            {
                double A = 1.0, Bm = 0.0;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const int i = lo + k;
                    ccu[k] = ccl[k] = 0.0;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        const double Fd_top = (k + 1 < CH && i + 1 < hi) ? Fd_reg[(k + 1) % CH] : Fd_hi;
                        double a = CF(4);
                        ccl[k] = CF(7) - CF(5) * Fcd_reg[k];
                        A = a * A;
                        Bm = a * Bm + ccl[k];
                        a = CF(0);
                        ccu[k] = CF(3) - CF(1) * Fd_top;
                        A = a * A;
                        Bm = a * Bm + ccu[k];
                    }
                }
                S.mUA[w * COLS + c] = A;
                S.mUB[w * COLS + c] = Bm;
            }
            __syncthreads();
            {
                double F = S.fu0[c];
                for (int v = 0; v < w; v++) F = S.mUA[v * COLS + c] * F + S.mUB[v * COLS + c];
                if (last && live && w == 0) F_up[col] = F;
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const int i = lo + k;
                    if (i < hi) {
                        const size_t o = (size_t)i * COLS + c;
                        Fu_reg[k] = F;
                        // no tiny-value clean-up on Fc_up: the reference applies it to index i, not i-1 (K:1763)
                        F = CF(4) * F + ccl[k];
                        Fcu_reg[k] = F;
                        if (last && live) Fc_up[col + (size_t)ncol * i] = F;
                        F = tiny_to_abs(CF(0) * F + ccu[k]);
                        if (last && live) F_up[col + (size_t)ncol * (i + 1)] = F;
                    }
                }
                S.edgeU[w * COLS + c] = F;
            }
            __syncthreads();
            Fu_reg[0] = (w == 0) ? S.fu0[c] : S.edgeU[(w - 1) * COLS + c];
        }
        __syncthreads();
    }
}
#undef CF

// ------------------------------------------------------------------------------------------------
// launch planning: pick (COLS, CH) so that the block fits (<= 32 chunks, <= 227 kB shared memory)
// ------------------------------------------------------------------------------------------------
struct CpPlan {
    int cols, ch, nchunk, threads;
    size_t smem;
    int grid;
};

static bool cp_plan(helios_ctx* ctx, int nlay, int ncol, int planes, int cols, int ch, int minb, CpPlan* p) {
    const int nchunk = (nlay + ch - 1) / ch;
    const int threads = nchunk * cols;
    const size_t smem = ((size_t)planes * nlay * cols + (size_t)6 * nchunk * cols + cols) * sizeof(double);
    if (nchunk > 32 || threads > 1024 || smem > 227 * 1024) return false;
    p->cols = cols;
    p->ch = ch;
    p->nchunk = nchunk;
    p->threads = threads;
    p->smem = smem;
    const int ntile = (ncol + cols - 1) / cols;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    const int by_threads = 2048 / ((threads + 31) / 32 * 32);
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm > minb) per_sm = minb;  // registers are budgeted for `minb` resident blocks
    if (per_sm < 1) per_sm = 1;
    const int cap = ctx->num_sms * per_sm;
    p->grid = ntile < cap ? ntile : cap;
    return true;
}

template <int COLS, int CH, int MINB>
static int launch_iso(helios_ctx* ctx, const CpPlan& p, double* F_down, double* F_up, const double* F_dir,
                      const double* planck, const double* w_0, const double* M, const double* N, const double* P,
                      const double* Gp, const double* Gm, const double* albedo, const double* g0tot, CpScalars s) {
    auto kern = k_fband_iso_cp<COLS, CH, MINB>;
    HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<p.grid, p.threads, p.smem, ctx->stream>>>(F_down, F_up, F_dir, planck, w_0, M, N, P, Gp, Gm, albedo,
                                                      g0tot, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

template <int COLS, int CH, int MINB>
static int launch_noniso(helios_ctx* ctx, const CpPlan& p, double* F_down, double* F_up, double* Fc_down,
                         double* Fc_up, const double* F_dir, const double* Fc_dir, const double* planck_lay,
                         const double* planck_int, CpNonisoCoef c, const double* albedo, const double* g0_lay,
                         const double* g0_int, CpScalars s) {
    auto kern = k_fband_noniso_cp<COLS, CH, MINB>;
    HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<p.grid, p.threads, p.smem, ctx->stream>>>(F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planck_lay,
                                                      planck_int, c, albedo, g0_lay, g0_int, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

static int cp_variant() {
    const char* v = getenv("HELIOS_CP_VARIANT");  // tuning hook (scratch/tune_fband.py); 0 = first that fits
    return v ? atoi(v) : 0;
}

// returns HELIOS_OK when launched, -1 when the shape does not fit this scheme (caller falls back to the
// one-thread-per-column kernel of fband.cu)
int fband_iso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, const double* F_dir, const double* planck,
                     const double* w_0, const double* M, const double* N, const double* P, const double* Gp,
                     const double* Gm, const double* albedo, const double* g0tot, double g_0, double Rstar,
                     double a, int nint, int nbin, double f_factor, double mu_star, int ny, double epsi,
                     int dir_beam, int clouds, int scat_corr, double i2s, int npass) {
    const int nlay = nint - 1, ncol = nbin * ny;
    CpPlan p;
    CpScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, 0.0, i2s, nint, nbin, ny, dir_beam, clouds, scat_corr, npass, 0};
    const int var = cp_variant();
#define TRY_ISO(ID, COLS, CH, MINB)                                                                        \
    if ((var == 0 || var == ID) && cp_plan(ctx, nlay, ncol, 4, COLS, CH, MINB, &p)) {                      \
        s.nchunk = p.nchunk;                                                                               \
        return launch_iso<COLS, CH, MINB>(ctx, p, F_down, F_up, F_dir, planck, w_0, M, N, P, Gp, Gm, albedo, g0tot, s); \
    }
    TRY_ISO(2, 16, 4, 2)
    TRY_ISO(1, 16, 4, 3)
    TRY_ISO(3, 16, 5, 3)
    TRY_ISO(4, 8, 5, 4)
    TRY_ISO(5, 16, 8, 2)
    TRY_ISO(6, 8, 8, 2)
    TRY_ISO(7, 8, 16, 2)
#undef TRY_ISO
    return -1;
}

int fband_noniso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                        const double* F_dir, const double* Fc_dir, const double* planck_lay,
                        const double* planck_int, CpNonisoCoef c, const double* albedo, const double* g0_lay,
                        const double* g0_int, double g_0, double Rstar, double a, int nint, int nbin,
                        double f_factor, double mu_star, int ny, double epsi, double delta_tau_limit,
                        int dir_beam, int clouds, int scat_corr, double i2s, int npass) {
    const int nlay = nint - 1, ncol = nbin * ny;
    CpPlan p;
    CpScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s, nint, nbin, ny, dir_beam, clouds, scat_corr, npass, 0};
    const int var = cp_variant();
#define TRY_NONISO(ID, COLS, CH, MINB)                                                                     \
    if ((var == 0 || var == ID) && cp_plan(ctx, nlay, ncol, 8, COLS, CH, MINB, &p)) {                      \
        s.nchunk = p.nchunk;                                                                               \
        return launch_noniso<COLS, CH, MINB>(ctx, p, F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planck_lay, \
                                             planck_int, c, albedo, g0_lay, g0_int, s);                    \
    }
    TRY_NONISO(2, 8, 4, 3)
    TRY_NONISO(1, 8, 5, 3)
    TRY_NONISO(3, 16, 4, 1)
    TRY_NONISO(4, 8, 5, 2)
    TRY_NONISO(5, 16, 5, 1)
    TRY_NONISO(6, 8, 8, 2)
    TRY_NONISO(7, 4, 8, 2)
#undef TRY_NONISO
    return -1;
}
