// Transmission functions / two-stream coefficients, layer heights and the direct stellar beam.
// From-scratch sm_100a kernels for K:128-290 and K:1015-1362 of the reference.
#include "common.cuh"

struct TransScalars {
    double g_0, epsi, epsi2, mu_star, w_0_limit, w_0_scat_limit, i2s_transition;
    int scat, nbin, ny, nlayer, clouds, scat_corr, debug;
    double inv_eps2, inv_mu2;  // pow(epsi, -2.0), pow(mu_star, -2.0): see trans_constants()
    int nbatch;                // atmospheres per launch (helios_ctx_set_batch), 1 otherwise
};

struct CellCoeffs {
    double w0, dtau, trans, M, N, P, Gp, Gm;
};

// G+/- limiter (K:218-231)
__device__ __forceinline__ double g_limit(double G) {
    return fabs(G) < 1e8 ? G : 1e8 * G / fabs(G);
}

// everything calc_trans_* derives for one (half-)layer cell: K:1076-1099 / K:1199-1237 with the helper
// functions K:128-290 inlined.  E, sqrt((E-w0)/(E(1-w0 g0))) and the G denominator are formed once
// (the reference re-evaluates them in five device functions).
__device__ __forceinline__ CellCoeffs cell_coeffs(double ray, double cloud_scat, double cloud_abs,
                                                  double opac, double mmm, double dcol, double dtau_cloud,
                                                  double g0, const TransScalars& s) {
    CellCoeffs c;
    c.w0 = fmin((ray + cloud_scat) / ((ray + cloud_scat) + (opac * mmm + cloud_abs)), s.w_0_limit);
    c.dtau = dcol * (opac + ray / mmm);
    const double del_tau = c.dtau + dtau_cloud;
    const double w0 = c.w0;
    const double E = s.scat_corr == 1 ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
    const double one_m_wg = 1.0 - w0 * g0;
    c.trans = exp(-1.0 / s.epsi * sqrt(E * one_m_wg * (E - w0)) * del_tau);
    const double root = sqrt((E - w0) / (E * one_m_wg));
    const double zm = 0.5 * (1.0 - root);
    const double zp = 0.5 * (1.0 + root);
    const double t2 = c.trans * c.trans;
    c.M = (zm * zm) * t2 - (zp * zp);
    c.N = zp * zm * (1.0 - t2);
    c.P = ((zm * zm) - (zp * zp)) * c.trans;
    // G+ / G- (K:149-213)
    const double num = w0 * (E * one_m_wg + g0 * s.epsi / s.epsi2);
    const double denom = E * s.inv_eps2 * (E - w0) * one_m_wg - s.inv_mu2;
    const double inv_eps = 1.0 / s.epsi;
    const double cross = 1.0 / (s.mu_star * E * one_m_wg);
    const double third = s.epsi * w0 * g0 * s.mu_star / (s.epsi2 * E * one_m_wg);
    c.Gp = g_limit(0.5 * (num / denom * (inv_eps + cross) + third));
    c.Gm = g_limit(0.5 * (num / denom * (inv_eps - cross) - third));
    return c;
}

// The G+/- denominator contains pow(epsi, -2.0) and pow(mu_star, -2.0) (K:168, K:202): loop-invariant
// scalars that the reference re-evaluates with the generic libdevice pow in every cell (four calls per cell).
// They are evaluated ONCE here -- on the device, with the same libdevice pow, because the denominator cancels
// to O(w0) and a last-bit difference of a host pow would be amplified -- and cached in the context until epsi
// or mu_star change.
__global__ void k_trans_constants(double epsi, double mu_star, double* __restrict__ out) {
    out[0] = pow(epsi, -2.0);
    out[1] = pow(mu_star, -2.0);
}

static int trans_constants(helios_ctx* ctx, TransScalars& s) {
    if (!(ctx->trans_cache_valid && ctx->trans_cache[0] == s.epsi && ctx->trans_cache[1] == s.mu_star)) {
        double* d = nullptr;
        int rc = helios_ctx_scratch(ctx, 2 * sizeof(double), &d);
        if (rc != HELIOS_OK) return rc;
        k_trans_constants<<<1, 1, 0, ctx->stream>>>(s.epsi, s.mu_star, d);
        HLAUNCHED(ctx);
        double h[2];
        HCUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        HCUDA(cudaStreamSynchronize(ctx->stream));
        ctx->trans_cache[0] = s.epsi;
        ctx->trans_cache[1] = s.mu_star;
        ctx->trans_cache[2] = h[0];
        ctx->trans_cache[3] = h[1];
        ctx->trans_cache_valid = true;
    }
    s.inv_eps2 = ctx->trans_cache[2];
    s.inv_mu2 = ctx->trans_cache[3];
    return HELIOS_OK;
}

// one thread per cell of the [i][x][y] arrays, flat index -> fully coalesced 8-array store
__global__ void __launch_bounds__(256)
k_calc_trans_iso(double* __restrict__ trans_wg, double* __restrict__ delta_tau_wg, double* __restrict__ M_term,
                 double* __restrict__ N_term, double* __restrict__ P_term, double* __restrict__ G_plus,
                 double* __restrict__ G_minus, const double* __restrict__ delta_colmass,
                 const double* __restrict__ opac_wg_lay, const double* __restrict__ meanmolmass_lay,
                 const double* __restrict__ scat_cross_lay, const double* __restrict__ abs_cross_cl,
                 const double* __restrict__ scat_cross_cl, double* __restrict__ delta_tau_all_clouds,
                 double* __restrict__ w_0, const double* __restrict__ g_0_tot_lay,
                 int* __restrict__ scat_trigger, TransScalars s) {
    const int ncol = s.nbin * s.ny;
    const long long per_atm = (long long)ncol * s.nlayer;
    const long long total = per_atm * s.nbatch;
    for (long long ee = blockIdx.x * (long long)blockDim.x + threadIdx.x; ee < total;
         ee += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(ee / per_atm);                       // atmosphere (0 outside batch mode)
        const long long el = ee - (long long)a * per_atm;
        const int i = (int)(el / ncol);
        const int col = (int)(el - (long long)i * ncol);
        const int x = col / s.ny;
        const int y = col - x * s.ny;
        const size_t e = (size_t)a * ncol * (s.nlayer + 1) + el;  // [i][x][y] arrays hold ninterface rows (Q:407)
        const size_t b = (size_t)a * s.nbin * s.nlayer + (size_t)x + (size_t)s.nbin * i;
        const size_t li = (size_t)a * s.nlayer + i;
        const double g0 = s.clouds == 1 ? g_0_tot_lay[b] : s.g_0;
        const double ray = s.scat == 1 ? scat_cross_lay[b] : 0.0;
        const double csc = s.scat == 1 ? scat_cross_cl[b] : 0.0;
        const double cab = abs_cross_cl[b];
        const double mmm = meanmolmass_lay[li];
        const double dcol = delta_colmass[li];
        const double dtc = dcol * (cab + csc) / mmm;
        if (y == 0) delta_tau_all_clouds[b] = dtc;
        const CellCoeffs c = cell_coeffs(ray, csc, cab, opac_wg_lay[e], mmm, dcol, dtc, g0, s);
        w_0[e] = c.w0;
        delta_tau_wg[e] = c.dtau;
        trans_wg[e] = c.trans;
        M_term[e] = c.M;
        N_term[e] = c.N;
        P_term[e] = c.P;
        G_plus[e] = c.Gp;
        G_minus[e] = c.Gm;
        if (c.w0 > s.w_0_scat_limit) scat_trigger[(size_t)a * ncol + col] = 1;  // benign race, as K:1102
    }
}

struct NonisoOut {
    double *trans_u, *trans_l, *dtau_u, *dtau_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u, *Gp_l, *Gm_u,
        *Gm_l, *dtc_u, *dtc_l, *w0_u, *w0_l;
};
struct NonisoIn {
    const double *dcol_u, *dcol_l, *opac_lay, *opac_int, *mmm_lay, *mmm_int, *scat_lay, *scat_int,
        *cab_lay, *cab_int, *csc_lay, *csc_int, *g0_lay, *g0_int;
};

__global__ void __launch_bounds__(256)
k_calc_trans_noniso(NonisoOut o, NonisoIn in, int* __restrict__ scat_trigger, TransScalars s) {
    const int ncol = s.nbin * s.ny;
    const long long per_atm = (long long)ncol * s.nlayer;
    const long long total = per_atm * s.nbatch;
    for (long long ee = blockIdx.x * (long long)blockDim.x + threadIdx.x; ee < total;
         ee += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(ee / per_atm);                       // atmosphere (0 outside batch mode)
        const long long el = ee - (long long)a * per_atm;
        const int i = (int)(el / ncol);
        const int col = (int)(el - (long long)i * ncol);
        const int x = col / s.ny;
        const int y = col - x * s.ny;
        const size_t e = (size_t)a * ncol * (s.nlayer + 1) + el;                         // [i][x][y] arrays
        const size_t b = (size_t)a * s.nbin * s.nlayer + (size_t)x + (size_t)s.nbin * i;        // [layer][x]
        const size_t bi = (size_t)a * s.nbin * (s.nlayer + 1) + (size_t)x + (size_t)s.nbin * i;  // [interface][x]
        const size_t bu = bi + s.nbin;                                                           // interface i+1
        const size_t li = (size_t)a * s.nlayer + i, ii = (size_t)a * (s.nlayer + 1) + i;
        double g0_up = s.g_0, g0_low = s.g_0;
        if (s.clouds == 1) {
            g0_up = (in.g0_lay[b] + in.g0_int[bu]) / 2.0;
            g0_low = (in.g0_int[bi] + in.g0_lay[b]) / 2.0;
        }
        double ray_up = 0.0, ray_low = 0.0, csc_up = 0.0, csc_low = 0.0;
        if (s.scat == 1) {
            ray_up = (in.scat_lay[b] + in.scat_int[bu]) / 2.0;
            ray_low = (in.scat_int[bi] + in.scat_lay[b]) / 2.0;
            csc_up = (in.csc_lay[b] + in.csc_int[bu]) / 2.0;
            csc_low = (in.csc_int[bi] + in.csc_lay[b]) / 2.0;
        }
        const double cab_up = (in.cab_lay[b] + in.cab_int[bu]) / 2.0;
        const double cab_low = (in.cab_int[bi] + in.cab_lay[b]) / 2.0;
        const double k_lay = in.opac_lay[e];
        const double opac_up = (k_lay + in.opac_int[e + ncol]) / 2.0;
        const double opac_low = (in.opac_int[e] + k_lay) / 2.0;
        const double mmm_up = (in.mmm_lay[li] + in.mmm_int[ii + 1]) / 2.0;
        const double mmm_low = (in.mmm_int[ii] + in.mmm_lay[li]) / 2.0;
        const double dtc_up = in.dcol_u[li] * (cab_up + csc_up) / mmm_up;
        const double dtc_low = in.dcol_l[li] * (cab_low + csc_low) / mmm_low;
        if (y == 0) {
            o.dtc_u[b] = dtc_up;
            o.dtc_l[b] = dtc_low;
        }
        const CellCoeffs u = cell_coeffs(ray_up, csc_up, cab_up, opac_up, mmm_up, in.dcol_u[li], dtc_up, g0_up, s);
        const CellCoeffs l = cell_coeffs(ray_low, csc_low, cab_low, opac_low, mmm_low, in.dcol_l[li], dtc_low, g0_low, s);
        o.w0_u[e] = u.w0;      o.w0_l[e] = l.w0;
        o.dtau_u[e] = u.dtau;  o.dtau_l[e] = l.dtau;
        o.trans_u[e] = u.trans; o.trans_l[e] = l.trans;
        o.M_u[e] = u.M;  o.M_l[e] = l.M;
        o.N_u[e] = u.N;  o.N_l[e] = l.N;
        o.P_u[e] = u.P;  o.P_l[e] = l.P;
        o.Gp_u[e] = u.Gp; o.Gp_l[e] = l.Gp;
        o.Gm_u[e] = u.Gm; o.Gm_l[e] = l.Gm;
        if (u.w0 > s.w_0_scat_limit || l.w0 > s.w_0_scat_limit) scat_trigger[(size_t)a * ncol + col] = 1;
    }
}

// K:1247-1261
__global__ void k_calc_delta_z(const double* __restrict__ tlay, const double* __restrict__ pint,
                               const double* __restrict__ mmm, double* __restrict__ dz, double g, int nlayer,
                               const double* __restrict__ g_batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t a = blockIdx.y;  // batch: T_lay and p_int hold nlayer + 1 values per atmosphere
    if (g_batch) g = g_batch[a];
    tlay += a * (nlayer + 1);
    pint += a * (nlayer + 1);
    mmm += a * nlayer;
    dz += a * nlayer;
    if (i < nlayer) dz[i] = hc::KBOLTZMANN * tlay[i] / (mmm[i] * g) * log(pint[i] / pint[i + 1]);
}

// ------------------------------------------------------------------------------------------
// Direct beam (K:1265-1362).  One thread per column walks down from TOA.
// Without the geometric zenith correction the attenuation is a running product
// F[i] = F[i+1] * exp(dtau_i / mu*), which is exactly the reference's multiplication order, done in O(n)
// instead of the reference's O(n^2) re-multiplication per interface.  With the correction mu depends
// on (i, j) and the product is rebuilt per interface as in the reference.
// NONISO additionally produces the layer-centre value Fc_dir (K:1358).
// ------------------------------------------------------------------------------------------
template <bool NONISO>
__global__ void __launch_bounds__(128)
k_fdir(double* __restrict__ F_dir, double* __restrict__ Fc_dir, const double* __restrict__ planck_lay,
       const double* __restrict__ dtau_a,  // iso: delta_tau_wg ; noniso: upper
       const double* __restrict__ dtau_b,  // noniso: lower
       const double* __restrict__ z_lay, double mu_star, double R_planet, double R_star, double a,
       int dir_beam, int geom, int nint, int nbin, int ny) {
    const int ncol = nbin * ny;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int x = col / ny;
    const int nlay = nint - 1;
    const double I_dir = ((R_star / a) * (R_star / a)) * hc::PI * planck_lay[nlay + (size_t)x * (nlay + 2)];
    const double F_toa = -dir_beam * mu_star * I_dir;
    F_dir[col + (size_t)ncol * nlay] = F_toa;
    if (geom != 1) {
        double F = F_toa;
        for (int i = nlay - 1; i >= 0; i--) {
            const size_t e = col + (size_t)ncol * i;
            if (NONISO) {
                const double du = dtau_a[e];
                Fc_dir[e] = F * exp(du / mu_star);
                F *= exp((du + dtau_b[e]) / mu_star);
            } else {
                F *= exp(dtau_a[e] / mu_star);
            }
            F_dir[e] = F;
        }
    } else {
        const double one_m_mu2 = 1.0 - mu_star * mu_star;
        for (int i = nlay - 1; i >= 0; i--) {
            const double ri = R_planet + z_lay[i];
            double F = F_toa, Fc = 0.0;
            for (int j = nlay - 1; j >= i; j--) {
                const double q = ri / (R_planet + z_lay[j]);
                const double mu_j = -sqrt(1.0 - (q * q) * one_m_mu2);
                const size_t e = col + (size_t)ncol * j;
                if (NONISO) {
                    const double du = dtau_a[e];
                    Fc = F * exp(du / mu_j);
                    F *= exp((du + dtau_b[e]) / mu_j);
                } else {
                    F *= exp(dtau_a[e] / mu_j);
                }
            }
            const size_t e = col + (size_t)ncol * i;
            F_dir[e] = F;
            if (NONISO) Fc_dir[e] = Fc;
        }
    }
}

// Layer-parallel form of the non-geometric branch above.  The attenuation factors exp(dtau/mu*) of a
// column do not depend on each other, only their running product does: a block takes FD_COLS columns, all of
// its threads evaluate the exponentials of every (layer, column) cell in parallel into shared memory, one warp
// then forms the running products in the same top-down order as k_fdir (bit-identical results), and all
// threads store the block's [layer][column] slab coalesced.  ~100 dependent multiplies per column instead of
// ~100 dependent (load -> exp -> multiply) round trips.
#define FD_COLS 32
#define FD_THREADS 256
template <bool NONISO>
__global__ void __launch_bounds__(FD_THREADS)
k_fdir_lp(double* __restrict__ F_dir, double* __restrict__ Fc_dir, const double* __restrict__ planck_lay,
          const double* __restrict__ dtau_a, const double* __restrict__ dtau_b, double mu_star, double R_star,
          double a, int dir_beam, int nint, int nbin, int ny) {
    extern __shared__ double fd_sm[];
    const int ncol = nbin * ny;
    const int nlay = nint - 1;
    double* s_full = fd_sm;                                   // [nlay][FD_COLS]: exp of the whole layer -> F_dir
    double* s_half = fd_sm + (size_t)nlay * FD_COLS;          // NONISO: exp of the upper half -> Fc_dir
    const int c = threadIdx.x % FD_COLS;
    const int r = threadIdx.x / FD_COLS;
    constexpr int ROWS = FD_THREADS / FD_COLS;
    const int col = blockIdx.x * FD_COLS + c;
    const bool live = col < ncol;
    {   // batch (blockIdx.y = atmosphere)
        const size_t a = blockIdx.y;
        const size_t wg = (size_t)ncol * nint;
        F_dir += a * wg;
        dtau_a += a * wg;
        if (NONISO) {
            Fc_dir += a * wg;
            dtau_b += a * wg;
        }
        planck_lay += a * (size_t)(nlay + 2) * nbin;
    }
    if (live) {
        for (int i = r; i < nlay; i += ROWS) {
            const size_t e = col + (size_t)ncol * i;
            if (NONISO) {
                const double du = dtau_a[e];
                s_half[i * FD_COLS + c] = exp(du / mu_star);
                s_full[i * FD_COLS + c] = exp((du + dtau_b[e]) / mu_star);
            } else {
                s_full[i * FD_COLS + c] = exp(dtau_a[e] / mu_star);
            }
        }
    }
    __syncthreads();
    if (r == 0 && live) {
        const int x = col / ny;
        const double I_dir = ((R_star / a) * (R_star / a)) * hc::PI * planck_lay[nlay + (size_t)x * (nlay + 2)];
        double F = -dir_beam * mu_star * I_dir;
        F_dir[col + (size_t)ncol * nlay] = F;
        for (int i = nlay - 1; i >= 0; i--) {
            if (NONISO) s_half[i * FD_COLS + c] = F * s_half[i * FD_COLS + c];
            F *= s_full[i * FD_COLS + c];
            s_full[i * FD_COLS + c] = F;
        }
    }
    __syncthreads();
    if (live) {
        for (int i = r; i < nlay; i += ROWS) {
            const size_t e = col + (size_t)ncol * i;
            F_dir[e] = s_full[i * FD_COLS + c];
            if (NONISO) Fc_dir[e] = s_half[i * FD_COLS + c];
        }
    }
}

template <bool NONISO>
static int launch_fdir(helios_ctx* ctx, double* F_dir, double* Fc_dir, const double* planck_lay,
                       const double* dtau_a, const double* dtau_b, const double* z_lay, double mu_star,
                       double R_planet, double R_star, double a, int dir_beam, int geom, int nint, int nbin,
                       int ny) {
    const int ncol = nbin * ny;
    const size_t smem = (size_t)(nint - 1) * FD_COLS * sizeof(double) * (NONISO ? 2 : 1);
    HBATCHDIMS(ctx, nint == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    if (geom != 1 && smem <= 200 * 1024) {
        auto kern = k_fdir_lp<NONISO>;
        HCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(ceil_div(ncol, FD_COLS), ctx->batch.nbatch), FD_THREADS, smem, ctx->stream>>>(
            F_dir, Fc_dir, planck_lay, dtau_a, dtau_b, mu_star, R_star, a, dir_beam, nint, nbin, ny);
    } else {
        HNOBATCH(ctx);
        k_fdir<NONISO><<<ceil_div(ncol, 128), 128, 0, ctx->stream>>>(F_dir, Fc_dir, planck_lay, dtau_a, dtau_b, z_lay,
                                                                     mu_star, R_planet, R_star, a, dir_beam, geom,
                                                                     nint, nbin, ny);
    }
    HLAUNCHED(ctx);
    // with dir_beam == 0 every entry written above is -0.0 (F_toa = -0 * mu* * I): remember that (common.cuh)
    ctx->zero_beam[0] = dir_beam == 0 ? F_dir : nullptr;
    ctx->zero_beam[1] = (dir_beam == 0 && NONISO) ? Fc_dir : nullptr;
    ctx->zero_beam_bytes = (size_t)ncol * nint * sizeof(double) * ctx->batch.nbatch;
    return HELIOS_OK;
}

extern "C" {

int helios_calc_trans_iso(helios_ctx* ctx, double* trans_wg, double* delta_tau_wg, double* M_term,
                          double* N_term, double* P_term, double* G_plus, double* G_minus,
                          const double* delta_colmass, const double* opac_wg_lay,
                          const double* meanmolmass_lay, const double* scat_cross_lay,
                          const double* abs_cross_all_clouds_lay, const double* scat_cross_all_clouds_lay,
                          double* delta_tau_all_clouds, double* w_0, const double* g_0_tot_lay,
                          int* scat_trigger, double g_0, double epsi, double epsi2, double mu_star,
                          double w_0_limit, double w_0_scat_limit, int scat, int nbin, int ny, int nlayer,
                          int clouds, int scat_corr, int debug, double i2s_transition) {
    HCTX(ctx);
    HARG(trans_wg && delta_tau_wg && M_term && N_term && P_term && G_plus && G_minus && delta_colmass &&
         opac_wg_lay && meanmolmass_lay && scat_cross_lay && abs_cross_all_clouds_lay &&
         scat_cross_all_clouds_lay && delta_tau_all_clouds && w_0 && scat_trigger);
    HARG(clouds == 0 || g_0_tot_lay != nullptr);
    HARG(nbin > 0 && ny > 0 && nlayer > 0);
    TransScalars s{g_0, epsi, epsi2, mu_star, w_0_limit, w_0_scat_limit, i2s_transition,
                   scat, nbin, ny, nlayer, clouds, scat_corr, debug, 0.0, 0.0, ctx->batch.nbatch};
    HBATCHDIMS(ctx, nbin == ctx->batch.nbin && ny == ctx->batch.ny && nlayer == ctx->batch.nlayer);
    {
        const int rc = trans_constants(ctx, s);
        if (rc != HELIOS_OK) return rc;
    }
    const long long total = (long long)nbin * ny * nlayer * ctx->batch.nbatch;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    k_calc_trans_iso<<<(int)blocks, 256, 0, ctx->stream>>>(
        trans_wg, delta_tau_wg, M_term, N_term, P_term, G_plus, G_minus, delta_colmass, opac_wg_lay,
        meanmolmass_lay, scat_cross_lay, abs_cross_all_clouds_lay, scat_cross_all_clouds_lay,
        delta_tau_all_clouds, w_0, g_0_tot_lay, scat_trigger, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_trans_noniso(
    helios_ctx* ctx, double* trans_wg_upper, double* trans_wg_lower, double* delta_tau_wg_upper,
    double* delta_tau_wg_lower, double* M_upper, double* M_lower, double* N_upper, double* N_lower,
    double* P_upper, double* P_lower, double* G_plus_upper, double* G_plus_lower, double* G_minus_upper,
    double* G_minus_lower, const double* delta_col_upper, const double* delta_col_lower,
    const double* opac_wg_lay, const double* opac_wg_int, const double* meanmolmass_lay,
    const double* meanmolmass_int, const double* scat_cross_lay, const double* scat_cross_int,
    const double* abs_cross_all_clouds_lay, const double* abs_cross_all_clouds_int,
    const double* scat_cross_all_clouds_lay, const double* scat_cross_all_clouds_int,
    double* delta_tau_all_clouds_upper, double* delta_tau_all_clouds_lower, double* w_0_upper,
    double* w_0_lower, const double* g_0_tot_lay, const double* g_0_tot_int, int* scat_trigger,
    double g_0, double epsi, double epsi2, double mu_star, double w_0_limit, double w_0_scat_limit,
    int scat, int nbin, int ny, int nlayer, int clouds, int scat_corr, int debug,
    double i2s_transition) {
    HCTX(ctx);
    HARG(trans_wg_upper && trans_wg_lower && delta_tau_wg_upper && delta_tau_wg_lower && M_upper &&
         M_lower && N_upper && N_lower && P_upper && P_lower && G_plus_upper && G_plus_lower &&
         G_minus_upper && G_minus_lower && delta_col_upper && delta_col_lower && opac_wg_lay &&
         opac_wg_int && meanmolmass_lay && meanmolmass_int && scat_cross_lay && scat_cross_int &&
         abs_cross_all_clouds_lay && abs_cross_all_clouds_int && scat_cross_all_clouds_lay &&
         scat_cross_all_clouds_int && delta_tau_all_clouds_upper && delta_tau_all_clouds_lower &&
         w_0_upper && w_0_lower && scat_trigger);
    HARG(clouds == 0 || (g_0_tot_lay != nullptr && g_0_tot_int != nullptr));
    HARG(nbin > 0 && ny > 0 && nlayer > 0);
    TransScalars s{g_0, epsi, epsi2, mu_star, w_0_limit, w_0_scat_limit, i2s_transition,
                   scat, nbin, ny, nlayer, clouds, scat_corr, debug, 0.0, 0.0, ctx->batch.nbatch};
    HBATCHDIMS(ctx, nbin == ctx->batch.nbin && ny == ctx->batch.ny && nlayer == ctx->batch.nlayer);
    {
        const int rc = trans_constants(ctx, s);
        if (rc != HELIOS_OK) return rc;
    }
    NonisoOut o{trans_wg_upper, trans_wg_lower, delta_tau_wg_upper, delta_tau_wg_lower, M_upper, M_lower,
                N_upper, N_lower, P_upper, P_lower, G_plus_upper, G_plus_lower, G_minus_upper,
                G_minus_lower, delta_tau_all_clouds_upper, delta_tau_all_clouds_lower, w_0_upper,
                w_0_lower};
    NonisoIn in{delta_col_upper, delta_col_lower, opac_wg_lay, opac_wg_int, meanmolmass_lay,
                meanmolmass_int, scat_cross_lay, scat_cross_int, abs_cross_all_clouds_lay,
                abs_cross_all_clouds_int, scat_cross_all_clouds_lay, scat_cross_all_clouds_int,
                g_0_tot_lay, g_0_tot_int};
    const long long total = (long long)nbin * ny * nlayer * ctx->batch.nbatch;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    k_calc_trans_noniso<<<(int)blocks, 256, 0, ctx->stream>>>(o, in, scat_trigger, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_delta_z(helios_ctx* ctx, const double* tlay, const double* pint, const double* play,
                        const double* meanmolmass_lay, double* delta_z_lay, double g, int nlayer) {
    HCTX(ctx);
    (void)play;
    HARG(tlay && pint && meanmolmass_lay && delta_z_lay && nlayer > 0);
    HBATCHDIMS(ctx, nlayer == ctx->batch.nlayer);
    k_calc_delta_z<<<dim3(ceil_div(nlayer, 128), ctx->batch.nbatch), 128, 0, ctx->stream>>>(
        tlay, pint, meanmolmass_lay, delta_z_lay, g, nlayer, ctx->batch.active ? ctx->batch.g : nullptr);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_fdir_iso(helios_ctx* ctx, double* F_dir_wg, const double* planckband_lay,
                    const double* delta_tau_wg, const double* z_lay, double mu_star, double R_planet,
                    double R_star, double a, int dir_beam, int geom_zenith_corr, int ninterface, int nbin,
                    int ny) {
    HCTX(ctx);
    HARG(F_dir_wg && planckband_lay && delta_tau_wg && ninterface > 1 && nbin > 0 && ny > 0);
    HARG(geom_zenith_corr != 1 || z_lay != nullptr);
    return launch_fdir<false>(ctx, F_dir_wg, nullptr, planckband_lay, delta_tau_wg, nullptr, z_lay, mu_star,
                              R_planet, R_star, a, dir_beam, geom_zenith_corr, ninterface, nbin, ny);
}

int helios_fdir_noniso(helios_ctx* ctx, double* F_dir_wg, double* Fc_dir_wg, const double* planckband_lay,
                       const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
                       const double* z_lay, double mu_star, double R_planet, double R_star, double a,
                       int dir_beam, int geom_zenith_corr, int ninterface, int nbin, int ny) {
    HCTX(ctx);
    HARG(F_dir_wg && Fc_dir_wg && planckband_lay && delta_tau_wg_upper && delta_tau_wg_lower &&
         ninterface > 1 && nbin > 0 && ny > 0);
    HARG(geom_zenith_corr != 1 || z_lay != nullptr);
    return launch_fdir<true>(ctx, F_dir_wg, Fc_dir_wg, planckband_lay, delta_tau_wg_upper, delta_tau_wg_lower,
                             z_lay, mu_star, R_planet, R_star, a, dir_beam, geom_zenith_corr, ninterface, nbin, ny);
}

}  // extern "C"
