// Convective adjustment and on-the-fly VMR interpolation on the device.
//
// The reference does both on the host every iteration of the radiative-convective loop: ~12 host<->device round trips
// per iteration (C:1053-1139) around host_functions.py:337-635 (conv_check, conv_correct, mark_convective_layers,
// stitching, check_for_radiative_eq), and per refresh a scipy RectBivariateSpline per layer and species (H:874-959).
// The algorithms are serial over ~100 layers and tiny; what costs is the ping-pong.  Here each is ONE small launch:
//   k_convective_adjustment   one block per atmosphere: the adiabat factors of every layer (the only transcendental work,
//                             six pow() per layer, T-independent) are formed in parallel; thread 0 then runs the
//                             check / mark / correct cycle to stability and the final damped correction, exactly the
//                             control flow of H:509-538 (restated in helios_b200/host.py, which is pinned to the reference
//                             module by tests/golden/host_golden.npz).
//   k_convection_marks        mark_convective_layers(stitching = 1) + check_for_radiative_eq (H:545-582, 251-286) after
//                             the flux solve: conv_layer, marked_red and three counters for the loop's exit test.
//   k_vmr_interpol            bilinear (T, log10 P) interpolation of a species' pre-tabulated VMR (RectBivariateSpline with
//                             kx = ky = 1 is piecewise bilinear, clamped at the grid edges) and the VMR-weighted mean
//                             molecular mass accumulation (H:927-959).
#include "common.cuh"

namespace {

constexpr int CV_THREADS = 128;
constexpr int CV_MAXL = 1024;  // layers

struct ConvArgs {
    double* T_lay;                // [n + 1] in / out (index n = surface)
    const double *p_lay, *p_int, *kappa_lay, *kappa_int, *c_p_lay, *mmm_lay;
    const double *F_add_heat_sum, *F_smooth_sum, *F_down_tot, *F_up_tot;
    int* conv_layer;              // [n + 1] in / out
    int* conv_unstable;           // [n + 1] out
    int* status;                  // [4] out: adjustment cycles, zones, unstable flags at entry, error
    double F_intern, T_star, dampara;  // dampara <= 0: "automatic" (H:441-449)
    int n, iter_value;
};

// adiabat factors of layer i, exponents stretched by `slack` (H:348-350, 557-559, 470-478):
//   up[i]  = (p_int[i+1] / p_lay[i]) ^ (kappa_lay[i] slack)      layer centre  -> interface above
//   cen[i] = (p_lay[i]   / p_int[i]) ^ (kappa_int[i] slack)      interface i   -> layer centre
struct Factors {
    double* up[3];   // slack 1 + 1e-6, 1 - 1e-6, 1
    double* cen[3];
};

__device__ void mark_layers(const ConvArgs& a, const Factors& f, int* mark, int stitching) {
    const int n = a.n;
    const double* T = a.T_lay;
    mark[n] = 0;
    mark[0] = 0;
    for (int i = 0; i < n - 1; i++) {
        if (a.p_lay[i] <= 1e1) break;  // the uppermost atmosphere keeps its previous marks
        const double Tad = (T[i] * f.up[1][i]) * f.cen[1][i + 1];
        if (T[i + 1] < Tad) {
            mark[i] = mark[i + 1] = 1;
        } else {
            mark[i + 1] = 0;
        }
    }
    for (int i = 0; i < n - 1; i++)  // no kink at the top edge of a zone
        if (T[i + 1] > T[i]) mark[i] = 0;
    if (T[0] < T[n] * f.cen[1][0]) mark[n] = mark[0] = 1;
    if (stitching == 1 && a.iter_value > 5000) {
        // close radiative gaps thinner than a scale height between convective zones (H:585-635)
        int prev_end = -2;   // last layer of the previous zone (-1 = the surface "layer"), -2 = none yet
        bool in_zone = mark[n] == 1 && mark[0] == 0;  // a zone consisting of the surface alone
        if (in_zone) prev_end = -1;
        int i = 0;
        while (i < n) {
            if (mark[i] != 1) { i++; continue; }
            const int start = i;
            const int below = i > 0 ? mark[i - 1] : mark[n];
            int end = i;
            while (end + 1 < n && mark[end + 1] == 1) end++;
            if (below == 0 && prev_end != -2) {
                const double p_top = a.p_lay[start];
                const double p_bot = prev_end != -1 ? a.p_lay[prev_end] : a.p_int[0];
                if (p_top / p_bot > 1.0 / 2.718281828459045)
                    for (int m = prev_end + 1; m < start; m++) mark[m] = 1;
            }
            prev_end = end;
            i = end + 1;
        }
    }
}

__device__ int check_layers(const ConvArgs& a, const Factors& f, int* flag) {
    const int n = a.n;
    const double* T = a.T_lay;
    int any = 0;
    for (int i = 0; i <= n; i++) flag[i] = 0;
    for (int i = 0; i < n - 1; i++) {
        if (a.p_lay[i] <= 1e1) break;
        const double Tad = (T[i] * f.up[0][i]) * f.cen[0][i + 1];
        if (T[i + 1] < Tad) {
            flag[i] = flag[i + 1] = 1;
            any = 1;
        }
    }
    if (T[0] < T[n] * f.cen[0][0]) {
        flag[n] = flag[0] = 1;
        any = 1;
    }
    return any;
}

// replaces unstable lapse rates by adiabats that conserve the zone's enthalpy (H:368-506)
__device__ int correct_layers(const ConvArgs& a, const Factors& f, const int* unstable, const int* mark, int fudging,
                              int* z_start, int* z_end) {
    const int n = a.n;
    double* T = a.T_lay;
    // zones: maximal runs of layers that are unstable or marked; the surface (index n) sits below layer 0 as "-1"
    int nz = 0;
    const bool surf = unstable[n] == 1 || mark[n] == 1;
    int prev = surf ? -1 : -3;
    if (surf) { z_start[0] = -1; z_end[0] = -1; nz = 1; }
    for (int i = 0; i < n; i++) {
        if (!(unstable[i] == 1 || mark[i] == 1)) continue;
        if (nz > 0 && prev == i - 1) {
            z_end[nz - 1] = i;
        } else {
            z_start[nz] = i;
            z_end[nz] = i;
            nz++;
        }
        prev = i;
    }
    for (int z = 0; z < nz; z++) {
        double fudge = 1.0;
        if (fudging == 1) {
            int probe = 0;
            bool found = false;
            for (int m = z; m < nz && !found; m++) {
                if (m != nz - 1) {
                    const double p_top = a.p_lay[z_start[m + 1]];
                    const double p_bot = z_end[m] != -1 ? a.p_lay[z_end[m]] : a.p_int[0];
                    if (p_top / p_bot < 1.0 / 2.718281828459045) {  // a radiative zone thicker than a scale height follows
                        probe = (z_end[m] + z_start[m + 1]) / 2;
                        found = true;
                    }
                } else {
                    probe = (int)(0.8 * z_end[m] + 0.2 * n);  // ninterface - 1 == n
                    found = true;
                }
            }
            double dampara = a.dampara;
            if (dampara <= 0.0) dampara = a.T_star > 10.0 ? (z < nz - 1 ? 0.5 : 4.0) : 8.0;
            // Python's negative index: probe - 1 == -1 addresses the last element (H:452)
            const int pm = probe - 1 >= 0 ? probe - 1 : n - 1;
            const double r = (a.F_intern + a.F_add_heat_sum[pm] + a.F_smooth_sum[pm] + a.F_down_tot[probe]) / a.F_up_tot[probe];
            fudge = fmin(1.01, fmax(0.99, pow(r, 1.0 / dampara)));
        }
        const int lo = max(0, z_start[z]), hi = max(0, z_end[z]);
        double num = 0.0, den = 0.0, climb = 1.0;
        for (int i = lo; i <= hi; i++) {
            const double weight = a.c_p_lay[i] / a.mmm_lay[i];
            const double dp = a.p_int[i] - a.p_int[i + 1];
            num += weight * T[i] * dp;
            den += climb * (f.cen[2][i] * a.c_p_lay[i] / a.mmm_lay[i] * dp);
            climb = climb * (f.cen[2][i] * f.up[2][i]);
        }
        double theta = num / den;
        theta *= fudge;
        climb = 1.0;
        for (int i = lo; i <= hi; i++) {
            T[i] = theta * (climb * f.cen[2][i]);
            climb = climb * (f.cen[2][i] * f.up[2][i]);
        }
        if (z_start[z] == -1) T[n] = theta;
    }
    return nz;
}

__global__ void __launch_bounds__(CV_THREADS) k_convective_adjustment(ConvArgs a, int nbatch_stride_n) {
    extern __shared__ double sm[];
    const int n = a.n;
    if (blockIdx.x > 0) {  // batch: atmosphere b's vectors follow each other with the reference's allocation sizes
        const size_t v1 = (size_t)blockIdx.x * (n + 1), v0 = (size_t)blockIdx.x * n;
        a.T_lay += v1; a.p_int += v1; a.kappa_int += v1; a.F_down_tot += v1; a.F_up_tot += v1;
        a.conv_layer += v1; a.conv_unstable += v1;
        a.p_lay += v0; a.kappa_lay += v0; a.c_p_lay += v0; a.mmm_lay += v0; a.F_add_heat_sum += v0; a.F_smooth_sum += v0;
        a.status += 4 * blockIdx.x;
    }
    (void)nbatch_stride_n;
    Factors f;
    for (int k = 0; k < 3; k++) {
        f.up[k] = sm + (size_t)(2 * k) * n;
        f.cen[k] = sm + (size_t)(2 * k + 1) * n;
    }
    int* z_start = reinterpret_cast<int*>(sm + 6 * (size_t)n);
    int* z_end = z_start + (n + 2);
    const double slack[3] = {1.0 + 1e-6, 1.0 - 1e-6, 1.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double xu = a.p_int[i + 1] / a.p_lay[i], xc = a.p_lay[i] / a.p_int[i];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            f.up[k][i] = pow(xu, a.kappa_lay[i] * slack[k]);
            f.cen[k][i] = pow(xc, a.kappa_int[i] * slack[k]);
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int cycles = 0, zones = 0;
    int any = check_layers(a, f, a.conv_unstable);
    const int unstable_at_entry = any;
    while (any && cycles < 100000) {
        mark_layers(a, f, a.conv_layer, 0);
        correct_layers(a, f, a.conv_unstable, a.conv_layer, 0, z_start, z_end);
        any = check_layers(a, f, a.conv_unstable);
        cycles++;
    }
    mark_layers(a, f, a.conv_layer, 1);
    zones = correct_layers(a, f, a.conv_unstable, a.conv_layer, 1, z_start, z_end);
    a.status[0] = cycles;
    a.status[1] = zones;
    a.status[2] = unstable_at_entry;
    a.status[3] = any ? 1 : 0;  // 1: did not reach stability within the cycle limit
}

struct MarkArgs {
    const double *T_lay, *p_lay, *p_int, *kappa_lay, *kappa_int;
    const double *F_net, *F_down_tot, *F_add_heat_sum, *F_smooth_sum;
    int *conv_layer, *marked_red;
    int* status;  // [4] out: converged radiative layers, radiative layers, convective layers, zero-temperature layers
    double F_intern, limit;
    int n, iter_value;
};

__global__ void __launch_bounds__(CV_THREADS) k_convection_marks(MarkArgs m) {
    extern __shared__ double sm[];
    const int n = m.n;
    if (blockIdx.x > 0) {
        const size_t v1 = (size_t)blockIdx.x * (n + 1), v0 = (size_t)blockIdx.x * n;
        m.T_lay += v1; m.p_int += v1; m.kappa_int += v1; m.F_net += v1; m.F_down_tot += v1;
        m.conv_layer += v1; m.marked_red += v1;
        m.p_lay += v0; m.kappa_lay += v0; m.F_add_heat_sum += v0; m.F_smooth_sum += v0;
        m.status += 4 * blockIdx.x;
    }
    Factors f;
    for (int k = 0; k < 3; k++) {
        f.up[k] = sm;            // only the (1 - 1e-6) pair is used here
        f.cen[k] = sm + n;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        f.up[1][i] = pow(m.p_int[i + 1] / m.p_lay[i], m.kappa_lay[i] * (1.0 - 1e-6));
        f.cen[1][i] = pow(m.p_lay[i] / m.p_int[i], m.kappa_int[i] * (1.0 - 1e-6));
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    ConvArgs a{};
    a.T_lay = const_cast<double*>(m.T_lay);
    a.p_lay = m.p_lay;
    a.p_int = m.p_int;
    a.n = n;
    a.iter_value = m.iter_value;
    mark_layers(a, f, m.conv_layer, 1);
    // local radiative equilibrium over the radiative layers (H:251-286)
    const double scale = m.limit * (m.F_down_tot[n] + m.F_intern);
    int converged = 0, nconv = 0, zero_T = 0;
    for (int i = 0; i <= n; i++) {
        m.marked_red[i] = 0;
        if (m.T_lay[i] == 0.0) zero_T++;
        if (m.conv_layer[i] != 0) {
            nconv++;
            continue;
        }
        const double miss = i < n ? fabs(m.F_intern + m.F_add_heat_sum[i] + m.F_smooth_sum[i] - m.F_net[i + 1])
                                  : fabs(m.F_intern - m.F_net[0]);
        if (miss < scale) converged++;
        else m.marked_red[i] = 1;
    }
    m.status[0] = converged;
    m.status[1] = (n + 1) - nconv;
    m.status[2] = nconv;
    m.status[3] = zero_T;
}

// bilinear interpolation in (T, log10 P), clamped at the grid edges; optional accumulation of the mean molecular mass
__global__ void k_vmr_interpol(const double* __restrict__ T, const double* __restrict__ P, const double* __restrict__ ktemp,
                               const double* __restrict__ kpress, const double* __restrict__ table, double* __restrict__ vmr,
                               int ntemp, int npress, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t = fmin(fmax(T[i], ktemp[0]), ktemp[ntemp - 1]);
    const double lp0 = log10(kpress[0]), lp1 = log10(kpress[npress - 1]);
    const double p = fmin(fmax(log10(P[i]), lp0), lp1);
    // searchsorted(side = "right") - 1, clipped to [0, size - 2]
    int it = 0, ip = 0;
    {
        int lo = 0, hi = ntemp;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (ktemp[mid] <= t) lo = mid + 1; else hi = mid; }
        it = min(max(lo - 1, 0), ntemp - 2);
        lo = 0; hi = npress;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (log10(kpress[mid]) <= p) lo = mid + 1; else hi = mid; }
        ip = min(max(lo - 1, 0), npress - 2);
    }
    const double lpa = log10(kpress[ip]), lpb = log10(kpress[ip + 1]);
    const double ft = (t - ktemp[it]) / (ktemp[it + 1] - ktemp[it]);
    const double fp = (p - lpa) / (lpb - lpa);
    const double v00 = table[(size_t)it * npress + ip], v10 = table[(size_t)(it + 1) * npress + ip];
    const double v01 = table[(size_t)it * npress + ip + 1], v11 = table[(size_t)(it + 1) * npress + ip + 1];
    vmr[i] = v00 * (1 - ft) * (1 - fp) + v10 * ft * (1 - fp) + v01 * (1 - ft) * fp + v11 * ft * fp;
}

__global__ void k_mmm_accumulate(const double* __restrict__ vmr, double weight, double* __restrict__ sum_w,
                                 double* __restrict__ sum_v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum_w[i] += vmr[i] * weight;
    sum_v[i] += vmr[i];
}

__global__ void k_mmm_finish(const double* __restrict__ sum_w, const double* __restrict__ sum_v, double* __restrict__ mmm,
                             int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mmm[i] = sum_w[i] / sum_v[i] * hc::AMU;
}

}  // namespace

extern "C" {

int helios_convective_adjustment(helios_ctx* ctx, double* T_lay, const double* p_lay, const double* p_int,
                                 const double* kappa_lay, const double* kappa_int, const double* c_p_lay,
                                 const double* meanmolmass_lay, const double* F_add_heat_sum, const double* F_smooth_sum,
                                 const double* F_down_tot, const double* F_up_tot, int* conv_layer, int* conv_unstable,
                                 int* status_dev, double F_intern, double T_star, double dampara, int iter_value,
                                 int nlayer) {
    HCTX(ctx);
    HARG(T_lay && p_lay && p_int && kappa_lay && kappa_int && c_p_lay && meanmolmass_lay && F_add_heat_sum && F_smooth_sum &&
         F_down_tot && F_up_tot && conv_layer && conv_unstable && status_dev);
    HARG(nlayer >= 2 && nlayer <= CV_MAXL);
    HBATCHDIMS(ctx, nlayer == ctx->batch.nlayer);
    ConvArgs a{T_lay, p_lay, p_int, kappa_lay, kappa_int, c_p_lay, meanmolmass_lay, F_add_heat_sum, F_smooth_sum, F_down_tot,
               F_up_tot, conv_layer, conv_unstable, status_dev, F_intern, T_star, dampara, nlayer, iter_value};
    const size_t smem = (size_t)6 * nlayer * sizeof(double) + (size_t)2 * (nlayer + 2) * sizeof(int);
    k_convective_adjustment<<<ctx->batch.nbatch, CV_THREADS, smem, ctx->stream>>>(a, nlayer);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_convection_marks(helios_ctx* ctx, const double* T_lay, const double* p_lay, const double* p_int,
                            const double* kappa_lay, const double* kappa_int, const double* F_net,
                            const double* F_down_tot, const double* F_add_heat_sum, const double* F_smooth_sum,
                            int* conv_layer, int* marked_red, int* status_dev, double F_intern,
                            double rad_convergence_limit, int iter_value, int nlayer) {
    HCTX(ctx);
    HARG(T_lay && p_lay && p_int && kappa_lay && kappa_int && F_net && F_down_tot && F_add_heat_sum && F_smooth_sum &&
         conv_layer && marked_red && status_dev);
    HARG(nlayer >= 2 && nlayer <= CV_MAXL);
    HBATCHDIMS(ctx, nlayer == ctx->batch.nlayer);
    MarkArgs m{T_lay, p_lay, p_int, kappa_lay, kappa_int, F_net, F_down_tot, F_add_heat_sum, F_smooth_sum, conv_layer,
               marked_red, status_dev, F_intern, rad_convergence_limit, nlayer, iter_value};
    k_convection_marks<<<ctx->batch.nbatch, CV_THREADS, (size_t)2 * nlayer * sizeof(double), ctx->stream>>>(m);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_vmr_interpol(helios_ctx* ctx, const double* temp, const double* press, const double* ktemp,
                        const double* kpress, const double* vmr_pretab, double* vmr_out, int npress, int ntemp, int n) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(temp && press && ktemp && kpress && vmr_pretab && vmr_out && npress >= 2 && ntemp >= 2 && n > 0);
    k_vmr_interpol<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(temp, press, ktemp, kpress, vmr_pretab, vmr_out, ntemp, npress, n);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_meanmolmass_accumulate(helios_ctx* ctx, const double* vmr, double weight, double* sum_weighted, double* sum_vmr,
                                  int n) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(vmr && sum_weighted && sum_vmr && n > 0);
    k_mmm_accumulate<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(vmr, weight, sum_weighted, sum_vmr, n);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_meanmolmass_finish(helios_ctx* ctx, const double* sum_weighted, const double* sum_vmr, double* meanmolmass, int n) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(sum_weighted && sum_vmr && meanmolmass && n > 0);
    k_mmm_finish<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(sum_weighted, sum_vmr, meanmolmass, n);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

}  // extern "C"
