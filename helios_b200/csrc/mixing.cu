// On-the-fly species mixing: correlated-k summation and random overlap (RO) with a warp-level
// bitonic sort in shared memory, plus the small scattering helpers.
// From-scratch sm_100a kernels for K:472-492 and K:3143-3459 of the reference.
#include "common.cuh"
#include <cfloat>

#define RO_NY 20
#define RO_N2 400     // ny*ny k-combinations per cell (K:3315)
#define RO_PAD 512    // bitonic network size
#define RO_WARPS 4

// compare-exchange building blocks of the warp-level bitonic network (element e = PER * lane + r in register r)
__device__ __forceinline__ bool ro_greater(double a, int ta, double b, int tb) {
    return (a > b) | ((a == b) & ((ta & 511) > (tb & 511)));  // total order on (key, slot); no short-circuit branches
}

template <int PER, int J>
__device__ __forceinline__ void ro_in_lane(double (&kv)[PER], int (&tg)[PER], int k, int lane) {
#pragma unroll
    for (int r = 0; r < PER; r++) {
        if ((r & J) == 0) {
            constexpr int dummy = 0;
            (void)dummy;
            const int h = r | J;
            const bool up = (((lane * PER) | r) & k) == 0;
            const bool sw = ro_greater(kv[r], tg[r], kv[h], tg[h]) == up;
            const double lo_k = sw ? kv[h] : kv[r], hi_k = sw ? kv[r] : kv[h];
            const int lo_t = sw ? tg[h] : tg[r], hi_t = sw ? tg[r] : tg[h];
            kv[r] = lo_k; kv[h] = hi_k;
            tg[r] = lo_t; tg[h] = hi_t;
        }
    }
}

template <int PER>
__device__ __forceinline__ void ro_cross_lane(double (&kv)[PER], int (&tg)[PER], int lm, int k, int lane) {
    const bool is_lo = (lane & lm) == 0;
#pragma unroll
    for (int r = 0; r < PER; r++) {
        const double ok = __shfl_xor_sync(0xffffffffu, kv[r], lm);
        const int ot = __shfl_xor_sync(0xffffffffu, tg[r], lm);
        const bool up = (((lane * PER) | r) & k) == 0;
        const bool take = (is_lo == up) == ro_greater(kv[r], tg[r], ok, ot);  // keep the smaller one iff (lo == up)
        kv[r] = take ? ok : kv[r];
        tg[r] = take ? ot : tg[r];
    }
}

// One warp per (x, i) cell.  The reference gives each cell to one thread, which bubble-sorts the
// 400 sums in local memory (K:3152-3171).  Here the warp builds the 400 (k-sum, slot) pairs in shared
// memory, sorts them with a 512-wide bitonic network (ties broken by the reference's slot index, which
// reproduces the stable order of its exchange sort), scans the weights and rebins to the 20 Gauss
// points.
__global__ void __launch_bounds__(RO_WARPS * 32)
k_add_to_mixed_opac(const double* __restrict__ vmr, const double* __restrict__ opac_spec,
                    double* __restrict__ opac_wg, const double* __restrict__ meanmolmass,
                    const double* __restrict__ gauss_weight, const double* __restrict__ gauss_y,
                    double mass_spec, int s, int ro_method, int ny, int nbin, int n_i) {
    __shared__ double s_key[RO_WARPS][RO_PAD];
    __shared__ double s_yg[RO_WARPS][RO_N2];
    __shared__ int s_tag[RO_WARPS][RO_PAD];
    __shared__ double s_mixed[RO_WARPS][32];
    __shared__ double s_new[RO_WARPS][32];
    __shared__ double s_hw[RO_WARPS][32];  // 0.5 * gauss_weight

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const long long cell = (long long)blockIdx.x * RO_WARPS + wid;
    if (cell >= (long long)nbin * n_i) return;  // whole warp leaves together
    const int i = (int)(cell / nbin);
    const int x = (int)(cell - (long long)i * nbin);
    double* __restrict__ out = opac_wg + ((size_t)i * nbin + x) * ny;
    const double* __restrict__ spec = opac_spec + ((size_t)i * nbin + x) * ny;

    const double scale_a = vmr[i] * mass_spec / meanmolmass[i];  // K:3293, evaluated left to right

    if (ny != RO_NY || ro_method == 0 || s == 0) {
        // correlated-k for arbitrary ny (K:3304-3310)
        for (int y = lane; y < ny; y += 32) out[y] += scale_a * spec[y];
        return;
    }

    if (lane < RO_NY) {
        s_mixed[wid][lane] = out[lane];
        s_new[wid][lane] = scale_a * spec[lane];
        s_hw[wid][lane] = 0.5 * gauss_weight[lane];
    }
    __syncwarp();
    const double* mixed = s_mixed[wid];
    const double* newk = s_new[wid];

    // 1 % rule (K:3297)
    const bool negligible = (0.01 * mixed[0] > newk[RO_NY - 1]) || (0.01 * newk[0] > mixed[RO_NY - 1]);
    if (negligible) {
        if (lane < RO_NY) out[lane] = mixed[lane] + newk[lane];
        return;
    }

    // intersection of the two k-curves: last y at which their ordering flips (K:3321-3329)
    bool flip = false;
    if (lane >= 1 && lane < RO_NY)
        flip = (mixed[lane] > newk[lane]) != (mixed[lane - 1] > newk[lane - 1]);
    const unsigned fm = __ballot_sync(0xffffffffu, flip);
    const int yi = fm ? (31 - __clz(fm)) : RO_NY;
    const bool mixed_outer = mixed[0] > newk[0];  // K:3332
    const int split = yi * RO_NY;

    // build the 400 sums in the reference's slot layout (K:3332-3365), padded with +inf to the 512 inputs of the
    // network.  Element e = 16 * lane + r lives in register r of its lane: the compare-exchange distances below 16
    // stay inside a lane (pure register work), the larger ones pair lanes through shuffles; shared memory only
    // receives the sorted result.
    constexpr int PER = RO_PAD / 32;  // 16 elements per lane
    double kv[PER];
    int tg[PER];
#pragma unroll
    for (int r = 0; r < PER; r++) {
        const int pos = lane * PER + r;
        double v = DBL_MAX;
        int t = pos;
        if (pos < RO_N2) {
            int y1, y2;
            if (mixed_outer) {
                if (pos < split) { y1 = pos / yi; y2 = pos - y1 * yi; }
                else { y2 = pos / RO_NY; y1 = pos - y2 * RO_NY; }
            } else {
                if (pos < split) { y2 = pos / yi; y1 = pos - y2 * yi; }
                else { y1 = pos / RO_NY; y2 = pos - y1 * RO_NY; }
            }
            v = mixed[y1] + newk[y2];
            t = pos | (y1 << 9) | (y2 << 14);
        }
        kv[r] = v;
        tg[r] = t;
    }

    // bitonic sort, ascending in (key, slot); ties are broken by the reference's slot index, which reproduces the
    // stable order of its exchange sort (K:3152-3171).  The network is driven by run-time (k, j) over five small
    // code blocks (four in-lane distances, one cross-lane exchange) so that it stays resident in the instruction
    // cache; compare-exchanges are branch-free selects.
    for (int k = 2; k <= RO_PAD; k <<= 1) {
        for (int j = k >> 1; j >= PER; j >>= 1) ro_cross_lane<PER>(kv, tg, j / PER, k, lane);
        if (k > 8) ro_in_lane<PER, 8>(kv, tg, k, lane);
        if (k > 4) ro_in_lane<PER, 4>(kv, tg, k, lane);
        if (k > 2) ro_in_lane<PER, 2>(kv, tg, k, lane);
        ro_in_lane<PER, 1>(kv, tg, k, lane);
    }
    double* key = s_key[wid];
    int* tag = s_tag[wid];
#pragma unroll
    for (int r = 0; r < PER; r++) {
        key[lane * PER + r] = kv[r];
        tag[lane * PER + r] = tg[r];
    }
    __syncwarp();

    // abscissae of the sorted k-function: yg[w] = sum_{m<w} wt[m] + 0.5 wt[w]  (K:3371-3376)
    const int CH = 13;  // 32 * 13 >= 400
    const int w0 = lane * CH;
    double local = 0.0;
    for (int m = 0; m < CH; m++) {
        const int w = w0 + m;
        if (w < RO_N2) {
            const int t = tag[w];
            local += s_hw[wid][(t >> 9) & 31] * s_hw[wid][(t >> 14) & 31];
        }
    }
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    double run = incl - local;  // exclusive prefix
    double* yg = s_yg[wid];
    for (int m = 0; m < CH; m++) {
        const int w = w0 + m;
        if (w < RO_N2) {
            const int t = tag[w];
            const double wt = s_hw[wid][(t >> 9) & 31] * s_hw[wid][(t >> 14) & 31];
            yg[w] = run + 0.5 * wt;
            run += wt;
        }
    }
    __syncwarp();

    // rebinning (K:3379-3396): y advances by at most one per w, starting at w = 1
    int first = RO_N2;  // smallest w >= 1 with yg[w] > gauss_y[y]
    double gy = 0.0;
    if (lane < RO_NY) {
        gy = gauss_y[lane];
        int lo = 1, hi = RO_N2;  // search in [1, 400)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (yg[mid] > gy) hi = mid; else lo = mid + 1;
        }
        first = lo;
    }
    int v = first - lane;  // w_y = y + max_{m<=y}(first(m) - m)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = max(v, o);
    }
    const int w = v + lane;
    if (lane < RO_NY && w < RO_N2) {
        out[lane] = (key[w - 1] * (yg[w] - gy) + key[w] * (gy - yg[w - 1])) / (yg[w] - yg[w - 1]);
    }
}

// K:3404-3440 with K:3174-3205.  The cross-section goes through (n^2 - 1) with n - 1 ~ 1e-7, i.e. it
// amplifies any last-bit difference in n by ~1e7; the libdevice calls (pow with the reference's exponents)
// are therefore kept exactly as the reference makes them.
__device__ __forceinline__ double h2o_refr_index(double wave, double press, double temp, double f_h2o,
                                                 double mass_h2o) {
    const double dens = f_h2o * press * mass_h2o / (hc::KBOLTZMANN * temp);
    const double lamda = wave / 0.589e-4;
    const double delta = fmin(1.0, dens) / 1.0;
    const double theta = temp / 273.15;
    const double lamda_UV = 0.229202, lamda_IR = 5.432937;
    const double a0 = 0.244257733, a1 = 0.974634476e-2, a2 = -0.373234996e-2, a3 = 0.268678472e-3,
                 a4 = 0.158920570e-2, a5 = 0.245934259e-2, a6 = 0.900704920, a7 = -0.166626219e-1;
    const double A = delta * (a0 + a1 * delta + a2 * theta + a3 * pow(1.0 * lamda, 2.0) * theta +
                              a4 * pow(1.0 * lamda, -2.0) +
                              a5 / (pow(1.0 * lamda, 2.0) - pow(1.0 * lamda_UV, 2.0)) +
                              a6 / (pow(1.0 * lamda, 2.0) - pow(1.0 * lamda_IR, 2.0)) +
                              a7 * pow(1.0 * delta, 2.0));
    return pow((2.0 * A + 1.0) / (1.0 - A), 0.5);
}

__global__ void k_calc_h2o_scat(const double* __restrict__ temp, const double* __restrict__ press,
                                const double* __restrict__ wave, double* __restrict__ scat_cross,
                                const double* __restrict__ vmr, double mass_h2o, int nbin, int n_i) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)nbin * n_i) return;
    const int i = (int)(e / nbin), x = (int)(e - (long long)i * nbin);
    double sc = 0.0;
    if (wave[x] < 2.5e-4) {
        const double index = h2o_refr_index(wave[x], press[i], temp[i], vmr[i], mass_h2o);
        const double n_ref = vmr[i] * press[i] / (hc::KBOLTZMANN * temp[i]);
        const double King = (6.0 + 3.0 * 3e-4) / (6.0 - 7.0 * 3e-4);
        sc = 24.0 * pow(1.0 * hc::PI, 3.0) / (pow(1.0 * n_ref, 2.0) * pow(1.0 * wave[x], 4.0)) *
             pow((pow(1.0 * index, 2.0) - 1.0) / (pow(1.0 * index, 2.0) + 2.0), 2.0) * King;
    }
    scat_cross[e] = sc;
}

__global__ void k_add_to_mixed_scat(const double* __restrict__ vmr, const double* __restrict__ spec,
                                    double* __restrict__ scat, int nbin, int n_i) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)nbin * n_i) return;
    const int i = (int)(e / nbin);
    scat[e] += vmr[i] * spec[e];
}

__global__ void k_total_g0(const double* __restrict__ scat, const double* __restrict__ g0c,
                           const double* __restrict__ scatc, double* __restrict__ g0tot, double g_0,
                           long long n) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double num = g_0 * scat[e] + g0c[e] * scatc[e];
    const double den = scat[e] + scatc[e];
    g0tot[e] = num / den;
}

extern "C" {

int helios_add_to_mixed_opac(helios_ctx* ctx, const double* vmr, const double* opac_spec,
                             double* opac_wg, const double* meanmolmass, const double* gauss_weight,
                             const double* gauss_y, double mass_spec, int s, int ro_method, int ny,
                             int nbin, int nlay_or_nint) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(vmr && opac_spec && opac_wg && meanmolmass && gauss_weight && gauss_y);
    HARG(ny > 0 && nbin > 0 && nlay_or_nint > 0 && s >= 0);
    if (ro_method != 0 && s != 0 && ny != 1 && ny != RO_NY) {
        helios_set_error("helios_add_to_mixed_opac: random overlap needs ny == 20 (got %d), as in the "
                         "reference (kernels.cu:3314)", ny);
        return HELIOS_ERR_ARG;
    }
    const long long cells = (long long)nbin * nlay_or_nint;
    k_add_to_mixed_opac<<<ceil_div(cells, RO_WARPS), RO_WARPS * 32, 0, ctx->stream>>>(
        vmr, opac_spec, opac_wg, meanmolmass, gauss_weight, gauss_y, mass_spec, s, ro_method, ny, nbin,
        nlay_or_nint);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_h2o_scat(helios_ctx* ctx, const double* temp, const double* press, const double* wave,
                         double* scat_cross, const double* vmr, double mass_h2o, int nbin,
                         int nlay_or_nint) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(temp && press && wave && scat_cross && vmr && nbin > 0 && nlay_or_nint > 0);
    const long long n = (long long)nbin * nlay_or_nint;
    k_calc_h2o_scat<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(temp, press, wave, scat_cross, vmr,
                                                               mass_h2o, nbin, nlay_or_nint);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_add_to_mixed_scat(helios_ctx* ctx, const double* vmr, const double* scat_cross_spec,
                             double* scat_cross, int nbin, int nlay_or_nint) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(vmr && scat_cross_spec && scat_cross && nbin > 0 && nlay_or_nint > 0);
    const long long n = (long long)nbin * nlay_or_nint;
    k_add_to_mixed_scat<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(vmr, scat_cross_spec, scat_cross, nbin,
                                                                   nlay_or_nint);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_total_g_0_of_gas_and_clouds(helios_ctx* ctx, const double* scat_cross,
                                            const double* g_0_all_clouds,
                                            const double* scat_cross_all_clouds, double* g_0_tot,
                                            double g_0, int nbin, int nlay_or_nint) {
    HCTX(ctx);
    HARG(scat_cross && g_0_all_clouds && scat_cross_all_clouds && g_0_tot && nbin > 0 && nlay_or_nint > 0);
    // element-wise over [i][x] arrays that all have exactly nlay_or_nint rows per atmosphere: a batch is
    // simply nbatch times as many elements
    const long long n = (long long)nbin * nlay_or_nint * ctx->batch.nbatch;
    k_total_g0<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(scat_cross, g_0_all_clouds,
                                                          scat_cross_all_clouds, g_0_tot, g_0, n);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

}  // extern "C"
