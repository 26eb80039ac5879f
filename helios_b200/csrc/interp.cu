// Planck table, stellar energy correction, temperature / Planck / opacity interpolation.
// From-scratch sm_100a kernels for K:362-1011 and K:3209-3259 of the reference.
#include "common.cuh"
#include <cstdlib>

#ifndef PT_GATHER_TMA_DEFAULT
#define PT_GATHER_TMA_DEFAULT 0  // measured on B200: 35.8 us staged vs 31.7 us streamed for the C2 refresh
#endif

// ------------------------------------------------------------------------------------------
// Planck table (K:362-416).  One thread per table entry (x, row); rows 0..dim-1 have
// T = 1 + row*step, row `dim` is the stellar temperature.  The 199-term series is summed in the
// reference's order so that the table agrees to rounding.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double planck_series_term(int n, double y1, double y2) {
    // K:95-105
    const double dn = n;
    const double d2 = dn * dn, d3 = dn * dn * dn, d4 = dn * dn * dn * dn;
    return exp(-dn * y2) * ((y2 * y2 * y2) / dn + 3.0 * (y2 * y2) / d2 + 6.0 * y2 / d3 + 6.0 / d4) -
           exp(-dn * y1) * ((y1 * y1 * y1) / dn + 3.0 * (y1 * y1) / d2 + 6.0 * y1 / d3 + 6.0 / d4);
}

__global__ void __launch_bounds__(256)
k_plancktable(double* __restrict__ grid, const double* __restrict__ lambda_edge,
              const double* __restrict__ deltalambda, int nwave, double Tstar, int dim, int step) {
    const long long total = (long long)(dim + 1) * nwave;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(e / nwave);
        const int x = (int)(e - (long long)row * nwave);
        const double T = (row < dim) ? (double)(row * step + 1) : Tstar;
        double acc = 0.0;
        if (T > 0.01) {
            const double kh = hc::KBOLTZMANN / hc::HCONST;
            const double D = 2.0 * ((kh * kh * kh) * hc::KBOLTZMANN * (T * T * T * T)) /
                             (hc::CSPEED * hc::CSPEED);
            double y_top = hc::HCONST * hc::CSPEED / (lambda_edge[x + 1] * hc::KBOLTZMANN * T);
            double y_bot = hc::HCONST * hc::CSPEED / (lambda_edge[x] * hc::KBOLTZMANN * T);
            if (y_bot < y_top) {
                const double s = y_top;
                y_top = y_bot;
                y_bot = s;
            }
            for (int n = 1; n < 200; n++) acc += D * planck_series_term(n, y_bot, y_top);
        }
        grid[e] = acc / deltalambda[x];
    }
}

// ------------------------------------------------------------------------------------------
// Stellar energy correction (K:420-468): one block sums the incident flux in a fixed order,
// a second kernel rescales.  (The reference lets every thread redo the whole O(nbin) sum.)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_inc_energy_sum(const double* __restrict__ planck_grid, const double* __restrict__ starflux,
                 const double* __restrict__ deltalambda, int realstar, int nwave, double Tstar, int dim,
                 double* __restrict__ corr_out) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (int x = threadIdx.x; x < nwave; x += blockDim.x) {
        acc += realstar == 1 ? deltalambda[x] * starflux[x]
                             : deltalambda[x] * hc::PI * planck_grid[x + (size_t)dim * nwave];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double theo = hc::STEFANBOLTZMANN * pow(Tstar, 4.0);
        corr_out[0] = theo / red[0];
    }
}

__global__ void k_inc_energy_scale(double* __restrict__ planck_grid, double* __restrict__ starflux,
                                   int realstar, int nwave, int dim, const double* __restrict__ corr) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nwave) return;
    const double f = corr[0];
    if (realstar == 1) starflux[x] *= f;
    else planck_grid[x + (size_t)dim * nwave] *= f;
}

// ------------------------------------------------------------------------------------------
// temp_inter (K:496-520)
// ------------------------------------------------------------------------------------------
__global__ void k_temp_inter(const double* __restrict__ tlay, double* __restrict__ tint, int nint) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nint) return;
    tlay += (size_t)blockIdx.y * nint;  // batch: T_lay has nlayer + 1 = nint entries per atmosphere
    tint += (size_t)blockIdx.y * nint;
    if (i == 0) tint[i] = tlay[i] - 0.5 * (tlay[i + 1] - tlay[i]);
    else if (i == nint - 1) tint[i] = tlay[i - 1] + 0.5 * (tlay[i - 1] - tlay[i - 2]);
    else tint[i] = tlay[i - 1] + 0.5 * (tlay[i] - tlay[i - 1]);
}

// ------------------------------------------------------------------------------------------
// Planck interpolation (K:923-1011).  Output layout is the reference's [x][i] (i fastest); a
// 32x32 shared tile turns the x-coalesced table reads into i-coalesced stores.
//   mode 0: layer version, rows i<nl from temp[i], row nl = star slot, row nl+1 = temp[nl] (BOA)
//   mode 1: interface version, rows i<nrows from temp[i]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_planck_interpol(const double* __restrict__ temp, double* __restrict__ out,
                  const double* __restrict__ planck_grid, const double* __restrict__ starflux,
                  int realstar, int mode, int nl, int nrows, int nwave, int dim, int step,
                  const double* __restrict__ planck_star) {
    __shared__ double tile[32][33];
    {   // batch (blockIdx.z = atmosphere): T_lay holds nl + 1 values, T_int nrows
        const size_t b = blockIdx.z;
        temp += b * (size_t)(mode == 0 ? nl + 1 : nrows);
        out += b * (size_t)nrows * nwave;
        if (starflux) starflux += b * (size_t)nwave;
        if (planck_star) planck_star += b * (size_t)nwave;
    }
    const int x0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, x = x0 + tx;
        double v = 0.0;
        if (i < nrows && x < nwave) {
            if (mode == 0 && i == nl) {
                v = realstar == 1 ? starflux[x] / hc::PI
                                  : (planck_star ? planck_star[x] : planck_grid[x + (size_t)dim * nwave]);
            } else {
                const double Ti = (mode == 0 && i == nl + 1) ? temp[nl] : temp[i];
                double t = (Ti - 1.0) / step;
                t = fmax(0.001, fmin(dim - 1.001, t));
                const int tdown = (int)floor(t), tup = (int)ceil(t);
                if (tdown != tup)
                    v = planck_grid[x + (size_t)tdown * nwave] * (tup - t) +
                        planck_grid[x + (size_t)tup * nwave] * (t - tdown);
                else
                    v = planck_grid[x + (size_t)tdown * nwave];
            }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, i = i0 + tx;
        if (x < nwave && i < nrows) out[i + (size_t)x * nrows] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------
// Fused per-iteration preparation: temp_inter + planck_interpol_layer + planck_interpol_interface in ONE launch
// (three launches in the reference, C:856-857; they are the only work besides the flux solve in 9 of 10 iterations).
// Same formulas, same operation order as the three kernels above -> bitwise the same results.
//   blockIdx.y <  ytl : 32-row tile of the layer table (rows 0..nl-1 layers, nl = star, nl+1 = surface); the blocks
//                       of the first x-tile also write T_int (every interface index is < nl + 2)
//   blockIdx.y >= ytl : 32-row tile of the interface table (non-isothermal only)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double interface_temperature(const double* __restrict__ tlay, int i, int nint) {
    if (i == 0) return tlay[i] - 0.5 * (tlay[i + 1] - tlay[i]);
    if (i == nint - 1) return tlay[i - 1] + 0.5 * (tlay[i - 1] - tlay[i - 2]);
    return tlay[i - 1] + 0.5 * (tlay[i] - tlay[i - 1]);
}

__global__ void __launch_bounds__(256)
k_iter_prep(const double* __restrict__ tlay, double* __restrict__ tint, double* __restrict__ planck_lay,
            double* __restrict__ planck_int, const double* __restrict__ planck_grid,
            const double* __restrict__ starflux, int realstar, int nl, int nwave, int dim, int step, int ytl,
            const double* __restrict__ planck_star) {
    __shared__ double tile[32][33];
    const int nint = nl + 1;
    const bool lay = (int)blockIdx.y < ytl;
    const int nrows = lay ? nl + 2 : nint;
    {   // batch (blockIdx.z = atmosphere)
        const size_t b = blockIdx.z;
        tlay += b * (size_t)nint;
        tint += b * (size_t)nint;
        planck_lay += b * (size_t)(nl + 2) * nwave;
        if (planck_int) planck_int += b * (size_t)nint * nwave;
        if (starflux) starflux += b * (size_t)nwave;
        if (planck_star) planck_star += b * (size_t)nwave;
    }
    double* __restrict__ out = lay ? planck_lay : planck_int;
    const int x0 = blockIdx.x * 32, i0 = (lay ? blockIdx.y : blockIdx.y - ytl) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    if (lay && blockIdx.x == 0 && ty == 0) {
        const int i = i0 + tx;
        if (i < nint) tint[i] = interface_temperature(tlay, i, nint);
    }
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, x = x0 + tx;
        double v = 0.0;
        if (i < nrows && x < nwave) {
            if (lay && i == nl) {
                v = realstar == 1 ? starflux[x] / hc::PI
                                  : (planck_star ? planck_star[x] : planck_grid[x + (size_t)dim * nwave]);
            } else {
                const double Ti = lay ? (i == nl + 1 ? tlay[nl] : tlay[i]) : interface_temperature(tlay, i, nint);
                double t = (Ti - 1.0) / step;
                t = fmax(0.001, fmin(dim - 1.001, t));
                const int tdown = (int)floor(t), tup = (int)ceil(t);
                if (tdown != tup)
                    v = planck_grid[x + (size_t)tdown * nwave] * (tup - t) +
                        planck_grid[x + (size_t)tup * nwave] * (t - tdown);
                else
                    v = planck_grid[x + (size_t)tdown * nwave];
            }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, i = i0 + tx;
        if (x < nwave && i < nrows) out[i + (size_t)x * nrows] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------
// (P,T) table interpolation (K:524-645, 649-919, 3209-3259).
// A tiny prep kernel resolves each layer's box once (the reference recomputes the log10 of the
// grid ends in every thread, K:545-554); the gather kernel then streams the four table rows of
// that box, which are contiguous nbin*ny-long vectors, into the output row.
// ------------------------------------------------------------------------------------------
struct PTBox {
    double p, t;
    int pdown, pup, tdown, tup;
};

// clamp_mode 0: [0.001, n-1.001] (K:549, 556);  1: [0, n-1] (K:3233, 3238)
__global__ void k_pt_prep(const double* __restrict__ temp, const double* __restrict__ press,
                          const double* __restrict__ gtemp, const double* __restrict__ gpress, int ntemp,
                          int npress, int n, int clamp_mode, int log_t, PTBox* __restrict__ box, int tstride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // batch (blockIdx.y = atmosphere): temperatures are tstride apart (T_lay: nlayer + 1), pressures n
    temp += (size_t)blockIdx.y * tstride;
    press += (size_t)blockIdx.y * n;
    box += (size_t)blockIdx.y * n;
    double t;
    if (log_t) {
        const double dT = (log10(gtemp[ntemp - 1]) - log10(gtemp[0])) / (ntemp - 1.0);
        t = (log10(temp[i]) - log10(gtemp[0])) / dT;
    } else {
        const double dT = (gtemp[ntemp - 1] - gtemp[0]) / (ntemp - 1.0);
        t = (temp[i] - gtemp[0]) / dT;
    }
    const double dP = (log10(gpress[npress - 1]) - log10(gpress[0])) / (npress - 1.0);
    double p = (log10(press[i]) - log10(gpress[0])) / dP;
    if (clamp_mode == 0) {
        t = fmin(ntemp - 1.001, fmax(0.001, t));
        p = fmin(npress - 1.001, fmax(0.001, p));
    } else {
        t = fmin(ntemp - 1.0, fmax(0.0, t));
        p = fmin(npress - 1.0, fmax(0.0, p));
    }
    PTBox b;
    b.p = p;
    b.t = t;
    b.tdown = (int)floor(t);
    b.tup = (int)ceil(t);
    b.pdown = (int)floor(p);
    b.pup = (int)ceil(p);
    box[i] = b;
}

// the four-branch bilinear form of K:561-608 / K:613-645
__device__ __forceinline__ double bilin4(double dd, double ud, double du, double uu, const PTBox& b) {
    const bool pe = b.pdown == b.pup, te = b.tdown == b.tup;
    if (!pe && !te)
        return dd * (b.pup - b.p) * (b.tup - b.t) + ud * (b.p - b.pdown) * (b.tup - b.t) +
               du * (b.pup - b.p) * (b.t - b.tdown) + uu * (b.p - b.pdown) * (b.t - b.tdown);
    if (te && !pe) return dd * (b.pup - b.p) + ud * (b.p - b.pdown);
    if (pe && !te) return dd * (b.tup - b.t) + du * (b.t - b.tdown);
    return dd;
}

// out[i][c] for c in [0, rowlen), table[t][p][c]; optional second table (cross sections)
__global__ void __launch_bounds__(256)
k_pt_gather(const PTBox* __restrict__ box, const double* __restrict__ table, double* __restrict__ out,
            int rowlen, const double* __restrict__ table2, double* __restrict__ out2, int rowlen2,
            int npress, int n, const int* __restrict__ table_index, size_t tstride, size_t tstride2,
            size_t ostride, size_t ostride2) {
    const int i = blockIdx.y;
    {   // batch (blockIdx.z = atmosphere)
        const size_t a = blockIdx.z;
        const size_t which = table_index ? (size_t)table_index[a] : 0;
        box += a * (size_t)n;
        if (table != nullptr) table += which * tstride;
        out += a * ostride;
        if (table2 != nullptr) {
            table2 += which * tstride2;
            out2 += a * ostride2;
        }
    }
    const PTBox b = box[i];
    const size_t r_dd = (size_t)b.pdown + (size_t)npress * b.tdown;
    const size_t r_ud = (size_t)b.pup + (size_t)npress * b.tdown;
    const size_t r_du = (size_t)b.pdown + (size_t)npress * b.tup;
    const size_t r_uu = (size_t)b.pup + (size_t)npress * b.tup;
    if (table != nullptr) {
        const double* __restrict__ t_dd = table + r_dd * rowlen;
        const double* __restrict__ t_ud = table + r_ud * rowlen;
        const double* __restrict__ t_du = table + r_du * rowlen;
        const double* __restrict__ t_uu = table + r_uu * rowlen;
        double* __restrict__ o = out + (size_t)i * rowlen;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < rowlen; c += gridDim.x * blockDim.x)
            o[c] = bilin4(__ldg(t_dd + c), __ldg(t_ud + c), __ldg(t_du + c), __ldg(t_uu + c), b);
    }
    if (table2 != nullptr) {
        const double* __restrict__ t_dd = table2 + r_dd * rowlen2;
        const double* __restrict__ t_ud = table2 + r_ud * rowlen2;
        const double* __restrict__ t_du = table2 + r_du * rowlen2;
        const double* __restrict__ t_uu = table2 + r_uu * rowlen2;
        double* __restrict__ o = out2 + (size_t)i * rowlen2;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < rowlen2; c += gridDim.x * blockDim.x)
            o[c] = bilin4(__ldg(t_dd + c), __ldg(t_ud + c), __ldg(t_du + c), __ldg(t_uu + c), b);
    }
}

// The same gather with the table rows staged through shared memory by TMA (north_star: "tables staged through
// TMA/shared memory for the interpolation gathers").  A CTA owns one chunk of GT_CHUNK columns of one layer: one thread
// requests the chunk of each of the four rows of the layer's (P, T) box as a cp.async.bulk copy (SASS UBLKCP) that
// completes on an mbarrier; the block then combines them from shared memory and stores the output row coalesced.
// Rows start at r * rowlen doubles: the 16-byte alignment TMA needs holds for even rowlen (else the LDG form above).
constexpr int GT_CHUNK = 1024;

__global__ void __launch_bounds__(128)
k_pt_gather_tma(const PTBox* __restrict__ box, const double* __restrict__ table, double* __restrict__ out, int rowlen,
                int npress, int n, const int* __restrict__ table_index, size_t tstride, size_t ostride) {
    __shared__ __align__(16) double rows[4][GT_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    const int i = blockIdx.y;
    {
        const size_t a = blockIdx.z;
        box += a * (size_t)n;
        table += (table_index ? (size_t)table_index[a] : 0) * tstride;
        out += a * ostride;
    }
    const PTBox b = box[i];
    const int c0 = blockIdx.x * GT_CHUNK;
    const int len = min(GT_CHUNK, rowlen - c0);
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t r[4] = {(size_t)b.pdown + (size_t)npress * b.tdown, (size_t)b.pup + (size_t)npress * b.tdown,
                             (size_t)b.pdown + (size_t)npress * b.tup, (size_t)b.pup + (size_t)npress * b.tup};
        const unsigned bytes = (unsigned)len * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(4u * bytes) : "memory");
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&rows[k][0]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(table + r[k] * rowlen + c0), "r"(bytes), "r"(bar_s)
                         : "memory");
        }
    }
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar_s),
        "r"(0u)
        : "memory");
    double* __restrict__ o = out + (size_t)i * rowlen + c0;
    for (int c = threadIdx.x; c < len; c += blockDim.x) o[c] = bilin4(rows[0][c], rows[1][c], rows[2][c], rows[3][c], b);
}

// scalar tables tab[p + npress*t] (K:649-919)
__global__ void k_pt_scalar(const double* __restrict__ temp, const double* __restrict__ press,
                            const double* __restrict__ gtemp, const double* __restrict__ gpress, int ntemp,
                            int npress, int n, int log_t, const double* __restrict__ tab,
                            double* __restrict__ out, int tstride, const int* __restrict__ table_index,
                            size_t tabstride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    temp += (size_t)blockIdx.y * tstride;  // batch (blockIdx.y = atmosphere)
    press += (size_t)blockIdx.y * n;
    out += (size_t)blockIdx.y * n;
    if (table_index) tab += (size_t)table_index[blockIdx.y] * tabstride;
    double t;
    if (log_t) {
        const double dT = (log10(gtemp[ntemp - 1]) - log10(gtemp[0])) / (ntemp - 1.0);
        t = (log10(temp[i]) - log10(gtemp[0])) / dT;
    } else {
        const double dT = (gtemp[ntemp - 1] - gtemp[0]) / (ntemp - 1.0);
        t = (temp[i] - gtemp[0]) / dT;
    }
    const double dP = (log10(gpress[npress - 1]) - log10(gpress[0])) / (npress - 1.0);
    double p = (log10(press[i]) - log10(gpress[0])) / dP;
    PTBox b;
    b.t = fmin(ntemp - 1.001, fmax(0.001, t));
    b.p = fmin(npress - 1.001, fmax(0.001, p));
    b.tdown = (int)floor(b.t);
    b.tup = (int)ceil(b.t);
    b.pdown = (int)floor(b.p);
    b.pup = (int)ceil(b.p);
    out[i] = bilin4(tab[b.pdown + npress * b.tdown], tab[b.pup + npress * b.tdown],
                    tab[b.pdown + npress * b.tup], tab[b.pup + npress * b.tup], b);
}

// ------------------------------------------------------------------------------------------
extern "C" {

int helios_plancktable(helios_ctx* ctx, double* planck_grid, const double* lambda_edge,
                       const double* deltalambda, int nwave, double Tstar, int dim, int step) {
    HCTX(ctx);
    HARG(planck_grid && lambda_edge && deltalambda && nwave > 0 && dim >= 0 && step > 0);
    HNOBATCH(ctx);
    const long long total = (long long)(dim + 1) * nwave;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->num_sms * 64;
    if (blocks > cap) blocks = cap;
    k_plancktable<<<(int)blocks, 256, 0, ctx->stream>>>(planck_grid, lambda_edge, deltalambda, nwave,
                                                        Tstar, dim, step);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_corr_inc_energy(helios_ctx* ctx, double* planck_grid, double* starflux,
                           const double* deltalambda, int realstar, int nwave, double Tstar, int dim,
                           double* corr_factor_host) {
    HCTX(ctx);
    HARG(planck_grid && deltalambda && nwave > 0 && dim >= 0);
    HARG(realstar == 0 || starflux != nullptr);
    HNOBATCH(ctx);
    double* scratch = nullptr;
    int rc = helios_ctx_scratch(ctx, sizeof(double), &scratch);
    if (rc) return rc;
    k_inc_energy_sum<<<1, 1024, 0, ctx->stream>>>(planck_grid, starflux, deltalambda, realstar, nwave,
                                                  Tstar, dim, scratch);
    HLAUNCHED(ctx);
    k_inc_energy_scale<<<ceil_div(nwave, 256), 256, 0, ctx->stream>>>(planck_grid, starflux, realstar,
                                                                      nwave, dim, scratch);
    HLAUNCHED(ctx);
    if (corr_factor_host) {
        HCUDA(cudaMemcpyAsync(corr_factor_host, scratch, sizeof(double), cudaMemcpyDeviceToHost,
                              ctx->stream));
        HCUDA(cudaStreamSynchronize(ctx->stream));
    }
    return HELIOS_OK;
}

int helios_temp_inter(helios_ctx* ctx, const double* tlay, double* tint, int numinterfaces) {
    HCTX(ctx);
    HARG(tlay && tint && numinterfaces >= 3);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint());
    k_temp_inter<<<dim3(ceil_div(numinterfaces, 128), ctx->batch.nbatch), 128, 0, ctx->stream>>>(tlay, tint,
                                                                                                  numinterfaces);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_planck_interpol_layer(helios_ctx* ctx, const double* temp, double* planckband_lay,
                                 const double* planck_grid, const double* starflux, int realstar,
                                 int numlayers, int nwave, int dim, int step) {
    HCTX(ctx);
    HARG(temp && planckband_lay && planck_grid && numlayers > 0 && nwave > 0 && dim > 1 && step > 0);
    HARG(realstar == 0 || starflux != nullptr);
    const int nrows = numlayers + 2;
    HBATCHDIMS(ctx, numlayers == ctx->batch.nlayer && nwave == ctx->batch.nbin);
    dim3 grid(ceil_div(nwave, 32), ceil_div(nrows, 32), ctx->batch.nbatch);
    k_planck_interpol<<<grid, 256, 0, ctx->stream>>>(temp, planckband_lay, planck_grid, starflux,
                                                     realstar, 0, numlayers, nrows, nwave, dim, step,
                                                     ctx->batch.active ? ctx->batch.planck_star : nullptr);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_planck_interpol_interface(helios_ctx* ctx, const double* temp, double* planckband_int,
                                     const double* planck_grid, int numinterfaces, int nwave, int dim,
                                     int step) {
    HCTX(ctx);
    HARG(temp && planckband_int && planck_grid && numinterfaces > 0 && nwave > 0 && dim > 1 && step > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nwave == ctx->batch.nbin);
    dim3 grid(ceil_div(nwave, 32), ceil_div(numinterfaces, 32), ctx->batch.nbatch);
    k_planck_interpol<<<grid, 256, 0, ctx->stream>>>(temp, planckband_int, planck_grid, nullptr, 0, 1, 0,
                                                     numinterfaces, nwave, dim, step, nullptr);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_iteration_prepare(helios_ctx* ctx, const double* tlay, double* tint, double* planckband_lay,
                             double* planckband_int, const double* planck_grid, const double* starflux,
                             int realstar, int numlayers, int nwave, int dim, int step) {
    HCTX(ctx);
    HARG(tlay && tint && planckband_lay && planck_grid && numlayers > 1 && nwave > 0 && dim > 1 && step > 0);
    HARG(realstar == 0 || starflux != nullptr);
    HBATCHDIMS(ctx, numlayers == ctx->batch.nlayer && nwave == ctx->batch.nbin);
    const int ytl = ceil_div(numlayers + 2, 32);
    const int yti = planckband_int ? ceil_div(numlayers + 1, 32) : 0;
    dim3 grid(ceil_div(nwave, 32), ytl + yti, ctx->batch.nbatch);
    k_iter_prep<<<grid, 256, 0, ctx->stream>>>(tlay, tint, planckband_lay, planckband_int, planck_grid, starflux,
                                               realstar, numlayers, nwave, dim, step, ytl,
                                               ctx->batch.active ? ctx->batch.planck_star : nullptr);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

static int g_pt_gather_tma = -1;  // -1: not decided yet (environment / default), 0: streamed, 1: TMA-staged

extern "C" int helios_set_pt_gather_tma(int on) {
    const int before = g_pt_gather_tma;
    g_pt_gather_tma = on < 0 ? -1 : (on ? 1 : 0);
    return before;
}

static int pt_table_interp(helios_ctx* ctx, const double* temp, const double* gtemp, const double* press,
                           const double* gpress, const double* table, double* out, int rowlen,
                           const double* table2, double* out2, int rowlen2, int npress, int ntemp, int n,
                           int clamp_mode) {
    const BatchDesc& bd = ctx->batch;
    const int nb = bd.nbatch;
    double* scratch = nullptr;
    int rc = helios_ctx_scratch(ctx, sizeof(PTBox) * (size_t)n * nb + 64, &scratch);
    if (rc) return rc;
    // keep clear of the first 64 bytes, which small reductions use
    PTBox* box = reinterpret_cast<PTBox*>(reinterpret_cast<char*>(scratch) + 64);
    // batch: T_lay carries the surface value behind the nlayer layer values, T_int has exactly n entries
    const int tstride = (bd.active && n == bd.nlayer) ? n + 1 : n;
    k_pt_prep<<<dim3(ceil_div(n, 128), nb), 128, 0, ctx->stream>>>(temp, press, gtemp, gpress, ntemp, npress, n,
                                                                   clamp_mode, 0, box, tstride);
    HLAUNCHED(ctx);
    // HELIOS_PT_GATHER: "tma" = table rows staged by cp.async.bulk (k_pt_gather_tma), "ldg" = streamed with __ldg;
    // default: see the measurement in DESIGN.md 6b
    if (g_pt_gather_tma < 0) {
        const char* e = getenv("HELIOS_PT_GATHER");
        g_pt_gather_tma = e == nullptr ? PT_GATHER_TMA_DEFAULT : (e[0] == 't' ? 1 : 0);
    }
    if (g_pt_gather_tma == 1 && rowlen % 2 == 0) {
        dim3 tgrid(ceil_div(rowlen, GT_CHUNK), n, nb);
        k_pt_gather_tma<<<tgrid, 128, 0, ctx->stream>>>(box, table, out, rowlen, npress, n, bd.active ? bd.table_index : nullptr,
                                                       bd.ktable_stride, bd.active ? bd.wg() : 0);
        HLAUNCHED(ctx);
        if (table2 == nullptr) return HELIOS_OK;
        // the cross-section table (rowlen2 = nbin doubles per row) stays on the streaming form
        table = nullptr;
    }
    int bx = ceil_div(rowlen, 256);
    const int cap = (ctx->num_sms * 8 + n * nb - 1) / (n * nb);
    if (bx > cap) bx = cap > 0 ? cap : 1;
    dim3 grid(bx, n, nb);
    // per-atmosphere output strides: every [i][x][y] array is allocated with ninterface rows (Q:407), the
    // [i][x] cross-section arrays with exactly n rows
    k_pt_gather<<<grid, 256, 0, ctx->stream>>>(box, table, out, rowlen, table2, out2, rowlen2, npress, n,
                                               bd.active ? bd.table_index : nullptr, bd.ktable_stride, bd.cross_stride,
                                               bd.active ? bd.wg() : 0, (size_t)n * rowlen2);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_opac_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                         const double* press, const double* opacpress, const double* ktable,
                         double* opac, const double* crosstable, double* scat_cross, int npress,
                         int ntemp, int ny, int nbin, int nlay_or_nint) {
    HCTX(ctx);
    HARG(temp && opactemp && press && opacpress && ktable && opac && crosstable && scat_cross);
    HARG(npress > 1 && ntemp > 1 && ny > 0 && nbin > 0 && nlay_or_nint > 0);
    HBATCHDIMS(ctx, ny == ctx->batch.ny && nbin == ctx->batch.nbin &&
                    (nlay_or_nint == ctx->batch.nlayer || nlay_or_nint == ctx->batch.nint()));
    return pt_table_interp(ctx, temp, opactemp, press, opacpress, ktable, opac, ny * nbin, crosstable,
                           scat_cross, nbin, npress, ntemp, nlay_or_nint, 0);
}

int helios_opac_species_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                                 const double* press, const double* opacpress,
                                 const double* opac_opacity_pretab, double* opac_spec_wg, int npress,
                                 int ntemp, int ny, int nbin, int nlay_or_nint) {
    HCTX(ctx);
    HARG(temp && opactemp && press && opacpress && opac_opacity_pretab && opac_spec_wg);
    HARG(npress > 1 && ntemp > 1 && ny > 0 && nbin > 0 && nlay_or_nint > 0);
    HNOBATCH(ctx);
    return pt_table_interp(ctx, temp, opactemp, press, opacpress, opac_opacity_pretab, opac_spec_wg,
                           ny * nbin, nullptr, nullptr, 0, npress, ntemp, nlay_or_nint, 1);
}

static int pt_scalar(helios_ctx* ctx, const double* temp, const double* gtemp, const double* press,
                     const double* gpress, double* out, const double* tab, int npress, int ntemp, int n,
                     int log_t, bool batched = false) {
    const BatchDesc& bd = ctx->batch;
    if (!batched) HNOBATCH(ctx);
    const int nb = bd.nbatch;
    const int tstride = (bd.active && n == bd.nlayer) ? n + 1 : n;
    k_pt_scalar<<<dim3(ceil_div(n, 128), nb), 128, 0, ctx->stream>>>(temp, press, gtemp, gpress, ntemp, npress, n,
                                                                     log_t, tab, out, tstride,
                                                                     bd.active ? bd.table_index : nullptr,
                                                                     bd.mmass_stride);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

#define SCALAR_ARGS_OK(t, gt, p, gp, o, tab, np, nt, n) \
    HARG(t && gt && p && gp && o && tab && np > 1 && nt > 1 && n > 0)

int helios_meanmolmass_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                                double* meanmolmass, const double* opac_meanmass, const double* press,
                                const double* opacpress, int npress, int ntemp, int ninterface) {
    HCTX(ctx);
    SCALAR_ARGS_OK(temp, opactemp, press, opacpress, meanmolmass, opac_meanmass, npress, ntemp, ninterface);
    HBATCHDIMS(ctx, ninterface == ctx->batch.nlayer || ninterface == ctx->batch.nint());
    return pt_scalar(ctx, temp, opactemp, press, opacpress, meanmolmass, opac_meanmass, npress, ntemp,
                     ninterface, 0, true);
}

int helios_kappa_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                          const double* press, const double* entr_press, double* kappa,
                          const double* entr_kappa, int entr_npress, int entr_ntemp, int nlay_or_nint) {
    HCTX(ctx);
    SCALAR_ARGS_OK(temp, entr_temp, press, entr_press, kappa, entr_kappa, entr_npress, entr_ntemp,
                   nlay_or_nint);
    return pt_scalar(ctx, temp, entr_temp, press, entr_press, kappa, entr_kappa, entr_npress, entr_ntemp,
                     nlay_or_nint, 0);
}

int helios_cp_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp, const double* press,
                       const double* entr_press, double* cp_lay, const double* entr_cp, int entr_npress,
                       int entr_ntemp, int nlayer) {
    HCTX(ctx);
    SCALAR_ARGS_OK(temp, entr_temp, press, entr_press, cp_lay, entr_cp, entr_npress, entr_ntemp, nlayer);
    return pt_scalar(ctx, temp, entr_temp, press, entr_press, cp_lay, entr_cp, entr_npress, entr_ntemp,
                     nlayer, 1);
}

int helios_entropy_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                            const double* press, const double* entr_press, double* entropy,
                            const double* entr_entropy, int entr_npress, int entr_ntemp, int nlayer) {
    HCTX(ctx);
    SCALAR_ARGS_OK(temp, entr_temp, press, entr_press, entropy, entr_entropy, entr_npress, entr_ntemp,
                   nlayer);
    return pt_scalar(ctx, temp, entr_temp, press, entr_press, entropy, entr_entropy, entr_npress,
                     entr_ntemp, nlayer, 1);
}

int helios_phase_number_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                                 const double* press, const double* entr_press, double* state,
                                 const double* entr_state, int entr_npress, int entr_ntemp, int nlayer) {
    HCTX(ctx);
    SCALAR_ARGS_OK(temp, entr_temp, press, entr_press, state, entr_state, entr_npress, entr_ntemp, nlayer);
    return pt_scalar(ctx, temp, entr_temp, press, entr_press, state, entr_state, entr_npress, entr_ntemp,
                     nlayer, 0);
}

}  // extern "C"
