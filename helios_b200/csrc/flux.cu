// Band / Gauss-point flux integration, temperature stepping and the post-processing diagnostics.
// From-scratch sm_100a kernels for K:2428-3139 of the reference.
#include "common.cuh"
#include "comm.cuh"

// ------------------------------------------------------------------------------------------------
// integrate_flux (K:2428-2513).  The reference runs ONE 1024-thread block that funnels every cell
// through fp64 CAS-loop atomics.  Here:
//   grid (x-tiles, interfaces, atmospheres): a block stages XB*ny contiguous doubles of each of the three wg arrays
//   through shared memory (coalesced), one thread per bin adds its ny Gauss points in y order -> F_*_band[i][x];
//   the block then tree-sums its bins' contributions to the wavelength integral, and the last block of an
//   interface to finish adds the per-block partial sums in block order -> F_up_tot, F_down_tot, F_net.
// One launch; the summation order is fixed, so results are bitwise reproducible run to run.
// ------------------------------------------------------------------------------------------------
#define IF_THREADS 256
#ifndef IF_YB
#define IF_YB 4  // Gauss points whose loads are in flight together
#endif
#define IF_BLOCKS_PER_SM 5
#define IF_MAXW 64  // Gauss points per bin whose half weights are kept in shared memory

#ifdef HELIOS_INTEG_TIMING  // experiment builds only (scripts/exp_integ_timing.py): phase stamps of one block, SM cycles
__device__ long long g_integ_t[8];
#define ISTAMP(k) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 50 && blockIdx.z == 0) g_integ_t[k] = clock64(); } while (0)
extern "C" int helios_debug_integ_timing(long long* out8) {
    return cudaMemcpyFromSymbol(out8, g_integ_t, sizeof(g_integ_t)) == cudaSuccess ? 0 : 1;
}
#else
#define ISTAMP(k) do { } while (0)
#endif

__device__ __forceinline__ double block_sum(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// 5 blocks per SM (<= 51 registers): the C2 grid of 6 x 101 blocks is then ONE wave (740 slots); at 58 registers it was
// 592 slots and a second wave of 14 blocks doubled the kernel's duration
template <bool NY1>  // NY1: one point per bin (opacity sampling), no staging, next tile prefetched
__global__ void __launch_bounds__(IF_THREADS, IF_BLOCKS_PER_SM)
k_band_integrate(const double* __restrict__ F_down_wg, const double* __restrict__ F_up_wg,
                 const double* __restrict__ F_dir_wg, double* __restrict__ F_down_band,
                 double* __restrict__ F_up_band, double* __restrict__ F_dir_band,
                 const double* __restrict__ gauss_weight, int nbin, int ny, int xb,
                 const double* __restrict__ deltalambda, double* __restrict__ partial, unsigned* __restrict__ ticket,
                 double* __restrict__ F_down_tot, double* __restrict__ F_up_tot, double* __restrict__ F_net, FusedComm fc) {
    extern __shared__ double sm[];
    ISTAMP(0);
    ISTAMP(1);
    __shared__ double s_hw[IF_MAXW];  // 0.5 * gauss_weight[y]
    if ((int)threadIdx.x < ny && ny <= IF_MAXW) s_hw[threadIdx.x] = 0.5 * gauss_weight[threadIdx.x];  // (barrier: after staging)
    const int pitch = ny + 1;  // odd pitch keeps the per-bin reads off one bank
    double* s_dn = sm;
    double* s_up = sm + (size_t)xb * pitch;
    double* s_dr = sm + (size_t)2 * xb * pitch;
    const int i = blockIdx.y;
    {   // batch (blockIdx.z = atmosphere): [i][x][y] and [i][x] arrays both hold gridDim.y = ninterface rows
        const size_t wg = (size_t)blockIdx.z * gridDim.y * nbin * ny, bd = (size_t)blockIdx.z * gridDim.y * nbin;
        F_down_wg += wg; F_up_wg += wg; F_dir_wg += wg;
        F_down_band += bd; F_up_band += bd; F_dir_band += bd;
    }
    // A block walks the x-tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its interface (wide spectra: 1e5 bins are
    // ~400 tiles; a block per tile would be ~40,000 blocks of two barriers-and-a-ticket each) and keeps its share of the
    // wavelength integral in registers across them.
    double t_up = 0.0, t_dn = 0.0;  // this thread's bins in the sum over wavelength
    const int ntiles = (nbin + xb - 1) / xb;
    // ny == 1 (opacity sampling, one thread per bin, no staging): the values of the block's NEXT tile are requested before
    // the current one is consumed -- a block walks up to ~60 tiles, and each used to wait for its own loads
    double p_dr = 0.0, p_up = 0.0, p_dn = 0.0, p_dl = 0.0;
    auto fetch1 = [&](int tile) {
        const int x = tile * xb + (int)threadIdx.x;
        if (tile < ntiles && (int)threadIdx.x < xb && x < nbin) {
            const size_t k = (size_t)i * nbin + x;
            p_dr = F_dir_wg[k];
            p_up = F_up_wg[k];
            p_dn = F_down_wg[k];
            p_dl = deltalambda[x];
        }
    };
    if (NY1) fetch1(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int x0 = tile * xb;
        const int nx = min(xb, nbin - x0);
        const size_t base = ((size_t)i * nbin + x0) * ny;
        const double c_dr = p_dr, c_up = p_up, c_dn = p_dn, c_dl = p_dl;
        if (NY1) fetch1(tile + gridDim.x);
        if (!NY1) {
            const int n = nx * ny;
            // two rounds of loads in flight per thread before the first is consumed (more would spill at the register budget
            // of IF_BLOCKS_PER_SM): the staging is a handful of
            // round trips to HBM, and a loop that stores each value as it arrives pays every one of them in full
            constexpr int U = 2;
            for (int k0 = threadIdx.x; k0 < n; k0 += blockDim.x * U) {
                double v_dn[U], v_up[U], v_dr[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = k0 + u * blockDim.x;
                    if (k < n) {
                        v_dn[u] = F_down_wg[base + k];
                        v_up[u] = F_up_wg[base + k];
                        v_dr[u] = F_dir_wg[base + k];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = k0 + u * blockDim.x;
                    if (k < n) {
                        const int xl = k / ny, y = k - xl * ny;
                        const int d = xl * pitch + y;
                        s_dn[d] = v_dn[u];
                        s_up[d] = v_up[u];
                        s_dr[d] = v_dr[u];
                    }
                }
            }
            __syncthreads();
            ISTAMP(2);
        }
        for (int xl = threadIdx.x; xl < nx; xl += blockDim.x) {
            double a_dn = 0.0, a_up = 0.0, a_dr = 0.0;
            if (!NY1) {
                // half weights from shared memory (a global load per Gauss point sat in the dependent chain: 135 cycles
                // per point, phase stamps of scripts/exp_integ_timing.py), four points in flight; same operations, same order
                const double* __restrict__ r_dr = s_dr + xl * pitch;
                const double* __restrict__ r_up = s_up + xl * pitch;
                const double* __restrict__ r_dn = s_dn + xl * pitch;
                int y = 0;
                if (ny <= IF_MAXW) {
                    for (; y + IF_YB <= ny; y += IF_YB) {  // the loads of IF_YB points first, then the adds in y order
                        double w[IF_YB], v_dr[IF_YB], v_up[IF_YB], v_dn[IF_YB];
#pragma unroll
                        for (int u = 0; u < IF_YB; u++) {
                            w[u] = s_hw[y + u];
                            v_dr[u] = r_dr[y + u];
                            v_up[u] = r_up[y + u];
                            v_dn[u] = r_dn[y + u];
                        }
#pragma unroll
                        for (int u = 0; u < IF_YB; u++) {
                            a_dr += w[u] * v_dr[u];
                            a_up += w[u] * v_up[u];
                            a_dn += w[u] * v_dn[u];
                        }
                    }
                }
                for (; y < ny; y++) {
                    const double hw = ny <= IF_MAXW ? s_hw[y] : 0.5 * gauss_weight[y];
                    a_dr += hw * r_dr[y];
                    a_up += hw * r_up[y];
                    a_dn += hw * r_dn[y];
                }
            } else {  // opacity sampling: one point per bin, consecutive threads read consecutive bins (xl == threadIdx.x)
                const double hw = 0.5 * gauss_weight[0];
                a_dr += hw * c_dr;
                a_up += hw * c_up;
                a_dn += hw * c_dn;
            }
            const size_t o = (size_t)i * nbin + x0 + xl;
            F_dir_band[o] = a_dr;
            F_up_band[o] = a_up;
            F_down_band[o] = a_dn;
            const double dl = NY1 ? c_dl : deltalambda[x0 + xl];
            t_up += a_up * dl;
            t_dn += (a_dr + a_dn) * dl;
        }
        if (!NY1) __syncthreads();  // the next tile overwrites the staging area
        ISTAMP(3);
    }
    // Sum over wavelength in the same launch (K:2484-2509): fixed tree over this block's bins, then the LAST block
    // of the interface to finish adds the per-block partial sums in block order -- a fixed summation order, bitwise
    // reproducible run to run, and no second launch.
    // fixed-shape reduction: shuffle tree inside each warp, then warp 0 adds the 8 warp sums in warp order
    __shared__ double wsum[2][IF_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_up += __shfl_down_sync(0xffffffffu, t_up, d);
        t_dn += __shfl_down_sync(0xffffffffu, t_dn, d);
    }
    if ((threadIdx.x & 31) == 0) {
        wsum[0][threadIdx.x >> 5] = t_up;
        wsum[1][threadIdx.x >> 5] = t_dn;
    }
    __shared__ bool last;
    const int ntile = gridDim.x, nint = gridDim.y;
    const size_t slot = (size_t)blockIdx.z * nint + i;  // (atmosphere, interface)
    __syncthreads();
    if (threadIdx.x == 0) {
        t_up = t_dn = 0.0;
#pragma unroll
        for (int w = 0; w < IF_THREADS / 32; w++) {
            t_up += wsum[0][w];
            t_dn += wsum[1][w];
        }
        partial[(slot * ntile + blockIdx.x) * 2] = t_up;
        partial[(slot * ntile + blockIdx.x) * 2 + 1] = t_dn;
        ISTAMP(4);
        __threadfence();
        last = atomicAdd(ticket + slot, 1u) == (unsigned)ntile - 1;
        ISTAMP(5);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const volatile double* p = partial + slot * ntile * 2;
        double up = 0.0, dn = 0.0;
        for (int c = 0; c < ntile; c++) {
            up += p[2 * c];
            dn += p[2 * c + 1];
        }
        F_up_tot[slot] = up;
        F_down_tot[slot] = dn;
        F_net[slot] = up - dn;
        ticket[slot] = 0u;
    }
    ISTAMP(6);
    // Wavelength sharding, fused form (helios_comm_set_fused): the sum over ranks of the per-interface totals runs in THIS
    // launch, with no fence and no flag.  Every double travels as two 8-byte packets {32 data bits, round number}
    // (comm.cuh: ll_store / ll_load): the block that finishes an interface stores this rank's two totals straight into
    // slot `rank` of every mailbox over NVLink (thread r -> peer r); the block that finishes the LAST interface of this
    // rank polls its own mailbox until every packet of the round has arrived -- a packet is valid when it carries the
    // round number -- and adds the world's slots in rank order: every rank obtains bitwise the same totals, one NVLink
    // latency after the slowest rank's integration.
    if (fc.world > 0) {
        __shared__ bool all_done;
        __shared__ unsigned round_s;
        __shared__ double tot_s[2];
        if (threadIdx.x == 0) {
            all_done = false;
            // read before this block's ticket: the counter only advances after the last ticket of the launch
            round_s = (unsigned)(*fc.seq_dev + 1ull);
            if (last) {
                tot_s[0] = F_up_tot[i];
                tot_s[1] = F_down_tot[i];
            }
        }
        __syncthreads();
        const unsigned round = round_s;
        const int bank = (int)(round & 1u);
        if (last) {
            if ((int)threadIdx.x < fc.world) {
                uint4* box = reinterpret_cast<uint4*>(reinterpret_cast<char*>(fc.peers.p[threadIdx.x]) + fc.ll_off) +
                             ((size_t)bank * fc.world + fc.rank) * fc.slot;
                ll_store(box + i, tot_s[0], round);
                ll_store(box + nint + i, tot_s[1], round);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();  // this block's F_*_tot[i] before the final block's sums (same device)
                all_done = atomicAdd(fc.ticket, 1u) == (unsigned)nint - 1;
            }
        }
        __syncthreads();
        if (all_done) {
            const uint4* box = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(fc.peers.p[fc.rank]) + fc.ll_off) +
                               (size_t)bank * fc.world * fc.slot;
            for (int t = threadIdx.x; t < nint; t += blockDim.x) {
                double a = 0.0, b = 0.0;
                for (int r = 0; r < fc.world; r++) {
                    a += ll_load(box + (size_t)r * fc.slot + t, round);
                    b += ll_load(box + (size_t)r * fc.slot + nint + t, round);
                }
                F_up_tot[t] = a;
                F_down_tot[t] = b;
                F_net[t] = a - b;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                *fc.ticket = 0u;
                *fc.seq_dev += 1ull;  // the next exchanging launch (stream order) sees the advanced counter
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Temperature stepping (K:2606-2884).  One block; the reads of neighbouring temperatures, the
// smoothing prefix sums and the temperature update are separated by block barriers (the reference
// races across its 16-thread blocks when smoothing is on, K:2665-2669).
// ------------------------------------------------------------------------------------------------
struct TempScalars {
    int itervalue, foreplay, numlayers, adapt_interval, smooth, dim, step, no_atmo, conv;
    double f_factor, g, physical_tstep, local_limit, F_intern;
    const int* done;        // batch: atmospheres that have converged are left untouched
    const double* g_batch;  // batch: per-atmosphere gravity
    const int* iter_dev;    // batch with the on-device iteration counter: overrides itervalue
    // fused tail (helios_rad_temp_iter_latched): flag sum + convergence latch + counter advance in the same launch
    int* sums;
    int* done_w;
    int* converged_at;
    int* iter_w;
    unsigned* ticket;
};

__global__ void __launch_bounds__(256)
k_temp_iter(const double* __restrict__ F_down_tot, const double* __restrict__ F_net,
            double* __restrict__ F_net_diff, double* __restrict__ tlay, const double* __restrict__ play,
            const double* __restrict__ pint, int* __restrict__ abrt, double* __restrict__ T_store,
            double* __restrict__ prefactor, const int* __restrict__ marked_red,
            const double* __restrict__ F_add_heat_lay, const double* __restrict__ F_add_heat_sum,
            double* __restrict__ F_smooth, double* __restrict__ F_smooth_sum,
            const double* __restrict__ c_p_lay, const double* __restrict__ mmm_lay, TempScalars s) {
    const int nl = s.numlayers;
    bool skip = false;  // batch: this atmosphere has converged and is left untouched
    if (s.iter_dev != nullptr) s.itervalue = *s.iter_dev;
    if (gridDim.x > 1 || s.done != nullptr) {  // batch (blockIdx.x = atmosphere)
        const size_t a = blockIdx.x;
        skip = s.done != nullptr && s.done[a] != 0;
        if (s.g_batch != nullptr) s.g = s.g_batch[a];
        const size_t v1 = a * (nl + 1), v0 = a * nl;  // vectors with nlayer + 1 / nlayer entries
        if (F_down_tot) F_down_tot += v1;
        F_net += v1; tlay += v1; pint += v1; T_store += v1; prefactor += v1;
        if (abrt) abrt += v1;
        if (marked_red) marked_red += v1;
        F_net_diff += v0; play += v0; F_add_heat_lay += v0; F_smooth += v0; F_smooth_sum += v0;
        if (F_add_heat_sum) F_add_heat_sum += v0;
        if (c_p_lay) c_p_lay += v0;
        if (mmm_lay) mmm_lay += v0;
    }
    int my_flags = 0;  // convergence flags this thread has just written (the fused tail sums them without re-reading)
    if (!skip) {
    // phase 1: flux divergence and smoothing force, from the OLD temperatures
    for (int i = threadIdx.x; i < nl; i += blockDim.x) {
        F_net_diff[i] = F_net[i] - F_net[i + 1] + F_add_heat_lay[i];
        if (s.smooth == 1) {
            double t_mid = tlay[i];
            if (play[i] < 1e6 && i < nl - 1 && i > 0) t_mid = (tlay[i - 1] + tlay[i + 1]) / 2.0;
            F_smooth[i] = pow(t_mid - tlay[i], 7.0);
        }
    }
    __syncthreads();
    if (s.smooth == 1) {
        // running sum in layer order (K:2668-2669), one thread: nl is ~100
        if (threadIdx.x == 0) {
            double acc = 0.0;
            for (int j = 0; j < nl; j++) {
                acc += F_smooth[j];
                F_smooth_sum[j] = acc;
            }
        }
        __syncthreads();
    }
    // phase 2: step every layer and the surface "ghost layer" i = nl
    for (int i = threadIdx.x; i < nl + 1; i += blockDim.x) {
        double combined;
        if (i < nl) {
            combined = F_net_diff[i] + F_smooth[i];
        } else {
            combined = s.F_intern - F_net[0];
            if (s.conv == 0) {
                if (fabs(s.F_intern - F_net[1]) / (F_down_tot[nl] + s.F_intern) > 0.5 * s.local_limit)
                    combined = s.F_intern - F_net[1];
            } else {
                for (int j = 0; j < nl; j++) {
                    if (marked_red[j] == 1) {
                        combined = s.F_intern - F_net[j + 1];
                        break;
                    }
                }
            }
        }
        double delta_T = 0.0;
        const double T_old = tlay[i];
        if (s.conv == 0) {
            if (s.physical_tstep == 0) {
                if (s.itervalue == s.foreplay) prefactor[i] = 1e0;
                if (s.itervalue == 10000) prefactor[i] = 1e-1;
                double delta_t = 0.0;  // the reference leaves it uninitialised when combined == 0
                if (combined != 0) delta_t = prefactor[i] * play[0] / pow(fabs(combined), 0.9);
                delta_T = combined / (pint[0] - pint[1]) * delta_t;
                if (fabs(delta_T) > 500.0) delta_T = 500.0 * combined / fabs(combined);
                if (s.itervalue % s.adapt_interval == 0) T_store[i] = T_old;
                if (s.itervalue % s.adapt_interval == s.adapt_interval - 1) {
                    if (fabs(T_old - T_store[i]) < s.adapt_interval / 2.0 * fabs(delta_T)) prefactor[i] /= 1.5;
                    else prefactor[i] *= 1.1;
                }
            } else {
                const double delta_t = s.physical_tstep;
                const int k = i < nl ? i : 0;
                delta_T = s.g / (c_p_lay[k] / (mmm_lay[k] / hc::AMU)) * combined / (pint[k] - pint[k + 1]) * delta_t;
            }
            double T_new = T_old + delta_T;
            if (s.no_atmo == 1 && i != nl) T_new = 1.001;
            const double max_limit = s.dim * s.step - 1.001;
            T_new = fmin(fmax(T_new, 1.001), max_limit);
            tlay[i] = T_new;
            bool ok;
            if (i < nl)
                ok = fabs(s.F_intern + F_add_heat_sum[i] + F_smooth_sum[i] - F_net[i + 1]) /
                         (F_down_tot[nl] + s.F_intern) < s.local_limit;
            else
                ok = fabs(s.F_intern - F_net[0]) / (F_down_tot[nl] + s.F_intern) < s.local_limit;
            abrt[i] = ok ? 1 : 0;
            my_flags += ok ? 1 : 0;
        } else {
            if (s.itervalue == 0) prefactor[i] = 1e-2;
            if (s.itervalue == 6000) prefactor[i] = 1e-3;
            double delta_t = 0.0;
            if (combined != 0) delta_t = prefactor[i] * play[0] / pow(fabs(combined), 0.5);
            delta_T = combined / (pint[0] - pint[1]) * delta_t;
            if (fabs(delta_T) > 20.0) delta_T = 20.0 * combined / fabs(combined);
            if (s.itervalue % s.adapt_interval == 0) T_store[i] = T_old;
            if (s.itervalue % s.adapt_interval == s.adapt_interval - 1) {
                if (fabs(T_old - T_store[i]) < s.adapt_interval / 2.0 * fabs(delta_T)) prefactor[i] /= 1.5;
                else prefactor[i] *= 1.1;
            }
            tlay[i] = fmax(T_old + delta_T, 1.001);
        }
    }
    }  // !skip
    if (s.sums != nullptr) {
        // fused tail: what k_abort_sum + k_iter_advance do, without two more launches per iteration
        __shared__ int red_i[256];
        __syncthreads();
        int a = my_flags;
        if (skip || s.conv != 0) {  // flags not rewritten by this launch: read them
            a = 0;
            for (int i = threadIdx.x; i < nl + 1; i += blockDim.x) a += abrt[i];
        }
        red_i[threadIdx.x] = a;
        __syncthreads();
        for (int w = blockDim.x / 2; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red_i[threadIdx.x] += red_i[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int b = blockIdx.x;
            s.sums[b] = red_i[0];
            if (red_i[0] == nl + 1 && s.done_w[b] == 0) {
                s.done_w[b] = 1;
                s.converged_at[b] = s.itervalue + 1;  // iterations completed
            }
            // the LAST block to finish advances the iteration counter: by then every block has read it
            if (gridDim.x == 1) {
                *s.iter_w = s.itervalue + 1;
            } else {
                __threadfence();
                if (atomicAdd(s.ticket, 1u) == gridDim.x - 1) {
                    *s.ticket = 0u;
                    *s.iter_w = s.itervalue + 1;
                }
            }
        }
    }
}

// conv_temp_iter's smoothing branch differs slightly (no i > 0 guard in the reference, K:2808, which
// reads tlay[-1]); the guarded form is used for both.

__global__ void k_abort_sum(const int* __restrict__ abrt, int n, int* __restrict__ out, int* __restrict__ done,
                            int* __restrict__ converged_at, const int* __restrict__ iter_dev) {
    __shared__ int red[256];
    abrt += (size_t)blockIdx.x * n;  // batch (blockIdx.x = atmosphere)
    out += blockIdx.x;
    int a = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += abrt[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = red[0];
        if (done != nullptr && red[0] == n && done[blockIdx.x] == 0) {  // latched: see helios_ctx_set_batch
            done[blockIdx.x] = 1;
            if (iter_dev != nullptr) converged_at[blockIdx.x] = *iter_dev + 1;  // iterations completed
        }
    }
}

// ------------------------------------------------------------------------------------------------
// post-processing diagnostics (K:2888-3139)
// ------------------------------------------------------------------------------------------------
template <bool NONISO>
__global__ void k_optdepth_trans(const double* __restrict__ tr_a, const double* __restrict__ tr_b,
                                 double* __restrict__ trans_band, const double* __restrict__ dt_a,
                                 const double* __restrict__ dt_b, double* __restrict__ dtau_band,
                                 const double* __restrict__ gw, double* __restrict__ dtc,
                                 const double* __restrict__ dtc_u, const double* __restrict__ dtc_l, int nbin,
                                 int nlayer, int ny) {
    const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (o >= (long long)nbin * nlayer) return;
    double dsum = 0.0, tsum = 0.0;
    const size_t base = (size_t)o * ny;
    for (int y = 0; y < ny; y++) {
        if (NONISO) {
            dsum += 0.5 * gw[y] * (dt_a[base + y] + dt_b[base + y]);
            tsum += 0.5 * gw[y] * (tr_a[base + y] * tr_b[base + y]);
        } else {
            dsum += 0.5 * gw[y] * dt_a[base + y];
            tsum += 0.5 * gw[y] * tr_a[base + y];
        }
    }
    dtau_band[o] = dsum;
    trans_band[o] = tsum;
    if (NONISO) dtc[o] = dtc_l[o] + dtc_u[o];
}

// contribution function: one thread per bin walks every Gauss column from the top, carrying the
// transmission to TOA as a running product (the reference rebuilds that product for every layer,
// O(nlayer^2 ny) per thread, K:2970-2979).
template <bool NONISO>
__global__ void k_contr_func(const double* __restrict__ tr_a, const double* __restrict__ tr_b,
                             double* __restrict__ tw_band, double* __restrict__ cf_band,
                             const double* __restrict__ gw, const double* __restrict__ planck_lay,
                             double epsi, int nbin, int nlayer, int ny) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nbin) return;
    for (int y = 0; y < ny; y++) {
        double to_top = 1.0;
        for (int i = nlayer - 1; i >= 0; i--) {
            const size_t e = (size_t)y + (size_t)ny * x + (size_t)ny * nbin * i;
            const double t = NONISO ? tr_a[e] * tr_b[e] : tr_a[e];
            tw_band[x + (size_t)nbin * i] += 0.5 * gw[y] * (1.0 - t) * to_top;
            to_top *= t;
        }
    }
    for (int i = 0; i < nlayer; i++)
        cf_band[x + (size_t)nbin * i] =
            2.0 * hc::PI * epsi * planck_lay[i + (size_t)x * (nlayer + 2)] * tw_band[x + (size_t)nbin * i];
}

// K:294-329
__device__ __forceinline__ double dB_dT(double lambda, double T) {
    const double c3 = hc::CSPEED * hc::CSPEED * hc::CSPEED;
    const double l2 = lambda * lambda;
    const double l6 = l2 * l2 * l2;
    const double D = 2.0 * hc::HCONST * c3 * hc::HCONST / (l6 * hc::KBOLTZMANN * (T * T));
    const double ex = exp(hc::HCONST * hc::CSPEED / (lambda * hc::KBOLTZMANN * T));
    return D * ex / ((ex - 1.0) * (ex - 1.0));
}

__device__ __forceinline__ double integrated_dB_dT(const double* __restrict__ kw, const double* __restrict__ ky,
                                                   int ny, double lbot, double ltop, double T) {
    double r = 0.0;
    for (int y = 0; y < ny; y++) {
        const double xx = (ky[y] - 0.5) * 2.0;
        const double arg = (ltop - lbot) / 2.0 * xx + (ltop + lbot) / 2.0;
        r += (ltop - lbot) / 2.0 * kw[y] * dB_dT(arg, T);
    }
    return r;
}

// one block per layer; the eight spectral sums are reduced in a fixed tree over the bins
__global__ void __launch_bounds__(IF_THREADS)
k_mean_opacities(double* __restrict__ planck_pl, double* __restrict__ ross_pl, double* __restrict__ planck_st,
                 double* __restrict__ ross_st, const double* __restrict__ opac_wg_lay,
                 const double* __restrict__ abs_cl, const double* __restrict__ mmm,
                 const double* __restrict__ planck_lay, const double* __restrict__ interwave,
                 const double* __restrict__ deltawave, const double* __restrict__ T_lay,
                 const double* __restrict__ gw, const double* __restrict__ gy, double* __restrict__ opac_band,
                 int nlayer, int nbin, int ny, double T_star) {
    __shared__ double red[IF_THREADS];
    const int i = blockIdx.x;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double Ti = T_lay[i];
    for (int x = threadIdx.x; x < nbin; x += blockDim.x) {
        double band = 0.0;
        const size_t base = ((size_t)i * nbin + x) * ny;
        for (int y = 0; y < ny; y++) band += 0.5 * gw[y] * opac_wg_lay[base + y];
        opac_band[x + (size_t)nbin * i] = band;
        const double k = band + abs_cl[x + (size_t)nbin * i] / mmm[i];
        const double Bp = planck_lay[i + (size_t)x * (nlayer + 2)] * deltawave[x];
        const double Bs = planck_lay[nlayer + (size_t)x * (nlayer + 2)] * deltawave[x];
        const double dpl = integrated_dB_dT(gw, gy, ny, interwave[x], interwave[x + 1], Ti);
        const double dst = integrated_dB_dT(gw, gy, ny, interwave[x], interwave[x + 1], T_star);
        a[0] += k * Bp;  // planckband * deltawave grouped; agrees with K:3071 to rounding
        a[1] += Bp;
        a[2] += dpl;
        if (k > 0) a[3] += dpl / k;
        a[4] += k * Bs;
        a[5] += Bs;
        a[6] += dst;
        if (k > 0) a[7] += dst / k;
    }
    for (int q = 0; q < 8; q++) a[q] = block_sum(a[q], red);
    if (threadIdx.x == 0) {
        planck_pl[i] = a[0] / a[1];
        ross_pl[i] = a[2] / a[3];
        if (Ti < 70) ross_pl[i] = -3;
        planck_st[i] = a[4] / a[5];
        ross_st[i] = a[6] / a[7];
        if (T_star < 70) {
            planck_st[i] = -3;
            ross_st[i] = -3;
        }
    }
}

__global__ void __launch_bounds__(IF_THREADS)
k_integrate_beamflux(double* __restrict__ F_dir_tot, const double* __restrict__ F_dir_band,
                     const double* __restrict__ deltalambda, int nbin) {
    __shared__ double red[IF_THREADS];
    const int i = blockIdx.x;
    double acc = 0.0;
    for (int x = threadIdx.x; x < nbin; x += blockDim.x) acc += F_dir_band[x + (size_t)nbin * i] * deltalambda[x];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) F_dir_tot[i] = acc;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

int helios_integrate_flux_double(helios_ctx* ctx, const double* deltalambda, double* F_down_tot,
                                 double* F_up_tot, double* F_net, const double* F_down_wg,
                                 const double* F_up_wg, const double* F_dir_wg, double* F_down_band,
                                 double* F_up_band, double* F_dir_band, const double* gauss_weight,
                                 int nbin, int numinterfaces, int ny) {
    HCTX(ctx);
    HARG(deltalambda && F_down_tot && F_up_tot && F_net && F_down_wg && F_up_wg && F_dir_wg &&
         F_down_band && F_up_band && F_dir_band && gauss_weight);
    HARG(nbin > 0 && numinterfaces > 0 && ny > 0);
    // bins per block: as many as fit in ~36 kB of shared memory, at most 256
    int xb = (int)(36 * 1024 / (3 * sizeof(double) * (ny + 1)));
    if (xb > 256) xb = 256;
    if (xb < 1) {
        helios_set_error("helios_integrate_flux_double: ny = %d too large", ny);
        return HELIOS_ERR_ARG;
    }
    const size_t smem = (size_t)3 * xb * (ny + 1) * sizeof(double);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    const int nb = ctx->batch.nbatch;
    // blocks per (atmosphere, interface): one per x-tile while that keeps a single atmosphere's grid within IF_BLOCKS_PER_SM blocks per
    // SM, else a block walks several tiles.  Independent of the batch size: the summation order of an atmosphere is the
    // same alone and in a batch (bit-for-bit equality of the two paths, tests/test_gpu_batch.py).
    int ntile = ceil_div(nbin, xb);
    const int cap = ctx->num_sms * IF_BLOCKS_PER_SM / numinterfaces;  // one wave
    if (ntile > cap) ntile = cap < 1 ? 1 : cap;
    dim3 grid(ntile, numinterfaces, nb);
    // per-block partial sums (transient, scratch) and one ticket per (atmosphere, interface) (persistent, zeroed)
    double* scratch = nullptr;
    const size_t pbytes = (size_t)nb * numinterfaces * ntile * 2 * sizeof(double);
    int rc = helios_ctx_scratch(ctx, pbytes + 64, &scratch);
    if (rc) return rc;
    const size_t nticket = (size_t)nb * numinterfaces;
    if (ctx->integ_ticket_n < nticket) {
        if (ctx->integ_ticket) {
            HCUDA(cudaStreamSynchronize(ctx->stream));
            HCUDA(cudaFree(ctx->integ_ticket));
            ctx->integ_ticket = nullptr;
            ctx->integ_ticket_n = 0;
        }
        HCUDA(cudaMalloc((void**)&ctx->integ_ticket, nticket * sizeof(unsigned)));
        HCUDA(cudaMemsetAsync(ctx->integ_ticket, 0, nticket * sizeof(unsigned), ctx->stream));
        ctx->integ_ticket_n = nticket;
    }
    FusedComm fc;
    rc = helios_comm_fused_next(ctx, numinterfaces, &fc);
    if (rc) return rc;
    if (fc.world > 0 && nb != 1) {
        helios_set_error("helios_integrate_flux_double: the fused flux all-reduce is not available in batch mode");
        return HELIOS_ERR_STATE;
    }
    if (ny == 1)
        k_band_integrate<true><<<grid, IF_THREADS, smem, ctx->stream>>>(F_down_wg, F_up_wg, F_dir_wg, F_down_band, F_up_band,
                                                                        F_dir_band, gauss_weight, nbin, ny, xb, deltalambda,
                                                                        scratch + 8, ctx->integ_ticket, F_down_tot,
                                                                        F_up_tot, F_net, fc);
    else
        k_band_integrate<false><<<grid, IF_THREADS, smem, ctx->stream>>>(F_down_wg, F_up_wg, F_dir_wg, F_down_band, F_up_band,
                                                                         F_dir_band, gauss_weight, nbin, ny, xb, deltalambda,
                                                                         scratch + 8, ctx->integ_ticket, F_down_tot,
                                                                         F_up_tot, F_net, fc);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

static int rad_temp_iter_impl(helios_ctx* ctx, const double* F_down_tot, const double* F_net, double* F_net_diff,
                              double* tlay, const double* play, const double* pint, int* abrt, double* T_store,
                              double* deltat_prefactor, const double* F_add_heat_lay, const double* F_add_heat_sum,
                              double* F_smooth, double* F_smooth_sum, const double* c_p_lay,
                              const double* meanmolmass_lay, int itervalue, double f_factor, int foreplay, double g,
                              int numlayers, double physical_tstep, double local_limit, int adapt_interval, int smooth,
                              int dim, int step, double F_intern, int no_atmo, int* sum_dev, const char* who) {
    if (!(F_down_tot && F_net && F_net_diff && tlay && play && pint && abrt && T_store && deltat_prefactor &&
          F_add_heat_lay && F_add_heat_sum && F_smooth && F_smooth_sum) ||
        !(physical_tstep == 0 || (c_p_lay != nullptr && meanmolmass_lay != nullptr)) ||
        !(numlayers > 1 && adapt_interval > 0)) {
        helios_set_error("%s: invalid argument", who);
        return HELIOS_ERR_ARG;
    }
    HBATCHDIMS(ctx, numlayers == ctx->batch.nlayer);
    const BatchDesc& bd = ctx->batch;
    const bool batched = bd.active;
    if (sum_dev != nullptr && !(batched && bd.use_iter_dev)) {
        helios_set_error("%s: needs batch mode with the device iteration counter", who);
        return HELIOS_ERR_STATE;
    }
    TempScalars s{itervalue, foreplay, numlayers, adapt_interval, smooth, dim, step, no_atmo, 0,
                  f_factor, g, physical_tstep, local_limit, F_intern,
                  batched ? bd.done : nullptr, batched ? bd.g : nullptr,
                  (batched && bd.use_iter_dev) ? bd.iter_dev : nullptr,
                  sum_dev, bd.done, bd.converged_at, bd.iter_dev, bd.ticket};
    k_temp_iter<<<bd.nbatch, 256, 0, ctx->stream>>>(F_down_tot, F_net, F_net_diff, tlay, play, pint, abrt, T_store,
                                                    deltat_prefactor, nullptr, F_add_heat_lay, F_add_heat_sum,
                                                    F_smooth, F_smooth_sum, c_p_lay, meanmolmass_lay, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_rad_temp_iter(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                         const double* F_net, double* F_net_diff, double* tlay, const double* play,
                         const double* tint, const double* pint, int* abrt, double* T_store,
                         double* deltat_prefactor, const double* F_add_heat_lay,
                         const double* F_add_heat_sum, double* F_smooth, double* F_smooth_sum,
                         const double* c_p_lay, const double* meanmolmass_lay, int itervalue,
                         double f_factor, int foreplay, double g, int numlayers, double physical_tstep,
                         double local_limit, int adapt_interval, int smooth, int dim, int step,
                         double F_intern, int no_atmo) {
    HCTX(ctx);
    (void)F_up_tot; (void)tint;
    return rad_temp_iter_impl(ctx, F_down_tot, F_net, F_net_diff, tlay, play, pint, abrt, T_store, deltat_prefactor,
                              F_add_heat_lay, F_add_heat_sum, F_smooth, F_smooth_sum, c_p_lay, meanmolmass_lay,
                              itervalue, f_factor, foreplay, g, numlayers, physical_tstep, local_limit,
                              adapt_interval, smooth, dim, step, F_intern, no_atmo, nullptr, "helios_rad_temp_iter");
}

int helios_rad_temp_iter_latched(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                                 const double* F_net, double* F_net_diff, double* tlay, const double* play,
                                 const double* tint, const double* pint, int* abrt, double* T_store,
                                 double* deltat_prefactor, const double* F_add_heat_lay,
                                 const double* F_add_heat_sum, double* F_smooth, double* F_smooth_sum,
                                 const double* c_p_lay, const double* meanmolmass_lay, int itervalue,
                                 double f_factor, int foreplay, double g, int numlayers, double physical_tstep,
                                 double local_limit, int adapt_interval, int smooth, int dim, int step,
                                 double F_intern, int no_atmo, int* sum_dev) {
    HCTX(ctx);
    (void)F_up_tot; (void)tint;
    HARG(sum_dev != nullptr);
    return rad_temp_iter_impl(ctx, F_down_tot, F_net, F_net_diff, tlay, play, pint, abrt, T_store, deltat_prefactor,
                              F_add_heat_lay, F_add_heat_sum, F_smooth, F_smooth_sum, c_p_lay, meanmolmass_lay,
                              itervalue, f_factor, foreplay, g, numlayers, physical_tstep, local_limit,
                              adapt_interval, smooth, dim, step, F_intern, no_atmo, sum_dev,
                              "helios_rad_temp_iter_latched");
}

int helios_conv_temp_iter(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                          const double* F_net, double* F_net_diff, double* tlay, const double* play,
                          const double* pint, double* T_store, double* deltat_prefactor,
                          const int* marked_red, const double* F_add_heat_lay, double* F_smooth,
                          double* F_smooth_sum, int numlayers, int itervalue, int adapt_interval,
                          int smooth, double F_intern) {
    HCTX(ctx);
    (void)F_up_tot; (void)F_down_tot;
    HARG(F_net && F_net_diff && tlay && play && pint && T_store && deltat_prefactor && marked_red &&
         F_add_heat_lay && F_smooth && F_smooth_sum);
    HARG(numlayers > 1 && adapt_interval > 0);
    HNOBATCH(ctx);
    TempScalars s{itervalue, 0, numlayers, adapt_interval, smooth, 0, 0, 0, 1, 0.0, 0.0, 0.0, 0.0, F_intern,
                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    k_temp_iter<<<1, 256, 0, ctx->stream>>>(nullptr, F_net, F_net_diff, tlay, play, pint, nullptr, T_store,
                                            deltat_prefactor, marked_red, F_add_heat_lay, nullptr, F_smooth,
                                            F_smooth_sum, nullptr, nullptr, s);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_abort_sum(helios_ctx* ctx, const int* abrt, int n, int* sum_dev) {
    HCTX(ctx);
    HARG(abrt && sum_dev && n > 0);
    HBATCHDIMS(ctx, n == ctx->batch.nlayer + 1);
    const BatchDesc& bd = ctx->batch;
    k_abort_sum<<<bd.nbatch, 256, 0, ctx->stream>>>(abrt, n, sum_dev, bd.active ? bd.done : nullptr, bd.converged_at,
                                                    (bd.active && bd.use_iter_dev) ? bd.iter_dev : nullptr);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

__global__ void k_iter_advance(int* iter_dev) { iter_dev[0] += 1; }

int helios_batch_iter_advance(helios_ctx* ctx) {
    HCTX(ctx);
    if (!ctx->batch.active) {
        helios_set_error("helios_batch_iter_advance: not in batch mode");
        return HELIOS_ERR_STATE;
    }
    k_iter_advance<<<1, 1, 0, ctx->stream>>>(ctx->batch.iter_dev);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_integrate_optdepth_transmission_iso(helios_ctx* ctx, const double* trans_wg,
                                               double* trans_band, const double* delta_tau_wg,
                                               double* delta_tau_band, const double* gauss_weight,
                                               int nbin, int nlayer, int ny) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(trans_wg && trans_band && delta_tau_wg && delta_tau_band && gauss_weight && nbin > 0 && nlayer > 0 && ny > 0);
    const long long n = (long long)nbin * nlayer;
    k_optdepth_trans<false><<<ceil_div(n, 128), 128, 0, ctx->stream>>>(
        trans_wg, nullptr, trans_band, delta_tau_wg, nullptr, delta_tau_band, gauss_weight, nullptr, nullptr,
        nullptr, nbin, nlayer, ny);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_integrate_optdepth_transmission_noniso(
    helios_ctx* ctx, const double* trans_wg_upper, const double* trans_wg_lower, double* trans_band,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower, double* delta_tau_band,
    const double* gauss_weight, double* delta_tau_all_clouds, const double* delta_tau_all_clouds_upper,
    const double* delta_tau_all_clouds_lower, int nbin, int nlayer, int ny) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(trans_wg_upper && trans_wg_lower && trans_band && delta_tau_wg_upper && delta_tau_wg_lower &&
         delta_tau_band && gauss_weight && delta_tau_all_clouds && delta_tau_all_clouds_upper &&
         delta_tau_all_clouds_lower && nbin > 0 && nlayer > 0 && ny > 0);
    const long long n = (long long)nbin * nlayer;
    k_optdepth_trans<true><<<ceil_div(n, 128), 128, 0, ctx->stream>>>(
        trans_wg_upper, trans_wg_lower, trans_band, delta_tau_wg_upper, delta_tau_wg_lower, delta_tau_band,
        gauss_weight, delta_tau_all_clouds, delta_tau_all_clouds_upper, delta_tau_all_clouds_lower, nbin,
        nlayer, ny);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_contr_func_iso(helios_ctx* ctx, const double* trans_wg, double* trans_weight_band,
                               double* contr_func_band, const double* gauss_weight,
                               const double* planckband_lay, double epsi, int nbin, int nlayer, int ny) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(trans_wg && trans_weight_band && contr_func_band && gauss_weight && planckband_lay && nbin > 0 &&
         nlayer > 0 && ny > 0);
    k_contr_func<false><<<ceil_div(nbin, 64), 64, 0, ctx->stream>>>(
        trans_wg, nullptr, trans_weight_band, contr_func_band, gauss_weight, planckband_lay, epsi, nbin,
        nlayer, ny);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_contr_func_noniso(helios_ctx* ctx, const double* trans_wg_upper,
                                  const double* trans_wg_lower, double* trans_weight_band,
                                  double* contr_func_band, const double* gauss_weight,
                                  const double* planckband_lay, double epsi, int nbin, int nlayer, int ny) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(trans_wg_upper && trans_wg_lower && trans_weight_band && contr_func_band && gauss_weight &&
         planckband_lay && nbin > 0 && nlayer > 0 && ny > 0);
    k_contr_func<true><<<ceil_div(nbin, 64), 64, 0, ctx->stream>>>(
        trans_wg_upper, trans_wg_lower, trans_weight_band, contr_func_band, gauss_weight, planckband_lay,
        epsi, nbin, nlayer, ny);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_calc_mean_opacities(helios_ctx* ctx, double* planck_opac_T_pl, double* ross_opac_T_pl,
                               double* planck_opac_T_star, double* ross_opac_T_star,
                               const double* opac_wg_lay, const double* abs_cross_all_clouds_lay,
                               const double* meanmolmass_lay, const double* planckband_lay,
                               const double* opac_interwave, const double* opac_deltawave,
                               const double* T_lay, const double* gauss_weight, const double* gauss_y,
                               double* opac_band_lay, int nlayer, int nbin, int ny, double T_star) {
    HCTX(ctx);
    HNOBATCH(ctx);
    HARG(planck_opac_T_pl && ross_opac_T_pl && planck_opac_T_star && ross_opac_T_star && opac_wg_lay &&
         abs_cross_all_clouds_lay && meanmolmass_lay && planckband_lay && opac_interwave && opac_deltawave &&
         T_lay && gauss_weight && gauss_y && opac_band_lay && nlayer > 0 && nbin > 0 && ny > 0);
    k_mean_opacities<<<nlayer, IF_THREADS, 0, ctx->stream>>>(
        planck_opac_T_pl, ross_opac_T_pl, planck_opac_T_star, ross_opac_T_star, opac_wg_lay,
        abs_cross_all_clouds_lay, meanmolmass_lay, planckband_lay, opac_interwave, opac_deltawave, T_lay,
        gauss_weight, gauss_y, opac_band_lay, nlayer, nbin, ny, T_star);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_integrate_beamflux(helios_ctx* ctx, double* F_dir_tot, const double* F_dir_band,
                              const double* deltalambda, const double* gauss_weight, int nbin,
                              int numinterfaces) {
    HCTX(ctx);
    HNOBATCH(ctx);
    (void)gauss_weight;
    HARG(F_dir_tot && F_dir_band && deltalambda && nbin > 0 && numinterfaces > 0);
    k_integrate_beamflux<<<numinterfaces, IF_THREADS, 0, ctx->stream>>>(F_dir_tot, F_dir_band, deltalambda,
                                                                        nbin);
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

}  // extern "C"
