/*
 * helios_b200.h -- C-ABI of libhelios_b200.so, the B200-native (sm_100a) backend for the
 * radiative-transfer hot path of HELIOS.
 *
 * The reference has no FFI: its "interface" for this path is the set of PyCUDA launches
 *   SourceModule.get_function(name)(args..., block=, grid=)
 * in source/computation.py plus the gpuarray/mem_alloc buffer calls in source/quantities.py.
 * Every entry point below replaces exactly one of those launch sites (cited as K: = source/kernels.cu,
 * C: = source/computation.py, Q: = source/quantities.py, H: = source/host_functions.py) and keeps
 * the reference kernel's argument order and meaning, with the context handle prepended.
 *
 * Conventions
 *   - every function returns 0 (HELIOS_OK) or a non-zero status; helios_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - all `double*` / `int*` array arguments are DEVICE pointers obtained from helios_buf_alloc
 *     (the library owns device memory); scalars are passed by value.
 *   - launches are asynchronous on the context's stream; helios_buf_d2h and helios_ctx_sync
 *     synchronise.  (The reference synchronises the whole device after every launch, C:60 etc.)
 *   - array layouts are the reference's: "wg" arrays are [i][x][y] with y fastest
 *     (idx = y + ny*x + ny*nbin*i, K:1076); band arrays are [i][x] (K:2456); the Planck arrays are
 *     [x][i] with i fastest (K:940, K:1004).
 *   - fp64 only (`precision = double`, K:24-32).
 */
#ifndef HELIOS_B200_H
#define HELIOS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HELIOS_OK 0
#define HELIOS_ERR_CUDA 1
#define HELIOS_ERR_ARG 2
#define HELIOS_ERR_NOMEM 3
#define HELIOS_ERR_STATE 4

#define HELIOS_ABI_VERSION 1

typedef struct helios_ctx helios_ctx;
typedef struct helios_event helios_event;

/* ------------------------------------------------------------------ runtime ------------------ */

int helios_abi_version(void);
const char* helios_last_error(void);
int helios_device_count(int* count);

/* replaces `import pycuda.autoinit` (C:24): binds a context to one device and creates its stream */
int helios_ctx_create(int device, helios_ctx** out);
int helios_ctx_destroy(helios_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int helios_ctx_set_stream(helios_ctx* ctx, void* cuda_stream);
int helios_ctx_get_stream(helios_ctx* ctx, void** cuda_stream);
int helios_ctx_sync(helios_ctx* ctx);                       /* cuda.Context.synchronize(), C:60 */
int helios_ctx_device_info(helios_ctx* ctx, int* num_sms, size_t* l2_bytes, size_t* total_mem);
/* number of kernel launches issued through this context since creation (bench bookkeeping) */
int helios_ctx_launch_count(helios_ctx* ctx, unsigned long long* count);
/* bytes currently allocated through helios_buf_alloc */
int helios_ctx_bytes_allocated(helios_ctx* ctx, size_t* nbytes);

/* Benchmark hygiene, no reference counterpart: evict the L2 cache between timed steps (stream-ordered).
 * mode 0: overwrite a context-owned buffer of 2x the L2 size (L2 is left full of dirty lines, whose write-back
 *         overlaps whatever runs next);
 * mode 1: the same followed by a read sweep over that buffer, so that L2 ends up full of CLEAN foreign lines:
 *         everything is cold for the next kernel and no write-back traffic is charged to it. */
int helios_l2_flush(helios_ctx* ctx, int mode);

/* flux-sweep algorithm: 0 = automatic (layer-parallel kernel whenever the shape fits, default),
 * 1 = one thread per column (fband.cu), 2 = layer-parallel only (error if the shape does not fit) */
int helios_ctx_set_fband_mode(helios_ctx* ctx, int mode);

/* Batched atmospheres (BASELINE.json configs[4]; no reference counterpart -- the reference runs one atmosphere
 * per process).  After helios_ctx_set_batch(ctx, nbatch > 1, ...) the per-iteration entry points
 *   temp_inter, planck_interpol_layer/interface, opac_interpol, meanmolmass_interpol,
 *   calc_total_g_0_of_gas_and_clouds, calc_trans_iso/noniso, calc_delta_z, fdir_iso/noniso (no geometric zenith
 *   correction), fband_iso/noniso, integrate_flux_double, rad_temp_iter, abort_sum
 * process nbatch atmospheres of identical shape (nlayer, nbin, ny) in ONE launch each.  Conventions:
 *   - every per-atmosphere array argument holds nbatch consecutive single-atmosphere arrays, each of the size the
 *     reference allocates for it (Q:400-461, 613-665; e.g. all [i][x][y] arrays are ninterface*ny*nbin, Q:407);
 *   - shared arguments are passed once: the (P,T) grids, the opacity / Rayleigh / mean-molecular-mass tables
 *     (several tables may be stacked: atmosphere b reads table table_index[b], the strides give the distance
 *     between consecutive tables in doubles), the Planck table, wavelength grids, Gauss weights, surf_albedo;
 *   - g[b] replaces the scalar gravity argument; planck_star[b][x] replaces row `dim` of the Planck table
 *     (the stellar slot of planckband_lay, K:940) so that atmospheres with different T_star share one table;
 *   - scalar arguments (dimensions, flags, mu_star, epsi, ...) apply to all atmospheres;
 *   - abort_sum writes one sum per atmosphere and latches done[b] once all nlayer+1 flags of b are set;
 *     rad_temp_iter and fband_* skip atmospheres whose done flag is set, so a converged atmosphere keeps exactly
 *     the state a single-atmosphere run ends with while the others iterate on.
 * All other entry points return HELIOS_ERR_STATE in batch mode.  nbatch = 1 leaves batch mode. */
int helios_ctx_set_batch(helios_ctx* ctx, int nbatch, int nlayer, int nbin, int ny, const int* table_index,
                         size_t ktable_stride, size_t crosstable_stride, size_t meanmass_stride, const double* g,
                         const double* planck_star);
/* nbatch = 0 leaves batch mode; nbatch = 1 is a single atmosphere WITH the on-device bookkeeping (latch, counter).
 * Copies the done flags / the iteration counts at which they latched (both [nbatch]) / the device iteration
 * counter to the host (any pointer may be NULL; the call synchronises if one is given); reset != 0 clears all
 * three on the device first. */
int helios_ctx_batch_state(helios_ctx* ctx, int* done_host, int* converged_at_host, int* iter_host, int reset);
/* enable != 0: rad_temp_iter takes `itervalue` from the device iteration counter instead of its argument, and
 * abort_sum records the count at which an atmosphere latches; helios_batch_iter_advance (one tiny launch)
 * increments the counter.  This is what makes a whole block of RT iterations replayable as one CUDA graph. */
int helios_ctx_batch_device_iteration(helios_ctx* ctx, int enable);
int helios_batch_iter_advance(helios_ctx* ctx);

/* CUDA-graph capture of a launch sequence (no reference counterpart: the reference launches kernel by kernel with a
 * full device sync after each, C:60).  Everything issued on the context between begin and end is recorded instead
 * of executed; helios_graph_launch replays the recording with one driver call.  Entry points that synchronise
 * (helios_buf_d2h, helios_buf_h2d, helios_ctx_sync, helios_corr_inc_energy with a host result, a first-time
 * scratch allocation) must not be called while capturing: run the sequence once eagerly first. */
typedef struct helios_graph helios_graph;
int helios_graph_begin(helios_ctx* ctx);
int helios_graph_end(helios_ctx* ctx, helios_graph** out);
int helios_graph_launch(helios_ctx* ctx, helios_graph* graph);
int helios_graph_destroy(helios_graph* graph);

/* buffers: replace gpuarray.to_gpu / cuda.mem_alloc / .get() (Q:463-665) */
int helios_buf_alloc(helios_ctx* ctx, size_t nbytes, void** dptr);
int helios_buf_free(helios_ctx* ctx, void* dptr);
int helios_buf_h2d(helios_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);
int helios_buf_d2h(helios_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes);
int helios_buf_d2d(helios_ctx* ctx, void* dst_dev, const void* src_dev, size_t nbytes);
int helios_buf_zero(helios_ctx* ctx, void* dptr, size_t nbytes);
/* asynchronous copies for pinned host memory (no implicit sync) */
int helios_buf_h2d_async(helios_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);
int helios_buf_d2h_async(helios_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes);
int helios_host_alloc(size_t nbytes, void** hptr);          /* pinned host memory */
int helios_host_free(void* hptr);

/* events: replace cuda.Event() / record / time_till (C:838-841, 962-965) */
int helios_event_create(helios_ctx* ctx, helios_event** ev);
int helios_event_destroy(helios_event* ev);
int helios_event_record(helios_ctx* ctx, helios_event* ev);
int helios_event_synchronize(helios_event* ev);
int helios_event_elapsed_ms(helios_event* start, helios_event* stop, float* ms);

/* ------------------------------------------------------------------ set-up kernels ----------- */

/* K:362 plancktable, launched 10x at C:43-58.  One call fills rows 0..dim-1 (T = 1 + t*step) and
 * row dim (T = Tstar) of planck_grid[(dim+1)*nwave]. */
int helios_plancktable(helios_ctx* ctx, double* planck_grid, const double* lambda_edge,
                       const double* deltalambda, int nwave, double Tstar, int dim, int step);

/* K:420 corr_inc_energy, C:67-78.  corr_factor_host (may be NULL) receives theo_flux/num_flux,
 * which the reference prints from device code (K:455-456). */
int helios_corr_inc_energy(helios_ctx* ctx, double* planck_grid, double* starflux,
                           const double* deltalambda, int realstar, int nwave, double Tstar, int dim,
                           double* corr_factor_host);

/* ------------------------------------------------------------------ per-iteration kernels ---- */

/* K:496 temp_inter, C:107-115 */
int helios_temp_inter(helios_ctx* ctx, const double* tlay, double* tint, int numinterfaces);

/* K:923 planck_interpol_layer, C:298-311 */
int helios_planck_interpol_layer(helios_ctx* ctx, const double* temp, double* planckband_lay,
                                 const double* planck_grid, const double* starflux, int realstar,
                                 int numlayers, int nwave, int dim, int step);
/* K:981 planck_interpol_interface, C:316-327 */
int helios_planck_interpol_interface(helios_ctx* ctx, const double* temp, double* planckband_int,
                                     const double* planck_grid, int numinterfaces, int nwave,
                                     int dim, int step);

/* temp_inter + planck_interpol_layer + planck_interpol_interface (K:496, K:923, K:981; C:856-857) in ONE launch,
 * bitwise the same results as the three separate calls: what every RT iteration does before its flux solve.
 * planckband_int may be NULL (isothermal layers).  B200-side addition used by the device-resident loop. */
int helios_iteration_prepare(helios_ctx* ctx, const double* tlay, double* tint, double* planckband_lay,
                             double* planckband_int, const double* planck_grid, const double* starflux,
                             int realstar, int numlayers, int nwave, int dim, int step);

/* K:524 opac_interpol, C:122-159 */
int helios_opac_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                         const double* press, const double* opacpress, const double* ktable,
                         double* opac, const double* crosstable, double* scat_cross, int npress,
                         int ntemp, int ny, int nbin, int nlay_or_nint);

/* K:649 meanmolmass_interpol, C:166-195 */
int helios_meanmolmass_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                                double* meanmolmass, const double* opac_meanmass,
                                const double* press, const double* opacpress, int npress, int ntemp,
                                int ninterface);

/* K:703 kappa_interpol, K:761 cp_interpol (C:204-248); K:815 entropy_interpol (C:257-269);
 * K:869 phase_number_interpol (C:278-290) */
int helios_kappa_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                          const double* press, const double* entr_press, double* kappa,
                          const double* entr_kappa, int entr_npress, int entr_ntemp, int nlay_or_nint);
int helios_cp_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                       const double* press, const double* entr_press, double* cp_lay,
                       const double* entr_cp, int entr_npress, int entr_ntemp, int nlayer);
int helios_entropy_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                            const double* press, const double* entr_press, double* entropy,
                            const double* entr_entropy, int entr_npress, int entr_ntemp, int nlayer);
int helios_phase_number_interpol(helios_ctx* ctx, const double* temp, const double* entr_temp,
                                 const double* press, const double* entr_press, double* state,
                                 const double* entr_state, int entr_npress, int entr_ntemp, int nlayer);

/* K:3209 opac_species_interpol, C:1301-1334 */
int helios_opac_species_interpol(helios_ctx* ctx, const double* temp, const double* opactemp,
                                 const double* press, const double* opacpress,
                                 const double* opac_opacity_pretab, double* opac_spec_wg, int npress,
                                 int ntemp, int ny, int nbin, int nlay_or_nint);

/* K:3263 add_to_mixed_opac, C:1350-1386.  Random overlap requires ny == 20 as in the reference
 * (K:3314, K:3393); any other ny with ro_method=1 and s>0 is rejected with HELIOS_ERR_ARG unless
 * ny == 1 (which the reference routes to correlated-k, K:3302). */
int helios_add_to_mixed_opac(helios_ctx* ctx, const double* vmr, const double* opac_spec,
                             double* opac_wg, const double* meanmolmass, const double* gauss_weight,
                             const double* gauss_y, double mass_spec, int s, int ro_method, int ny,
                             int nbin, int nlay_or_nint);

/* K:3404 calc_h2o_scat, C:1394-1421 */
int helios_calc_h2o_scat(helios_ctx* ctx, const double* temp, const double* press,
                         const double* wave, double* scat_cross, const double* vmr, double mass_h2o,
                         int nbin, int nlay_or_nint);
/* K:3444 add_to_mixed_scat, C:1428-1450 */
int helios_add_to_mixed_scat(helios_ctx* ctx, const double* vmr, const double* scat_cross_spec,
                             double* scat_cross, int nbin, int nlay_or_nint);

/* K:472 calc_total_g_0_of_gas_and_clouds, C:334-360 */
int helios_calc_total_g_0_of_gas_and_clouds(helios_ctx* ctx, const double* scat_cross,
                                            const double* g_0_all_clouds,
                                            const double* scat_cross_all_clouds, double* g_0_tot,
                                            double g_0, int nbin, int nlay_or_nint);

/* K:1015 calc_trans_iso, C:371-406 */
int helios_calc_trans_iso(helios_ctx* ctx, double* trans_wg, double* delta_tau_wg, double* M_term,
                          double* N_term, double* P_term, double* G_plus, double* G_minus,
                          const double* delta_colmass, const double* opac_wg_lay,
                          const double* meanmolmass_lay, const double* scat_cross_lay,
                          const double* abs_cross_all_clouds_lay,
                          const double* scat_cross_all_clouds_lay, double* delta_tau_all_clouds,
                          double* w_0, const double* g_0_tot_lay, int* scat_trigger, double g_0,
                          double epsi, double epsi2, double mu_star, double w_0_limit,
                          double w_0_scat_limit, int scat, int nbin, int ny, int nlayer, int clouds,
                          int scat_corr, int debug, double i2s_transition);

/* K:1107 calc_trans_noniso, C:409-460 */
int helios_calc_trans_noniso(
    helios_ctx* ctx, double* trans_wg_upper, double* trans_wg_lower, double* delta_tau_wg_upper,
    double* delta_tau_wg_lower, double* M_upper, double* M_lower, double* N_upper, double* N_lower,
    double* P_upper, double* P_lower, double* G_plus_upper, double* G_plus_lower,
    double* G_minus_upper, double* G_minus_lower, const double* delta_col_upper,
    const double* delta_col_lower, const double* opac_wg_lay, const double* opac_wg_int,
    const double* meanmolmass_lay, const double* meanmolmass_int, const double* scat_cross_lay,
    const double* scat_cross_int, const double* abs_cross_all_clouds_lay,
    const double* abs_cross_all_clouds_int, const double* scat_cross_all_clouds_lay,
    const double* scat_cross_all_clouds_int, double* delta_tau_all_clouds_upper,
    double* delta_tau_all_clouds_lower, double* w_0_upper, double* w_0_lower,
    const double* g_0_tot_lay, const double* g_0_tot_int, int* scat_trigger, double g_0, double epsi,
    double epsi2, double mu_star, double w_0_limit, double w_0_scat_limit, int scat, int nbin, int ny,
    int nlayer, int clouds, int scat_corr, int debug, double i2s_transition);

/* K:1247 calc_delta_z, C:467-477 */
int helios_calc_delta_z(helios_ctx* ctx, const double* tlay, const double* pint, const double* play,
                        const double* meanmolmass_lay, double* delta_z_lay, double g, int nlayer);

/* K:1265 fdir_iso, C:486-502 */
int helios_fdir_iso(helios_ctx* ctx, double* F_dir_wg, const double* planckband_lay,
                    const double* delta_tau_wg, const double* z_lay, double mu_star, double R_planet,
                    double R_star, double a, int dir_beam, int geom_zenith_corr, int ninterface,
                    int nbin, int ny);
/* K:1313 fdir_noniso, C:506-524 */
int helios_fdir_noniso(helios_ctx* ctx, double* F_dir_wg, double* Fc_dir_wg,
                       const double* planckband_lay, const double* delta_tau_wg_upper,
                       const double* delta_tau_wg_lower, const double* z_lay, double mu_star,
                       double R_planet, double R_star, double a, int dir_beam, int geom_zenith_corr,
                       int ninterface, int nbin, int ny);

/* K:1366 fband_iso.  The reference launches it (3*scat+1) times per iteration, or 1000*scat+1 times in
 * post-processing (C:531-571), each launch a full device sync.  `npass` fuses those launches: one
 * call performs npass consecutive down+up sweeps with identical results (columns are independent,
 * so pass p+1 of a column only needs that column's own pass-p fluxes). */
int helios_fband_iso(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* F_dir_wg,
                     const double* planckband_lay, const double* w_0, const double* M_term,
                     const double* N_term, const double* P_term, const double* G_plus,
                     const double* G_minus, const double* surf_albedo, const double* g_0_tot_lay,
                     double g_0, int singlewalk, double Rstar, double a, int numinterfaces, int nbin,
                     double f_factor, double mu_star, int ny, double epsi, int dir_beam, int clouds,
                     int scat_corr, int debug, double i2s_transition, int npass);

/* K:1521 fband_noniso, C:575-621; npass as above */
int helios_fband_noniso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg, double* Fc_up_wg,
    const double* F_dir_wg, const double* Fc_dir_wg, const double* planckband_lay,
    const double* planckband_int, const double* w_0_upper, const double* w_0_lower,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower,
    const double* M_upper, const double* M_lower, const double* N_upper, const double* N_lower,
    const double* P_upper, const double* P_lower, const double* G_plus_upper,
    const double* G_plus_lower, const double* G_minus_upper, const double* G_minus_lower,
    const double* surf_albedo, const double* g_0_tot_lay, const double* g_0_tot_int, double g_0,
    int singlewalk, double Rstar, double a, int numinterfaces, int nbin, double f_factor,
    double mu_star, int ny, double epsi, double delta_tau_limit, int dir_beam, int clouds,
    int scat_corr, int debug, double i2s_transition, int npass);

/* Sweep plan for non-isothermal layers (B200-side addition, no reference counterpart).  Between two opacity refreshes
 * (10 RT iterations, C:860) only the Planck terms of fband_noniso's inputs change.  helios_fband_noniso_plan_build
 * evaluates everything else the sweep derives from the coefficient arrays -- P/M, N/M, the source and gradient
 * factors, the direct-beam sources -- once, into `plan` (helios_fband_noniso_plan_size doubles, laid out in the
 * order the sweep streams it; 0 = more than 128 layers, unsupported; in batch mode the size covers the batch).
 * Call it after calc_trans_noniso and fdir_noniso.  helios_fband_noniso_planned then performs the same npass fused sweeps
 * as helios_fband_noniso from the plan plus the current Planck arrays and the previous fluxes; results agree with
 * helios_fband_noniso to rounding (<= 1e-13 relative on the fluxes; the plan folds the Planck-independent factors
 * into affine coefficients).  The plan is invalid as soon as any coefficient array, F_dir or surf_albedo changes. */
int helios_fband_noniso_plan_size(helios_ctx* ctx, int numinterfaces, int nbin, int ny, size_t* ndoubles);
int helios_fband_noniso_plan_build(
    helios_ctx* ctx, double* plan, const double* F_dir_wg, const double* Fc_dir_wg, const double* w_0_upper,
    const double* w_0_lower, const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower, const double* M_upper,
    const double* M_lower, const double* N_upper, const double* N_lower, const double* P_upper, const double* P_lower,
    const double* G_plus_upper, const double* G_plus_lower, const double* G_minus_upper, const double* G_minus_lower,
    const double* surf_albedo, const double* g_0_tot_lay, const double* g_0_tot_int, double g_0, int numinterfaces,
    int nbin, double mu_star, int ny, double epsi, double delta_tau_limit, int clouds, int scat_corr,
    double i2s_transition);
int helios_fband_noniso_planned(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg,
                                double* Fc_up_wg, const double* plan, const double* planckband_lay,
                                const double* planckband_int, const double* surf_albedo, double Rstar, double a,
                                int numinterfaces, int nbin, double f_factor, int ny, int dir_beam, int npass);

/* The same for isothermal layers (one step per layer: a, b, the Planck factor and the two beam sources per cell).
 * Call helios_fband_iso_plan_build after calc_trans_iso and fdir_iso; when the beam is known to be zero (fdir_iso
 * ran with dir_beam == 0 and F_dir_wg was not written since) the sweep does not read the beam sources. */
int helios_fband_iso_plan_size(helios_ctx* ctx, int numinterfaces, int nbin, int ny, size_t* ndoubles);
int helios_fband_iso_plan_build(helios_ctx* ctx, double* plan, const double* F_dir_wg, const double* w_0,
                                const double* M_term, const double* N_term, const double* P_term,
                                const double* G_plus, const double* G_minus, const double* surf_albedo,
                                const double* g_0_tot_lay, double g_0, int numinterfaces, int nbin, double mu_star,
                                int ny, double epsi, int dir_beam, int clouds, int scat_corr, double i2s_transition);
int helios_fband_iso_planned(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* plan,
                             const double* planckband_lay, const double* surf_albedo, double Rstar, double a,
                             int numinterfaces, int nbin, double f_factor, int ny, int dir_beam, int npass);

/* K:1803 fband_matrix_iso, C:630-668.  alpha/beta/source_term_* are accepted for signature
 * compatibility but not touched: the Thomas coefficients are formed on the fly; c_prime/d_prime
 * (2*ninterface*ny*nbin doubles each) hold the forward elimination. */
int helios_fband_matrix_iso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* F_dir_wg,
    const double* planckband_lay, const double* w_0, const double* M_term, const double* N_term,
    const double* P_term, const double* G_plus, const double* G_minus, const double* g_0_tot_lay,
    double* alpha, double* beta, double* source_term_down, double* source_term_up, double* c_prime,
    double* d_prime, const int* scat_trigger, const double* trans_wg, const double* surf_albedo,
    double g_0, int singlewalk, double Rstar, double a, int numinterfaces, int nbin, double f_factor,
    double mu_star, int ny, double epsi, int dir_beam, int clouds, int scat_corr, int debug,
    double i2s_transition);

/* K:2028 fband_matrix_noniso, C:672-727.  c_prime/d_prime need (4*ninterface-2)*ny*nbin doubles. */
int helios_fband_matrix_noniso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg, double* Fc_up_wg,
    const double* F_dir_wg, const double* Fc_dir_wg, const double* planckband_lay,
    const double* planckband_int, const double* w_0_upper, const double* w_0_lower,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower,
    const double* M_upper, const double* M_lower, const double* N_upper, const double* N_lower,
    const double* P_upper, const double* P_lower, const double* G_plus_upper,
    const double* G_plus_lower, const double* G_minus_upper, const double* G_minus_lower,
    const double* g_0_tot_lay, const double* g_0_tot_int, double* alpha, double* beta,
    double* source_term_down, double* source_term_up, double* c_prime, double* d_prime,
    const int* scat_trigger, const double* trans_wg_upper, const double* trans_wg_lower,
    const double* surf_albedo, double g_0, int singlewalk, double Rstar, double a, int numinterfaces,
    int nbin, double f_factor, double mu_star, int ny, double epsi, double delta_tau_limit,
    int dir_beam, int clouds, int scat_corr, int debug, double i2s_transition);

/* K:2428 integrate_flux_double, C:739-755.  Multi-block, fixed summation order (the reference's
 * single-block CAS-atomic order is nondeterministic). */
int helios_integrate_flux_double(helios_ctx* ctx, const double* deltalambda, double* F_down_tot,
                                 double* F_up_tot, double* F_net, const double* F_down_wg,
                                 const double* F_up_wg, const double* F_dir_wg, double* F_down_band,
                                 double* F_up_band, double* F_dir_band, const double* gauss_weight,
                                 int nbin, int numinterfaces, int ny);

/* K:2606 rad_temp_iter, C:762-795 */
int helios_rad_temp_iter(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                         const double* F_net, double* F_net_diff, double* tlay, const double* play,
                         const double* tint, const double* pint, int* abrt, double* T_store,
                         double* deltat_prefactor, const double* F_add_heat_lay,
                         const double* F_add_heat_sum, double* F_smooth, double* F_smooth_sum,
                         const double* c_p_lay, const double* meanmolmass_lay, int itervalue,
                         double f_factor, int foreplay, double g, int numlayers,
                         double physical_tstep, double local_limit, int adapt_interval, int smooth,
                         int dim, int step, double F_intern, int no_atmo);

/* rad_temp_iter + abort_sum + helios_batch_iter_advance in ONE launch (batch mode with the device iteration
 * counter): the block of each atmosphere sums its flags into sum_dev[b] and latches done[b]; the last block to
 * finish advances the counter.  Same arguments as helios_rad_temp_iter plus sum_dev [nbatch]. */
int helios_rad_temp_iter_latched(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                                 const double* F_net, double* F_net_diff, double* tlay, const double* play,
                                 const double* tint, const double* pint, int* abrt, double* T_store,
                                 double* deltat_prefactor, const double* F_add_heat_lay,
                                 const double* F_add_heat_sum, double* F_smooth, double* F_smooth_sum,
                                 const double* c_p_lay, const double* meanmolmass_lay, int itervalue,
                                 double f_factor, int foreplay, double g, int numlayers,
                                 double physical_tstep, double local_limit, int adapt_interval, int smooth,
                                 int dim, int step, double F_intern, int no_atmo, int* sum_dev);

/* K:2768 conv_temp_iter, C:802-823 */
int helios_conv_temp_iter(helios_ctx* ctx, const double* F_down_tot, const double* F_up_tot,
                          const double* F_net, double* F_net_diff, double* tlay, const double* play,
                          const double* pint, double* T_store, double* deltat_prefactor,
                          const int* marked_red, const double* F_add_heat_lay, double* F_smooth,
                          double* F_smooth_sum, int numlayers, int itervalue, int adapt_interval,
                          int smooth, double F_intern);

/* replaces the per-iteration `dev_abort.get()` + Python loop (C:927-932): sums abrt[0..n) on the
 * device into *sum_dev (device int).  B200-side addition, no reference kernel. */
int helios_abort_sum(helios_ctx* ctx, const int* abrt, int n, int* sum_dev);

/* opac_interpol / opac_species_interpol: 1 = stage the four table rows of a layer's (P, T) box through shared memory with
 * TMA bulk copies (k_pt_gather_tma), 0 = stream them with read-only loads (default: measured faster on B200), -1 = back
 * to the default / HELIOS_PT_GATHER.  Process-wide; returns the previous setting.  Both forms give identical results. */
int helios_set_pt_gather_tma(int on);

/* ------------------------------------------------------------------ device-side host functions */

/* host_functions.py:509-538 convective_adjustment (conv_check / mark_convective_layers / conv_correct to stability, then
 * the damped correction), one launch instead of the reference's host round trips (C:1053-1062).  T_lay (nlayer + 1,
 * index nlayer = surface) and conv_layer are updated in place; conv_unstable and status_dev[4] = {adjustment cycles,
 * convective zones, instability found at entry, 1 if the cycle limit was hit} are outputs.  dampara <= 0: "automatic"
 * (H:441-449).  B200-side addition: the reference has no kernel for this. */
int helios_convective_adjustment(helios_ctx* ctx, double* T_lay, const double* p_lay, const double* p_int,
                                 const double* kappa_lay, const double* kappa_int, const double* c_p_lay,
                                 const double* meanmolmass_lay, const double* F_add_heat_sum, const double* F_smooth_sum,
                                 const double* F_down_tot, const double* F_up_tot, int* conv_layer, int* conv_unstable,
                                 int* status_dev, double F_intern, double T_star, double dampara, int iter_value,
                                 int nlayer);

/* host_functions.py:545-582 mark_convective_layers(stitching = 1) followed by H:251-286 check_for_radiative_eq, after the
 * flux solve of a radiative-convective iteration (C:1093-1115).  status_dev[4] = {radiative layers converged, radiative
 * layers, convective layers, layers with T == 0}. */
int helios_convection_marks(helios_ctx* ctx, const double* T_lay, const double* p_lay, const double* p_int,
                            const double* kappa_lay, const double* kappa_int, const double* F_net,
                            const double* F_down_tot, const double* F_add_heat_sum, const double* F_smooth_sum,
                            int* conv_layer, int* marked_red, int* status_dev, double F_intern,
                            double rad_convergence_limit, int iter_value, int nlayer);

/* host_functions.py:874-910: a species' pre-tabulated VMR [ntemp][npress] interpolated bilinearly in (T, log10 P), clamped
 * at the grid edges, along a profile of n points (the reference: scipy RectBivariateSpline(kx = ky = 1) per layer). */
int helios_vmr_interpol(helios_ctx* ctx, const double* temp, const double* press, const double* ktemp,
                        const double* kpress, const double* vmr_pretab, double* vmr_out, int npress, int ntemp, int n);

/* host_functions.py:927-959 calc_meanmolmass on the device: sum_weighted += vmr * weight, sum_vmr += vmr per species, then
 * meanmolmass = sum_weighted / sum_vmr * AMU */
int helios_meanmolmass_accumulate(helios_ctx* ctx, const double* vmr, double weight, double* sum_weighted, double* sum_vmr,
                                  int n);
int helios_meanmolmass_finish(helios_ctx* ctx, const double* sum_weighted, const double* sum_vmr, double* meanmolmass, int n);

/* ------------------------------------------------------------------ post-processing ---------- */

/* K:2888 / K:2916, C:1180-1212 */
int helios_integrate_optdepth_transmission_iso(helios_ctx* ctx, const double* trans_wg,
                                               double* trans_band, const double* delta_tau_wg,
                                               double* delta_tau_band, const double* gauss_weight,
                                               int nbin, int nlayer, int ny);
int helios_integrate_optdepth_transmission_noniso(
    helios_ctx* ctx, const double* trans_wg_upper, const double* trans_wg_lower, double* trans_band,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower, double* delta_tau_band,
    const double* gauss_weight, double* delta_tau_all_clouds,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower, int nbin,
    int nlayer, int ny);

/* K:2951 / K:2987, C:1220-1250.  trans_weight_band is accumulated into (+=) as in the reference. */
int helios_calc_contr_func_iso(helios_ctx* ctx, const double* trans_wg, double* trans_weight_band,
                               double* contr_func_band, const double* gauss_weight,
                               const double* planckband_lay, double epsi, int nbin, int nlayer,
                               int ny);
int helios_calc_contr_func_noniso(helios_ctx* ctx, const double* trans_wg_upper,
                                  const double* trans_wg_lower, double* trans_weight_band,
                                  double* contr_func_band, const double* gauss_weight,
                                  const double* planckband_lay, double epsi, int nbin, int nlayer,
                                  int ny);

/* K:3024 calc_mean_opacities, C:1257-1279 */
int helios_calc_mean_opacities(helios_ctx* ctx, double* planck_opac_T_pl, double* ross_opac_T_pl,
                               double* planck_opac_T_star, double* ross_opac_T_star,
                               const double* opac_wg_lay, const double* abs_cross_all_clouds_lay,
                               const double* meanmolmass_lay, const double* planckband_lay,
                               const double* opac_interwave, const double* opac_deltawave,
                               const double* T_lay, const double* gauss_weight,
                               const double* gauss_y, double* opac_band_lay, int nlayer, int nbin,
                               int ny, double T_star);

/* K:3119 integrate_beamflux, C:1286-1296 */
int helios_integrate_beamflux(helios_ctx* ctx, double* F_dir_tot, const double* F_dir_band,
                              const double* deltalambda, const double* gauss_weight, int nbin,
                              int numinterfaces);

/* ------------------------------------------------------------------ multi-GPU ---------------- */
/* Wavelength sharding: each rank integrates its own bins; the per-interface partial
 * totals are summed across ranks.  The exchange is a one-shot peer-memory reduction over
 * NVLink/NVSwitch: every rank stores its partial vector into a slot of every peer's mailbox
 * (cudaIpc-mapped), then sums the slots in rank order, so all ranks get bitwise-identical totals.
 * No reference counterpart (the reference is single-GPU). */
#define HELIOS_IPC_HANDLE_BYTES 64
/* allocate this rank's mailbox (two banks of world*slot_doubles doubles + flags for the stand-alone kernel, the same
 * again as 16-byte packet pairs for the fused form) and export its IPC handle */
int helios_comm_create(helios_ctx* ctx, int rank, int world, int slot_doubles,
                       unsigned char* handle_out /* HELIOS_IPC_HANDLE_BYTES */);
/* map the peers' mailboxes; handles = world consecutive handles (own entry ignored) */
int helios_comm_connect(helios_ctx* ctx, const unsigned char* handles);
/* vec[0..n) (device, n <= slot_doubles) <- sum over ranks, fixed rank order */
int helios_comm_allreduce_sum(helios_ctx* ctx, double* vec, int n);
/* the per-iteration exchange of a wavelength-sharded run, fused: F_up_tot and F_down_tot (the partial sums
 * over this rank's bins, as written by helios_integrate_flux_double, K:2474-2495) become the sums over all
 * ranks in ONE peer-memory round trip, and F_net = F_up_tot - F_down_tot is recomputed (K:2509).
 * slot_doubles must be >= 2*numinterfaces. */
int helios_comm_allreduce_flux_totals(helios_ctx* ctx, double* F_up_tot, double* F_down_tot, double* F_net,
                                      int numinterfaces);
/* on != 0: helios_integrate_flux_double performs the flux-total exchange in its own launch (the block that finishes an
 * interface pushes this rank's two totals to every mailbox as self-validating {data, round} packets -- no fence, no flag;
 * the block that finishes the last interface polls its own mailbox and sums the slots in rank order); helios_comm_allreduce_flux_totals must then NOT be
 * called for that step.  Every rank has to make the same sequence of exchanging calls. */
int helios_comm_set_fused(helios_ctx* ctx, int on);
int helios_comm_destroy(helios_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HELIOS_B200_H */
