// Planned sweeps (fband_plan.cu): internal interface used by the C-ABI glue in fband.cu.
#pragma once
#include "common.cuh"

// return HELIOS_OK when launched, -1 when the shape does not fit (more than 256 / 128 layers), -2 when `plan` was not
// built by the matching *_build call of this context (or was overwritten since)
size_t plan2_iso_size(int nint, int ncol, int nbatch);
size_t plan2_noniso_size(int nint, int ncol, int nbatch);
int plan2_iso_build(helios_ctx* ctx, double* plan, const double* F_dir, const double* w_0, const double* M,
                    const double* N, const double* P, const double* Gp, const double* Gm, const double* albedo,
                    const double* g0tot, double g_0, double mu_star, double epsi, int nint, int nbin, int ny,
                    int dir_beam, int clouds, int scat_corr, double i2s);
int plan2_iso_sweep(helios_ctx* ctx, double* F_down, double* F_up, const double* plan, const double* planck_lay,
                    const double* albedo, double Rstar, double a, int nint, int nbin, double f_factor, int ny,
                    int dir_beam, int npass);
// coef: the 16 coefficient arrays in the order w0_u, w0_l, dtau_u, dtau_l, dtc_u, dtc_l, M_u, M_l, N_u, N_l, P_u, P_l,
// Gp_u, Gp_l, Gm_u, Gm_l
int plan2_noniso_build(helios_ctx* ctx, double* plan, const double* F_dir, const double* Fc_dir, const double* const* coef,
                       const double* albedo, const double* g0_lay, const double* g0_int, double g_0, double mu_star,
                       double epsi, double delta_tau_limit, int nint, int nbin, int ny, int dir_beam, int clouds,
                       int scat_corr, double i2s);
int plan2_noniso_sweep(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up, const double* plan,
                       const double* planck_lay, const double* planck_int, const double* albedo, double Rstar, double a,
                       int nint, int nbin, double f_factor, int ny, int dir_beam, int npass);
