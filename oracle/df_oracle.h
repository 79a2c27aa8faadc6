/* oracle/df_oracle.h -- TEST INFRASTRUCTURE ONLY.  See df_oracle.c. */
#ifndef DF_ORACLE_H
#define DF_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

void orc_default_grid(int Ny, double d_i, double* y_vert, double* yc, double* dy);
void orc_linear_interpolate(int n, const double* y_data, const double* f_data, int m, const double* y_new,
                            double* f_new);
void orc_half_widths(int Ny, int Nz, const double* yc, const double* dy, const double* dz, double d_i,
                     double Iz_inn, double Iz_out, int* N_y, int* N_z, int* Ny_max, int* Nz_max);
void orc_coeffs(int N, double* b);
void orc_sweep_y(int Ny, int Nz, int Ny_max, int Nz_max, const int* N_y, const double* r_ys, double* r_zs);
void orc_sweep_z(int Ny, int Nz, int Nz_max, const int* N_z, const double* r_zs, double* filt);
void orc_correlate(long n_cells, const double* filt_old, double* filt, double dt, double Lt);
void orc_rst_scaling(int Ny, int Nz, const double* R11, const double* R21, const double* R22, const double* R33,
                     const double* ufilt, const double* vfilt, const double* wfilt,
                     double* ufluc, double* vfluc, double* wfluc,
                     double* ufilt_old, double* vfilt_old, double* wfilt_old);
void orc_sra(int Ny, int Nz, const double* Ms, const double* Us, const double* Ts, const double* rhos,
             const double* ufluc, double* T_fluc, double* rho_fluc);
void orc_step(int Ny, int Nz, const int* const* N_y, const int* const* N_z, const int* Ny_max, const int* Nz_max,
              const double* Lt, const double* rows, double dt, int first_step,
              const double* const* r_ys, double* const* r_zs,
              double* const* filt_old, double* const* filt, double* const* fluc,
              double* T_fluc, double* rho_fluc);

#ifdef __cplusplus
}
#endif
#endif
