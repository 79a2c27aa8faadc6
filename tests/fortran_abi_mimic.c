/* tests/fortran_abi_mimic.c -- makes the calls fortran/digital_filtering.f90 makes, the way a Fortran
 * compiler makes them: a bind(C) struct with blank-padded character(len=256) fields passed by
 * c_loc + len_trim, every scalar by reference through the *_f entry points.  No Fortran compiler
 * exists in the image; this is the stand-in that keeps the binding honest.
 *   usage: fortran_abi_mimic <RST.dat> <line.dat>      prints "OK Ny Nz rms_u rms_T" or exits non-zero */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dfb200.h"

static void blank_pad(char* dst, const char* src, int n) {
    int l = (int)strlen(src);
    memset(dst, ' ', (size_t)n);
    memcpy(dst, src, (size_t)(l < n ? l : n));
}
static int len_trim(const char* s, int n) {
    while (n > 0 && s[n - 1] == ' ') --n;
    return n;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s RST.dat line.dat\n", argv[0]); return 2; }
    char grid_file[256], vel_fluc_file[256], line_file[256];      /* character(len=256), no NUL */
    blank_pad(grid_file, "grid.dat", 256);
    blank_pad(vel_fluc_file, argv[1], 256);
    blank_pad(line_file, argv[2], 256);

    dfb_config c;
    if (dfb_config_init(&c) != DFB_OK) return 3;
    c.d_i = 0.0013; c.rho_e = 0.044; c.U_e = 869.1; c.mu_e = 7.1212e-6;      /* fortran-main.f90:12-15 (mu: df.cpp:10) */
    c.vel_file_offset = 0; c.vel_file_N_values = 330;       /* RST.dat layout (fortran-main.f90:18-19 uses 142/330 with the Stat file) */
    c.honor_flow_config = 1;
    c.grid_file = grid_file;         c.grid_file_len = len_trim(grid_file, 256);
    c.vel_fluc_file = vel_fluc_file; c.vel_fluc_file_len = len_trim(vel_fluc_file, 256);
    c.line_file = line_file;         c.line_file_len = len_trim(line_file, 256);
    c.seed = 12345; c.device = -1;

    dfb_handle h = NULL;
    if (dfb_create_f(&c, &h) != DFB_OK) { fprintf(stderr, "create: %s\n", dfb_last_error()); return 4; }
    int Ny = 0, Nz = 0;
    if (dfb_dims_f(&h, &Ny, &Nz) != DFB_OK) return 5;
    size_t n = (size_t)Ny * (size_t)Nz;
    double *u = malloc(8 * n), *v = malloc(8 * n), *w = malloc(8 * n), *T = malloc(8 * n), *rho = malloc(8 * n);
    double dt = 1e-8;                                                      /* fortran-main.f90:24 */
    for (int s = 0; s < 3; ++s)
        if (dfb_filter_to_host_f(&h, &dt, u, v, w, T, rho) != DFB_OK) { fprintf(stderr, "filter: %s\n", dfb_last_error()); return 6; }
    double su = 0, sT = 0;
    for (size_t i = 0; i < n; ++i) { su += u[i] * u[i]; sT += T[i] * T[i]; }
    if (!(su > 0) || !isfinite(su) || !isfinite(sT)) return 7;
    printf("OK %d %d %.6f %.6f\n", Ny, Nz, sqrt(su / (double)n), sqrt(sT / (double)n));
    if (dfb_destroy_f(&h) != DFB_OK || h != NULL) return 8;
    free(u); free(v); free(w); free(T); free(rho);
    return 0;
}
