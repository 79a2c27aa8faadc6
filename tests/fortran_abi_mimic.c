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

    /* ---- create_digital_filter_batch / filter_batch (BASELINE config 5, the us3d_user.f90-style caller: create once, then one call
     *      per timestep): arrays u(n_cells, nplanes) = [nplanes][n_cells] in C order, every scalar by reference ---- */
    {
        const int np = 3;
        dfb_handle b = NULL;
        c.plane_id = 10;
        if (dfb_create_batch_f(&c, &np, &b) != DFB_OK) { fprintf(stderr, "create_batch: %s\n", dfb_last_error()); return 9; }
        double *bu = malloc(8 * n * np), *bv = malloc(8 * n * np), *bw = malloc(8 * n * np), *bT = malloc(8 * n * np), *br = malloc(8 * n * np);
        for (int s = 0; s < 2; ++s)
            if (dfb_filter_to_host_f(&b, &dt, bu, bv, bw, bT, br) != DFB_OK) { fprintf(stderr, "filter_batch: %s\n", dfb_last_error()); return 10; }
        /* plane 2 of the batch == a single-plane handle with plane_id 11, bit for bit */
        dfb_handle one = NULL;
        c.plane_id = 11;
        if (dfb_create_f(&c, &one) != DFB_OK) return 11;
        for (int s = 0; s < 2; ++s)
            if (dfb_filter_to_host_f(&one, &dt, u, v, w, T, rho) != DFB_OK) return 12;
        if (memcmp(u, bu + n, 8 * n) || memcmp(rho, br + n, 8 * n)) { fprintf(stderr, "batch plane differs from the single-plane handle\n"); return 13; }
        const int plane = 2, which = DFB_W_FLUC;
        if (dfb_get_field_plane_f(&b, &plane, &which, u) != DFB_OK || memcmp(u, bw + 2 * n, 8 * n)) return 14;
        /* face_map: the centre of cell (j, k) maps to j*Nz + k */
        double yv[2], zv[2];
        double* tab = malloc(8 * (size_t)(Ny + Nz + 2));
        if (dfb_get_table(one, 12, 0, tab, Ny + 1) != DFB_OK || dfb_get_table(one, 13, 0, tab + Ny + 1, Nz + 1) != DFB_OK) return 15;
        const int jj = 17, kk = 123, nf = 2;
        yv[0] = 0.5 * (tab[jj] + tab[jj + 1]); zv[0] = 0.5 * (tab[Ny + 1 + kk] + tab[Ny + 1 + kk + 1]);
        yv[1] = tab[0]; zv[1] = tab[Ny + 1];
        int cell[2];
        if (dfb_face_map_f(&one, &nf, yv, zv, cell) != DFB_OK || cell[0] != jj * Nz + kk || cell[1] != 0) return 16;
        printf("BATCH_OK %d\n", np);
        free(tab); free(bu); free(bv); free(bw); free(bT); free(br);
        dfb_destroy_f(&one); dfb_destroy_f(&b);
    }
    free(u); free(v); free(w); free(T); free(rho);
    return 0;
}
