/* tests/comm_c_test.c -- BASELINE config 4 through the C ABI alone (include/dfb200.h), the way a C / C++ / Fortran CFD code
 * would drive it: one process per GPU (fork; the NCCL id travels through a pipe, as MPI_Bcast would carry it), each rank filters
 * its own spanwise slab, rank 0 gathers the finished plane (u', v', w' on the wire, T', rho' rebuilt) and checks it, bit for bit,
 * against the same plane filtered whole on its own GPU.  The gather of step t is overlapped with step t+1.
 *   usage: comm_c_test [world=2]      prints "OK world Ny Nz wire_bytes" or exits non-zero */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include "dfb200.h"

enum { NY = 96, NZ = 1408 };

static void make_config(dfb_config* c, double* geo, double* rows, double* scales) {
    dfb_config_init(c);
    const double I0 = 1.0, dy = 0.67 * I0 / 5.5, dz = I0 / 6.5;          /* N_y = 10, N_z = 12 everywhere (df.cpp:146-149,187-189) */
    double *yc = geo, *dyv = geo + NY, *dzv = geo + 2 * NY;
    for (int j = 0; j < NY; ++j) { yc[j] = (j + 0.5) * dy; dyv[j] = dy; dzv[j] = dz; }
    const double rv[8] = {4.0, -1.0, 2.0, 3.0, 100.0, 300.0, 1.0, 0.3};  /* R11,R21,R22,R33,Us,Ts,rhos,Ms */
    for (int t = 0; t < 8; ++t) for (int j = 0; j < NY; ++j) rows[t * NY + j] = rv[t] * (1.0 + 0.01 * j);
    for (int f = 0; f < 3; ++f) { scales[3 * f] = I0; scales[3 * f + 1] = I0; scales[3 * f + 2] = 1e-6 * (f + 1); }
    c->honor_flow_config = 1; c->d_i = 1.0; c->U_e = 100.0;
    c->Ny = NY; c->Nz = NZ; c->geom_per_row = 1; c->yc = yc; c->dy = dyv; c->dz = dzv; c->rows = rows; c->scales = scales;
    c->seed = 7;
}

#define CHECK(call) do { if ((call) != DFB_OK) { fprintf(stderr, "rank %d: %s: %s\n", rank, #call, dfb_last_error()); return 10; } } while (0)

static int run_rank(int rank, int world, const unsigned char* id) {
    static double geo[3 * NY], rows[8 * NY], scales[9];
    dfb_config c;
    make_config(&c, geo, rows, scales);
    c.device = rank;
    /* slabs cut on multiples of 16 columns (what parallel.slab_bounds produces) */
    int cut[17];
    for (int r = 0; r <= world; ++r) cut[r] = (int)((long)NZ * r / world / 16) * 16;
    cut[world] = NZ;
    c.k_begin = cut[rank]; c.k_end = cut[rank + 1];
    dfb_handle h = NULL;
    CHECK(dfb_create(&c, &h));
    CHECK(dfb_comm_init(h, id, rank, world));
    int r2 = -1, w2 = -1, bounds[34];
    CHECK(dfb_comm_info(h, &r2, &w2, bounds));
    if (r2 != rank || w2 != world || bounds[2 * rank] != c.k_begin) return 11;

    const double dts[3] = {1e-7, 3e-7, 2e-7};
    const size_t n = (size_t)NY * NZ;
    double* got = malloc(3 * 5 * n * sizeof(double));
    /* step 0, gather 0 begun; then step t+1 is enqueued BEFORE gather t is awaited (overlap) */
    CHECK(dfb_filter(h, dts[0]));
    CHECK(dfb_gather_begin(h, 0));
    for (int s = 0; s < 3; ++s) {
        if (s + 1 < 3) CHECK(dfb_filter(h, dts[s + 1]));
        CHECK(dfb_gather_end(h));
        if (rank == 0)
            for (int w = 0; w < 5; ++w) CHECK(dfb_gathered_to_host(h, w, got + ((size_t)s * 5 + w) * n));
        if (s + 1 < 3) CHECK(dfb_gather_begin(h, 0));
    }
    int64_t wire = 0;
    CHECK(dfb_gather_wire_bytes(h, &wire));
    int rc = 0;
    if (rank == 0) {
        /* the same plane filtered whole on this GPU */
        c.k_begin = c.k_end = 0;
        dfb_handle whole = NULL;
        CHECK(dfb_create(&c, &whole));
        double* ref = malloc(n * sizeof(double));
        for (int s = 0; s < 3 && rc == 0; ++s) {
            CHECK(dfb_filter(whole, dts[s]));
            for (int w = 0; w < 5; ++w) {
                CHECK(dfb_get_field(whole, w, ref, 0));
                if (memcmp(ref, got + ((size_t)s * 5 + w) * n, n * sizeof(double)) != 0) { fprintf(stderr, "step %d field %d differs\n", s, w); rc = 12; break; }
            }
        }
        free(ref);
        dfb_destroy(whole);
        if (rc == 0) printf("OK %d %d %d %lld\n", world, NY, NZ, (long long)wire);
    }
    free(got);
    CHECK(dfb_comm_destroy(h));
    dfb_destroy(h);
    return rc;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    if (world < 1 || world > 16) return 2;
    int pipes[16][2];
    pid_t pid[16];
    for (int r = 1; r < world; ++r) {
        if (pipe(pipes[r]) != 0) return 3;
        pid[r] = fork();                                   /* before any CUDA call */
        if (pid[r] == 0) {
            unsigned char id[DFB_COMM_ID_BYTES];
            close(pipes[r][1]);
            if (read(pipes[r][0], id, sizeof(id)) != (ssize_t)sizeof(id)) _exit(4);
            _exit(run_rank(r, world, id));
        }
        close(pipes[r][0]);
    }
    unsigned char id[DFB_COMM_ID_BYTES];
    if (dfb_comm_unique_id(id) != DFB_OK) { fprintf(stderr, "unique id: %s\n", dfb_last_error()); return 5; }
    for (int r = 1; r < world; ++r) if (write(pipes[r][1], id, sizeof(id)) != (ssize_t)sizeof(id)) return 6;
    int rc = run_rank(0, world, id);
    for (int r = 1; r < world; ++r) {
        int st = 0;
        waitpid(pid[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 20 + r;
    }
    return rc;
}
