/* oracle/normal_oracle.c -- TEST INFRASTRUCTURE ONLY.  Host restatement of the noise contract in
 * include/dfb_rng_spec.h (spec v2): counter-addressed pcg32 draws (the reference's engine,
 * pcg_random.hpp:1866, advanced with its own jump-ahead algorithm, pcg_random.hpp:640-669) fed
 * through a Box-Muller pair transform made only of correctly rounded IEEE operations.
 * Compiled with -ffp-contract=off; every fused operation is an explicit fma().
 * This replaces (does not replicate) df.cpp:332-349's std::normal_distribution stream, as
 * north_star prescribes; the device kernel must match these doubles bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "dfb_rng_spec.h"

typedef struct { uint64_t state, inc; } orc_pcg32;
void orc_pcg32_seed(orc_pcg32* g, uint64_t seed, uint64_t stream);
void orc_pcg32_advance(orc_pcg32* g, uint64_t delta);
uint32_t orc_pcg32_next(orc_pcg32* g);

static const double LOG1P_C[DFB_LOG1P_NC] = {DFB_LOG1P_C_LIST};
static const double LOGTAB[128] = {DFB_LOGTAB_LIST};        /* INV[i] at 2i, L[i] at 2i+1 */
static const double SIN_C[DFB_SIN_NC] = {DFB_SIN_C_LIST};
static const double COS_C[DFB_COS_NC] = {DFB_COS_C_LIST};

/* four consecutive 32-bit draws -> two N(0,1) doubles */
void orc_normal_pair(const uint32_t o[4], double z[2]) {
    uint64_t U1 = ((((uint64_t)o[1] << 32) | o[0]) >> 11) | 1u;
    uint64_t U2 = (((uint64_t)o[3] << 32) | o[2]) >> 11;

    /* ln(u1), u1 = U1 * 2^-53: table over the top six mantissa bits + log1p polynomial */
    double d = (double)U1;
    uint64_t bits; memcpy(&bits, &d, 8);
    unsigned i6 = (unsigned)(bits >> 46) & 63u, hi = i6 >> 5;
    int e = (int)(bits >> 52) - 1023 + (int)hi - 53;
    uint64_t mb = (bits & 0x000FFFFFFFFFFFFFULL) | ((uint64_t)(0x3FFu - hi) << 52);
    double m; memcpy(&m, &mb, 8);
    double rr = fma(m, LOGTAB[2 * i6], -1.0);
    double P = LOG1P_C[DFB_LOG1P_NC - 1];
    for (int k = DFB_LOG1P_NC - 2; k >= 0; --k) P = fma(P, rr, LOG1P_C[k]);
    double lp = fma(rr * rr, P, rr);
    double lnu = fma((double)e, DFB_LN2, LOGTAB[2 * i6 + 1] + lp);
    double r = sqrt(-2.0 * lnu);

    /* sin/cos(2 pi u2), u2 = U2 * 2^-53 : octant from the top 3 bits, 50-bit fraction */
    unsigned oct = (unsigned)(U2 >> 50);
    uint64_t T = U2 & ((1ULL << 50) - 1u);
    if (oct & 1u) T = (1ULL << 50) - T;
    double t = (double)T * 0x1p-50;
    double x = t * DFB_PIO4;
    double x2 = x * x;
    double S = SIN_C[DFB_SIN_NC - 1];
    for (int k = DFB_SIN_NC - 2; k >= 0; --k) S = fma(S, x2, SIN_C[k]);
    double sx = fma(x * x2, S, x);
    double C = COS_C[DFB_COS_NC - 1];
    for (int k = DFB_COS_NC - 2; k >= 0; --k) C = fma(C, x2, COS_C[k]);
    double cx = fma(x2, C, 1.0);
    if (oct & 1u) { double tmp = sx; sx = cx; cx = tmp; }
    double sn, cs;
    switch (oct >> 1) {
        case 0:  sn =  sx; cs =  cx; break;
        case 1:  sn =  cx; cs = -sx; break;
        case 2:  sn = -sx; cs = -cx; break;
        default: sn = -cx; cs =  sx; break;
    }
    z[0] = r * cs;
    z[1] = r * sn;
}

/* elements e0 .. e0+n-1 (global flat indices) of the logical array (seed, stream, step) whose full
 * length is len (npairs = (len+1)/2). */
void orc_noise_elements(uint64_t seed, uint64_t stream, uint64_t step, uint64_t len,
                        uint64_t e0, uint64_t n, double* out) {
    uint64_t npairs = (len + 1u) / 2u;
    uint64_t q = e0 >> 1, qend = (e0 + n + 1u) >> 1;   /* pairs [q, qend) cover the range */
    orc_pcg32 g;
    orc_pcg32_seed(&g, seed, stream);
    orc_pcg32_advance(&g, 4u * (step * npairs + q));
    for (; q < qend; ++q) {
        uint32_t o[4];
        double z[2];
        for (int i = 0; i < 4; ++i) o[i] = orc_pcg32_next(&g);
        orc_normal_pair(o, z);
        for (int i = 0; i < 2; ++i) {
            uint64_t e = 2u * q + (uint64_t)i;
            if (e >= e0 && e < e0 + n) out[e - e0] = z[i];
        }
    }
}

static uint64_t stream_of(int plane, int field, int array) {
    return (uint64_t)(((int64_t)plane * 3 + field) * 2 + array);
}

/* r_ys of one field for the spanwise slab [k0,k1) of an Nz_global-wide plane, in the reference's
 * layout (df.cpp:197,374): (Ny + 2*Ny_max) rows x (k1-k0) columns, row-major. */
void orc_noise_rys(uint64_t seed, int plane, int field, uint64_t step, int Ny, int Ny_max,
                   int NzG, int k0, int k1, double* out) {
    int rows = Ny + 2 * Ny_max, w = k1 - k0;
    uint64_t len = (uint64_t)rows * (uint64_t)NzG;
    for (int r = 0; r < rows; ++r)
        orc_noise_elements(seed, stream_of(plane, field, 0), step, len,
                           (uint64_t)r * (uint64_t)NzG + (uint64_t)k0, (uint64_t)w, out + (size_t)r * (size_t)w);
}
/* the 2*Nz_max raw-noise halo columns of r_zs (df.cpp:157,398): Ny x (2*Nz_max),
 * [left halo | right halo] per row. */
void orc_noise_halo(uint64_t seed, int plane, int field, uint64_t step, int Ny, int Nz_max, double* out) {
    uint64_t len = (uint64_t)Ny * 2u * (uint64_t)Nz_max;
    orc_noise_elements(seed, stream_of(plane, field, 1), step, len, 0, len, out);
}
