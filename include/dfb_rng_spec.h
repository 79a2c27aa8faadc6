/* include/dfb_rng_spec.h -- the white-noise specification of the B200 digital filter (spec v1).
 *
 * The reference draws its noise from ONE process-wide `static pcg32 rng{random_device{}()}` through
 * libstdc++'s std::normal_distribution (df.cpp:334-339): a Marsaglia-polar rejection loop with a cached
 * second variate, i.e. a data-dependent stream position that cannot be reproduced by jump-ahead and
 * is not even reproducible run to run ("Bit-exact RNG gate").  north_star
 * therefore REPLACES it: per-element pcg32 streams addressed with advance() (pcg_random.hpp:457-460,
 * 640-669) and a normal transform built only from IEEE-754 correctly rounded operations
 * (+, *, fma, /, sqrt, integer ops), so that the device kernel and a plain-C host restatement
 * (oracle/normal_oracle.c) agree BIT FOR BIT.  This header holds the constants and the contract;
 * the arithmetic is written twice, independently (device: csrc/noise.cuh, host: oracle/).
 *
 * Streams.  For plane p, field f (0=u,1=v,2=w) and array a (0 = r_ys, 1 = r_zs halo columns)
 *     stream = ((p*3 + f)*2 + a),  rng = pcg32(seed, stream)          (pcg_random.hpp:484-503)
 *     inc    = (stream << 1) | 1,  state0 = (seed + inc) * M + inc,   M = 6364136223846793005
 * Positions.  The logical array of step s (s = 0 is the constructor's first field, df.cpp:57) has
 * `len` elements indexed by the GLOBAL flat index e (r_ys: e = row*Nz_global + k, rows
 * 0..Ny+2*Ny_max-1, df.cpp:197;  halo: e = j*(2*Nz_max) + h, h < Nz_max = left halo, else right,
 * the only part of r_zs that is ever read, df.cpp:157,398).  Element pair q = e >> 1 consumes the
 * four draws at stream positions 4*(s*npairs + q) .. +3, npairs = (len+1)/2.  Because a position
 * depends only on (seed, plane, field, array, step, global index), any spanwise slab of the plane
 * can regenerate exactly the noise its neighbour would have used: no halo exchange.
 *
 * Pair transform (o0..o3 = the four 32-bit outputs, in stream order):
 *     U1 = ((o1:o0) >> 11) + 1   in [1, 2^53]        u1 = U1 * 2^-53 in (0,1]
 *     U2 =  (o3:o2) >> 11        in [0, 2^53)        u2 = U2 * 2^-53 in [0,1)
 *     r  = sqrt(-2 ln u1),  z0 = r cos(2 pi u2) -> element 2q,  z1 = r sin(2 pi u2) -> element 2q+1
 * ln u1:   d = (double)U1 (exact); E = unbiased exponent, m = mantissa in [1,2);
 *          if (m > SQRT2) { m *= 0.5; E += 1; }   f = m - 1;  s = f / (2 + f);  z = s*s;
 *          P = Horner over DFB_LOG_C[9..0] in z with fma;  lnm = fma(s*z, P, 2*s);
 *          lnu = fma((double)(E - 53), LN2, lnm);            r = sqrt(-2 * lnu)
 * sincos:  oct = U2 >> 50;  T = U2 & (2^50-1);  if (oct & 1) T = 2^50 - T;  t = T * 2^-50 (exact);
 *          x = t * PIO4;  x2 = x*x;
 *          S = Horner over DFB_SIN_C[7..0] in x2 with fma;  sx = fma(x*x2, S, x)
 *          C = Horner over DFB_COS_C[7..0] in x2 with fma;  cx = fma(x2, C, 1.0)
 *          if (oct & 1) swap(sx, cx);      (phi = pi/2 - x inside the quadrant)
 *          quadrant = oct >> 1:  0:(sin,cos)=(sx,cx)  1:(cx,-sx)  2:(-sx,-cx)  3:(-cx,sx)
 * Every operation above is a single correctly rounded IEEE-754 binary64 operation; no contraction
 * other than the fma()s written out.
 */
#ifndef DFB_RNG_SPEC_H
#define DFB_RNG_SPEC_H

#define DFB_PCG32_MULT 6364136223846793005ULL   /* pcg_random.hpp:158 */
#define DFB_PCG32_DEFAULT_INC 1442695040888963407ULL /* pcg_random.hpp:162, one-arg ctor */

#define DFB_LN2   0x1.62e42fefa39efp-1
#define DFB_SQRT2 0x1.6a09e667f3bcdp+0
#define DFB_PIO4  0x1.921fb54442d18p-1

/* 2/(2k+3), k = 0..9 : ln(m) = 2s + s z (c0 + c1 z + ...),  z = s^2 */
#define DFB_LOG_C_LIST \
    0x1.5555555555555p-1, 0x1.999999999999ap-2, 0x1.2492492492492p-2, 0x1.c71c71c71c71cp-3, \
    0x1.745d1745d1746p-3, 0x1.3b13b13b13b14p-3, 0x1.1111111111111p-3, 0x1.e1e1e1e1e1e1ep-4, \
    0x1.af286bca1af28p-4, 0x1.8618618618618p-4
/* (-1)^(k+1)/(2k+3)!, k = 0..7 : sin x = x + x x2 (c0 + c1 x2 + ...) */
#define DFB_SIN_C_LIST \
    -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19, \
    -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49
/* (-1)^(k+1)/(2k+2)!, k = 0..7 : cos x = 1 + x2 (c0 + c1 x2 + ...) */
#define DFB_COS_C_LIST \
    -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16, \
    -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45

#define DFB_LOG_NC 10
#define DFB_SIN_NC 8
#define DFB_COS_NC 8

#endif
