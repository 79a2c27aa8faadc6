/* include/dfb_rng_spec.h -- the white-noise specification of the B200 digital filter (spec v2).
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
 *     U1 = ((o1:o0) >> 11) | 1   odd, in [1, 2^53)   u1 = U1 * 2^-53 in (0,1)
 *     U2 =  (o3:o2) >> 11        in [0, 2^53)        u2 = U2 * 2^-53 in [0,1)
 *     r  = sqrt(-2 ln u1),  z0 = r cos(2 pi u2) -> element 2q,  z1 = r sin(2 pi u2) -> element 2q+1
 * ln u1 (v2: table-driven, no division):
 *          d = (double)U1 (exact); E = unbiased exponent, m = mantissa in [1,2); i = top six mantissa bits; hi = i >> 5;
 *          m' = m * 2^-hi in [0.75, 1.5)  (an exponent-field edit);   e = E + hi - 53  (<= 0)
 *          r = fma(m', INV[i], -1)        |r| <= 2^-7 (1 + 2^-6):  INV[i] = RN(1/c_i), c_i = centre of m' over interval i
 *          q = Horner over DFB_LOG1P_C[5..0] in r with fma;  lp = fma(r*r, q, r)          = log1p(r), error < 2e-18
 *          lnu = fma((double)e, LN2, L[i] + lp)   L[i] = RN(-ln INV[i]);                  rad = sqrt(-2 * lnu)
 *          (u1 < 1 strictly and every error term is far below 2^-53, so lnu < 0: no clamp)
 * sincos:  oct = U2 >> 50;  T = U2 & (2^50-1);  if (oct & 1) T = 2^50 - T;  t = T * 2^-50 (exact);
 *          x = t * PIO4;  x2 = x*x;
 *          S = Horner over DFB_SIN_C[6..0] in x2 with fma;  sx = fma(x*x2, S, x)
 *          C = Horner over DFB_COS_C[7..0] in x2 with fma;  cx = fma(x2, C, 1.0)
 *          if (oct & 1) swap(sx, cx);      (phi = pi/2 - x inside the quadrant)
 *          quadrant = oct >> 1:  0:(sin,cos)=(sx,cx)  1:(cx,-sx)  2:(-sx,-cx)  3:(-cx,sx)
 * Every operation above is a single correctly rounded IEEE-754 binary64 operation; no contraction
 * other than the fma()s written out.  (v1 used an atanh-series logarithm with a division and ten terms, and eight sine terms:
 * 256 -> ~190 instructions per pair on the device; the streams and positions are unchanged.)
 */
#ifndef DFB_RNG_SPEC_H
#define DFB_RNG_SPEC_H

#define DFB_PCG32_MULT 6364136223846793005ULL   /* pcg_random.hpp:158 */
#define DFB_PCG32_DEFAULT_INC 1442695040888963407ULL /* pcg_random.hpp:162, one-arg ctor */

#define DFB_LN2   0x1.62e42fefa39efp-1
#define DFB_PIO4  0x1.921fb54442d18p-1

/* (-1)^(k+1)/k, k = 2..7 : log1p(r) = r + r^2 (c2 + c3 r + ... + c7 r^5) */
#define DFB_LOG1P_C_LIST \
    -0x1.0000000000000p-1, 0x1.5555555555555p-2, -0x1.0000000000000p-2, 0x1.999999999999ap-3, \
    -0x1.5555555555555p-3, 0x1.2492492492492p-3
#define DFB_LOG1P_NC 6
/* per interval i of the top six mantissa bits: INV[i] = RN(1/c_i), L[i] = RN(-ln INV[i])  (tools/gen_rng_tables.py, 60-digit arithmetic) */
#define DFB_LOGTAB_LIST \
    0x1.fc07f01fc07f0p-1, 0x1.fe02a6b106799p-8, \
    0x1.f44659e4a4271p-1, 0x1.7b91b07d5b126p-6, \
    0x1.ecc07b301ecc0p-1, 0x1.39e87b9febd68p-5, \
    0x1.e573ac901e574p-1, 0x1.b42dd711971b9p-5, \
    0x1.de5d6e3f8868ap-1, 0x1.16536eea37ae3p-4, \
    0x1.d77b654b82c34p-1, 0x1.51b073f06183cp-4, \
    0x1.d0cb58f6ec074p-1, 0x1.8c345d6319b23p-4, \
    0x1.ca4b3055ee191p-1, 0x1.c5e548f5bc743p-4, \
    0x1.c3f8f01c3f8f0p-1, 0x1.fec9131dbeabcp-4, \
    0x1.bdd2b899406f7p-1, 0x1.1b72ad52f67a2p-3, \
    0x1.b7d6c3dda338bp-1, 0x1.371fc201e8f75p-3, \
    0x1.b2036406c80d9p-1, 0x1.526e5e3a1b438p-3, \
    0x1.ac5701ac5701bp-1, 0x1.6d60fe719d21bp-3, \
    0x1.a6d01a6d01a6dp-1, 0x1.87fa06520c911p-3, \
    0x1.a16d3f97a4b02p-1, 0x1.a23bc1fe2b561p-3, \
    0x1.9c2d14ee4a102p-1, 0x1.bc286742d8cd4p-3, \
    0x1.970e4f80cb872p-1, 0x1.d5c216b4fbb94p-3, \
    0x1.920fb49d0e229p-1, 0x1.ef0adcbdc5935p-3, \
    0x1.8d3018d3018d3p-1, 0x1.0402594b4d041p-2, \
    0x1.886e5f0abb04ap-1, 0x1.1058bf9ae4ad4p-2, \
    0x1.83c977ab2beddp-1, 0x1.1c898c16999fbp-2, \
    0x1.7f405fd017f40p-1, 0x1.2895a13de86a4p-2, \
    0x1.7ad2208e0ecc3p-1, 0x1.347dd9a987d56p-2, \
    0x1.767dce434a9b1p-1, 0x1.404308686a7e4p-2, \
    0x1.724287f46debcp-1, 0x1.4be5f957778a1p-2, \
    0x1.6e1f76b4337c7p-1, 0x1.5767717455a6cp-2, \
    0x1.6a13cd1537290p-1, 0x1.62c82f2b9c796p-2, \
    0x1.661ec6a5122f9p-1, 0x1.6e08eaa2ba1e4p-2, \
    0x1.623fa77016240p-1, 0x1.792a55fdd47a1p-2, \
    0x1.5e75bb8d015e7p-1, 0x1.842d1da1e8b18p-2, \
    0x1.5ac056b015ac0p-1, 0x1.8f11e873662c8p-2, \
    0x1.571ed3c506b3ap-1, 0x1.99d958117e08ap-2, \
    0x1.5390948f40febp+0, -0x1.214456d0eb8d5p-2, \
    0x1.5015015015015p+0, -0x1.16b5ccbacfb73p-2, \
    0x1.4cab88725af6ep+0, -0x1.0c42d676162e2p-2, \
    0x1.49539e3b2d067p+0, -0x1.01eae5626c691p-2, \
    0x1.460cbc7f5cf9ap+0, -0x1.ef5ade4dcffe5p-3, \
    0x1.42d6625d51f87p+0, -0x1.db13db0d48941p-3, \
    0x1.3fb013fb013fbp+0, -0x1.c6ffbc6f00f71p-3, \
    0x1.3c995a47babe7p+0, -0x1.b31d8575bce3bp-3, \
    0x1.3991c2c187f63p+0, -0x1.9f6c407089663p-3, \
    0x1.3698df3de0748p+0, -0x1.8beafeb38fe8fp-3, \
    0x1.33ae45b57bcb2p+0, -0x1.7898d85444c74p-3, \
    0x1.30d190130d190p+0, -0x1.6574ebe8c1339p-3, \
    0x1.2e025c04b8097p+0, -0x1.527e5e4a1b58dp-3, \
    0x1.2b404ad012b40p+0, -0x1.3fb45a59928cap-3, \
    0x1.288b01288b013p+0, -0x1.2d1610c86813dp-3, \
    0x1.25e22708092f1p+0, -0x1.1aa2b7e23f729p-3, \
    0x1.23456789abcdfp+0, -0x1.08598b59e3a07p-3, \
    0x1.20b470c67c0d9p+0, -0x1.ec739830a1126p-4, \
    0x1.1e2ef3b3fb874p+0, -0x1.c885801bc4b20p-4, \
    0x1.1bb4a4046ed29p+0, -0x1.a4e7640b1bc38p-4, \
    0x1.19453808ca29cp+0, -0x1.8197e2f40e3f0p-4, \
    0x1.16e0689427379p+0, -0x1.5e95a4d9791cdp-4, \
    0x1.1485f0e0acd3bp+0, -0x1.3bdf5a7d1ee5ep-4, \
    0x1.12358e75d3033p+0, -0x1.1973bd1465561p-4, \
    0x1.0fef010fef011p+0, -0x1.eea31c006b87cp-5, \
    0x1.0db20a88f4696p+0, -0x1.aaef2d0fb1108p-5, \
    0x1.0b7e6ec259dc8p+0, -0x1.67c94f2d4bb65p-5, \
    0x1.0953f39010954p+0, -0x1.252f32f8d1840p-5, \
    0x1.073260a47f7c6p+0, -0x1.c63d2ec14aad7p-6, \
    0x1.05197f7d73404p+0, -0x1.432a925980cbcp-6, \
    0x1.03091b51f5e1ap+0, -0x1.82448a388a283p-7, \
    0x1.0101010101010p+0, -0x1.0080559588b25p-8

/* (-1)^(k+1)/(2k+3)!, k = 0..6 : sin x = x + x x2 (c0 + c1 x2 + ...);  the next term is below 6e-17 x for |x| <= pi/4 */
#define DFB_SIN_C_LIST \
    -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19, \
    -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41
/* (-1)^(k+1)/(2k+2)!, k = 0..7 : cos x = 1 + x2 (c0 + c1 x2 + ...) */
#define DFB_COS_C_LIST \
    -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16, \
    -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45

#define DFB_SIN_NC 7
#define DFB_COS_NC 8

#endif
