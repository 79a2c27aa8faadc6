/* oracle/pcg32_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the one pcg-cpp engine the
 * reference instantiates: pcg32 = setseq_xsh_rr_64_32 (pcg_random.hpp:1866), used at df.cpp:334.
 * Pinned by the reference's own known-answer files pcg-cpp/test-high/expected/check-pcg32.out and
 * check-pcg32_oneseq.out (tests/test_oracle_pcg32.py re-creates the whole test program output).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "dfb_rng_spec.h"

typedef struct { uint64_t state, inc; } orc_pcg32;

/* LCG step, pcg_random.hpp:413-416 (bump): state*mult + inc */
static uint64_t lcg(uint64_t s, uint64_t inc) { return s * DFB_PCG32_MULT + inc; }

/* XSH-RR 64->32 output, pcg_random.hpp:845-872: xorshift high bits, random rotate by top 5 bits */
static uint32_t xsh_rr(uint64_t s) {
    uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27);
    uint32_t r = (uint32_t)(s >> 59);
    return (x >> r) | (x << ((32u - r) & 31u));
}

/* two-arg seeding pcg32(seed, stream): inc = (stream<<1)|1 (pcg_random.hpp:263-267),
 * state = bump(seed + inc) (pcg_random.hpp:497-503) */
void orc_pcg32_seed(orc_pcg32* g, uint64_t seed, uint64_t stream) {
    g->inc = (stream << 1) | 1u;
    g->state = lcg(seed + g->inc, g->inc);
}
/* one-arg seeding pcg32{seed}: default increment (pcg_random.hpp:162,484-489) -- df.cpp:334's form */
void orc_pcg32_seed1(orc_pcg32* g, uint64_t seed) {
    g->inc = DFB_PCG32_DEFAULT_INC;
    g->state = lcg(seed + g->inc, g->inc);
}
/* operator(): output of the PRE-advance state (output_previous = true for 64-bit state,
 * pcg_random.hpp:427-437) */
uint32_t orc_pcg32_next(orc_pcg32* g) {
    uint64_t old = g->state;
    g->state = lcg(old, g->inc);
    return xsh_rr(old);
}
/* advance(delta), pcg_random.hpp:640-669 (Brown's arbitrary-stride algorithm) */
void orc_lcg_jump(uint64_t delta, uint64_t inc, uint64_t* A, uint64_t* C) {
    uint64_t acc_mult = 1u, acc_plus = 0u, cur_mult = DFB_PCG32_MULT, cur_plus = inc;
    while (delta > 0) {
        if (delta & 1u) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1u) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    *A = acc_mult; *C = acc_plus;
}
void orc_pcg32_advance(orc_pcg32* g, uint64_t delta) {
    uint64_t A, C;
    orc_lcg_jump(delta, g->inc, &A, &C);
    g->state = A * g->state + C;
}
/* backstep(delta) = advance(-delta), pcg_random.hpp:462-465 */
void orc_pcg32_backstep(orc_pcg32* g, uint64_t delta) { orc_pcg32_advance(g, (uint64_t)0 - delta); }

/* distance, pcg_random.hpp:671-693 (non-MCG branch: the increment is odd, never 0) */
uint64_t orc_pcg32_distance(const orc_pcg32* from, const orc_pcg32* to) {
    uint64_t cur = from->state, cur_mult = DFB_PCG32_MULT, cur_plus = from->inc;
    uint64_t the_bit = 1u, dist = 0u;
    while (cur != to->state) {
        if ((cur & the_bit) != (to->state & the_bit)) { cur = cur * cur_mult + cur_plus; dist |= the_bit; }
        the_bit <<= 1;
        cur_plus = (cur_mult + 1u) * cur_plus;
        cur_mult *= cur_mult;
    }
    return dist;
}
/* bounded_rand, pcg_extras.hpp:540-552 */
uint32_t orc_pcg32_bounded(orc_pcg32* g, uint32_t bound) {
    uint32_t threshold = (uint32_t)(0u - bound) % bound;
    for (;;) { uint32_t r = orc_pcg32_next(g); if (r >= threshold) return r % bound; }
}

/* flat helpers for ctypes */
void orc_pcg32_draw(uint64_t seed, uint64_t stream, int has_stream, uint64_t delta, int n, uint32_t* out) {
    orc_pcg32 g;
    if (has_stream) orc_pcg32_seed(&g, seed, stream); else orc_pcg32_seed1(&g, seed);
    orc_pcg32_advance(&g, delta);
    for (int i = 0; i < n; ++i) out[i] = orc_pcg32_next(&g);
}
void orc_pcg32_state(uint64_t seed, uint64_t stream, int has_stream, uint64_t* state, uint64_t* inc) {
    orc_pcg32 g;
    if (has_stream) orc_pcg32_seed(&g, seed, stream); else orc_pcg32_seed1(&g, seed);
    *state = g.state; *inc = g.inc;
}

/* Re-creation of pcg-cpp/test-high/pcg-test.cpp:56-171 for RNG = pcg32 (TWO_ARG_INIT, check-pcg32.cpp)
 * or pcg32_oneseq-equivalent seeding (one-arg; see SURVEY 8c: identical outputs to pcg32{42}).
 * Writes the text the program prints for `rounds` rounds, without the 5-line banner whose "size"
 * and "period" lines differ between the two typedefs.  Returns bytes written. */
static int put(char* buf, int cap, int pos, const char* s) {
    int n = (int)strlen(s);
    if (pos + n < cap) memcpy(buf + pos, s, (size_t)n + 1);
    return pos + n;
}
int orc_pcg32_kat_text(int two_arg, int rounds, char* buf, int cap) {
    orc_pcg32 g;
    char tmp[64];
    int pos = 0;
    if (two_arg) orc_pcg32_seed(&g, 42u, 54u); else orc_pcg32_seed1(&g, 42u);
    for (int round = 1; round <= rounds; ++round) {
        snprintf(tmp, sizeof tmp, "Round %d:\n", round); pos = put(buf, cap, pos, tmp);
        pos = put(buf, cap, pos, "  32bit:");
        for (int i = 0; i < 6; ++i) { snprintf(tmp, sizeof tmp, " 0x%08x", orc_pcg32_next(&g)); pos = put(buf, cap, pos, tmp); }
        pos = put(buf, cap, pos, "\n  Again:");
        orc_pcg32_backstep(&g, 6);                                   /* pcg-test.cpp:122 */
        for (int i = 0; i < 6; ++i) { snprintf(tmp, sizeof tmp, " 0x%08x", orc_pcg32_next(&g)); pos = put(buf, cap, pos, tmp); }
        pos = put(buf, cap, pos, "\n  Coins: ");
        for (int i = 0; i < 65; ++i) pos = put(buf, cap, pos, orc_pcg32_bounded(&g, 2) ? "H" : "T");
        pos = put(buf, cap, pos, "\n");
        orc_pcg32 copy = g;
        pos = put(buf, cap, pos, "  Rolls:");
        for (int i = 0; i < 33; ++i) { snprintf(tmp, sizeof tmp, " %u", orc_pcg32_bounded(&g, 6) + 1u); pos = put(buf, cap, pos, tmp); }
        snprintf(tmp, sizeof tmp, "\n   -->   rolling dice used %llu random numbers\n",
                 (unsigned long long)orc_pcg32_distance(&copy, &g));   /* pcg-test.cpp:143 */
        pos = put(buf, cap, pos, tmp);
        /* pcg_extras::shuffle, pcg_extras.hpp:554-566 */
        char cards[52];
        for (int i = 0; i < 52; ++i) cards[i] = (char)i;
        { int count = 52; char* to = cards + 52;
          while (count > 1) { int chosen = (int)orc_pcg32_bounded(&g, (uint32_t)count); --count; --to;
                              char t = cards[chosen]; cards[chosen] = *to; *to = t; } }
        static const char number[] = "A23456789TJQK", suit[] = "hcds";
        pos = put(buf, cap, pos, "  Cards:");
        for (int i = 0; i < 52; ++i) {
            snprintf(tmp, sizeof tmp, " %c%c", number[cards[i] / 4], suit[cards[i] % 4]); pos = put(buf, cap, pos, tmp);
            if ((i + 1) % 22 == 0) pos = put(buf, cap, pos, "\n\t");
        }
        pos = put(buf, cap, pos, "\n\n");
    }
    return pos;
}
