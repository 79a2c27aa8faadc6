"""tools/gen_rng_tables.py -- generates the 64-entry logarithm table of include/dfb_rng_spec.h (spec v2) with 60-digit arithmetic.

For the top six mantissa bits i of d = (double)U1 (mantissa m in [1,2)):
    i < 32:   m' = m      in [1, 1.5),    centre c = 1 + (i + 0.5)/64
    i >= 32:  m' = m / 2  in [0.75, 1),   centre c = (1 + (i + 0.5)/64) / 2
INV[i] = RN(1/c) (binary64), L[i] = RN(-ln(INV[i])) with INV[i] taken at its rounded value, so that
ln(m') = L[i] + log1p(r), r = fma(m', INV[i], -1), |r| <= 2^-7 (1 + 2^-6), holds to the last bit of L[i].
Prints the C initialiser list (hex floats)."""
from decimal import Decimal, getcontext
from fractions import Fraction
import struct

getcontext().prec = 60


def rn(x):            # Decimal -> nearest binary64 (via exact Fraction rounding)
    f = Fraction(x)
    d = float(f)      # Python rounds Fraction -> float correctly (round-half-even)
    return d


rows = []
for i in range(64):
    c = Fraction(2 * 64 + 2 * i + 1, 128)
    if i >= 32:
        c = c / 2
    inv = float(1 / c)                        # correctly rounded
    L = rn(-(Decimal(Fraction(inv).numerator) / Decimal(Fraction(inv).denominator)).ln())
    rows.append((inv, L))
print("#define DFB_LOGTAB_LIST \\")
for i, (inv, L) in enumerate(rows):
    print("    %s, %s%s" % (inv.hex(), L.hex(), ", \\" if i < 63 else ""))
