"""Exact-arithmetic proof that the collision kernel's three-instruction division by a constant
(sgc_lb.cu, divc<K>:  q = RN(x*rc);  r = fma(-q, c, x);  result = fma(r, rc, q),  rc = RN(1/c))
returns the correctly rounded x / c for EVERY normal double x (away from over/underflow, which the
kernel routes to the real division) for the four constants of LBPhysics::collision.

Argument.  Let Q = x/c, delta = |rc*c - 1|.  q is within beta = 1/2 + delta*2^53 ulp of Q, so
r = x - q*c is a multiple of ulp(q)*ulp(c) smaller than beta*C ulps (C = the 53-bit mantissa of c),
hence exact when beta*C < 2^53.  The value rounded by the last fma is v = q + r*rc and
|v - Q| = |r|*|rc - 1/c| <= beta*delta ulp =: tau ulp.  RN(v) can differ from RN(Q) only if a rounding
boundary m = M*2^(k-53) (M odd) lies within 2*tau ulp of Q, i.e. |X*2^s - M*C| <= 4*tau*C (X the
mantissa of x) — a linear congruence in M with a handful of solutions, which check_constant()
enumerates and runs through the algorithm with exact rational arithmetic (both signs)."""
from fractions import Fraction as Fr
import math

def decomp(c):
    m, e = math.frexp(c)            # c = m * 2^e, 0.5 <= m < 1
    C = int(m * (1 << 53)); ec = e - 53
    assert Fr(C) * Fr(2) ** ec == Fr(c)
    return C, ec

def fast_div(X, c, rc):
    x = Fr(X)
    q0 = float(x * Fr(rc))
    r = x - Fr(q0) * Fr(c)
    rf = float(r)
    exact = Fr(rf) == r
    q1 = float(Fr(q0) + Fr(rf) * Fr(rc))
    return q1, exact

def check_constant(name, c, dmax_override=None):
    C, ec = decomp(c)
    rc = 1.0 / c
    delta = abs(Fr(rc) * Fr(c) - 1)
    beta = delta * 2 ** 53 + Fr(1, 2)
    assert beta * C < 2 ** 53, "r might not be exact"
    tau = beta * delta
    Dmax = int(4 * tau * C) + 1 if dmax_override is None else dmax_override
    t = (C & -C).bit_length() - 1
    cands = set()
    # Q = X / c for X in [2^52, 2^53): Q*2^ec*... find binades k
    qlo, qhi = Fr(2 ** 52) / Fr(c), Fr(2 ** 53) / Fr(c)
    klo, khi = math.floor(math.log2(qlo)), math.floor(math.log2(qhi))
    for k in range(klo - 1, khi + 2):
        s = -(k - 53 + ec)
        if s <= 0:
            continue
        for D in range(-Dmax, Dmax + 1):
            if D % (1 << t):
                continue
            mod = 1 << (s - t)
            Cr = C >> t
            base = (-(D >> t) * pow(Cr, -1, mod)) % mod if mod > 1 else 0
            M = base
            # M odd in (2^53, 2^54)
            j0 = ((1 << 53) - base) // mod
            for j in range(j0 - 1, j0 + (1 << 54) // mod + 3):
                M = base + j * mod
                if M <= (1 << 53) or M >= (1 << 54) or M % 2 == 0:
                    continue
                num = M * C + D
                if num % (1 << s):
                    continue
                X = num >> s
                if (1 << 52) <= X < (1 << 53):
                    cands.add(X)
    bad = []
    for X in sorted(cands):
        for sign in (1, -1):
            q1, exact = fast_div(sign * X, c, rc)
            want = float(Fr(sign * X) / Fr(c))
            if q1 != want or not exact:
                bad.append((X, sign, q1, want, exact))
    return sorted(cands), bad



CS2 = 1.0 / 3                                   # LBParameters::g_cs2
CONSTANTS = {"cs2": CS2, "two_cs2": 2 * CS2, "two_cs2_sq": 2 * CS2 * CS2, "tau": 0.516}
