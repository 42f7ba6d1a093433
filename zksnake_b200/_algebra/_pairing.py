"""Optimal-ate pairing on BN254 and BLS12-381 in plain Python integers.

CPU support code only: the reference's `pairing` / `multi_pairing` (/root/reference/src/bn254/curve.rs:417-437) are used by
Groth16.verify / Plonk.verify, never by a prover, so they stay on the host.  Fq12 is represented as Fq[w]/(w^12 - c6 w^6 - c0)
(w^6 = xi, u = w^6 - a), G2 points are untwisted into E(Fq12) and the Miller loop uses affine line functions; simple rather
than fast (about a second per pairing).
"""

_PARAMS = {
    0: dict(
        q=21888242871839275222246405745257275088696311157297823662689037894645226208583,
        r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
        xi_a=9,                       # xi = 9 + u  ->  u = w^6 - 9, w^12 = 18 w^6 - 82
        loop=29793968203157093288,    # 6x + 2
        bn=True, mtwist=False,
    ),
    1: dict(
        q=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
        r=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
        xi_a=1,                       # xi = 1 + u  ->  u = w^6 - 1, w^12 = 2 w^6 - 2
        loop=0xD201000000010000,      # |x|
        bn=False, mtwist=True,
    ),
}


class _Fq12:
    def __init__(self, curve):
        P = _PARAMS[curve]
        self.q = P["q"]
        a = P["xi_a"]
        # w^12 = 2a w^6 - (a^2 + 1)
        self.c6 = 2 * a % self.q
        self.c0 = (-(a * a + 1)) % self.q
        self.one = [1] + [0] * 11
        self.zero = [0] * 12

    def add(self, x, y):
        q = self.q
        return [(a + b) % q for a, b in zip(x, y)]

    def sub(self, x, y):
        q = self.q
        return [(a - b) % q for a, b in zip(x, y)]

    def neg(self, x):
        q = self.q
        return [(-a) % q for a in x]

    def mul(self, x, y):
        q = self.q
        t = [0] * 23
        for i, a in enumerate(x):
            if a:
                for j, b in enumerate(y):
                    t[i + j] += a * b
        for k in range(22, 11, -1):
            v = t[k]
            if v:
                t[k - 6] += v * self.c6
                t[k - 12] += v * self.c0
        return [v % q for v in t[:12]]

    def scalar(self, x, k):
        q = self.q
        return [a * k % q for a in x]

    def pow(self, x, e):
        res = self.one
        base = x
        while e:
            if e & 1:
                res = self.mul(res, base)
            base = self.mul(base, base)
            e >>= 1
        return res

    def inv(self, x):
        """Extended Euclid on polynomials over Fq modulo m(w) = w^12 - c6 w^6 - c0."""
        q = self.q
        m = [(-self.c0) % q] + [0] * 5 + [(-self.c6) % q] + [0] * 5 + [1]

        def deg(p):
            d = len(p) - 1
            while d >= 0 and p[d] == 0:
                d -= 1
            return d

        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = list(x) + [0], m
        while deg(low) > 0:
            # r = high / low (polynomial quotient)
            dl = deg(low)
            temp = list(high)
            quo = [0] * 13
            inv_lead = pow(low[dl], -1, q)
            for i in range(deg(temp) - dl, -1, -1):
                c = temp[dl + i] * inv_lead % q
                quo[i] = c
                if c:
                    for j in range(dl + 1):
                        temp[i + j] = (temp[i + j] - c * low[j]) % q
            nm = list(hm)
            for i in range(13):
                if lm[i]:
                    for j in range(13 - i):
                        if quo[j]:
                            nm[i + j] = (nm[i + j] - lm[i] * quo[j]) % q
            lm, low, hm, high = nm, temp, lm, low
        c = pow(low[0], -1, q)
        return [v * c % q for v in lm[:12]]

    def from_fq2(self, a, xi_a):
        """a0 + a1 u  with u = w^6 - xi_a."""
        out = [0] * 12
        out[0] = (a[0] - xi_a * a[1]) % self.q
        out[6] = a[1] % self.q
        return out

    def from_fq(self, a):
        return [a % self.q] + [0] * 11


def _untwist(curve, Q, F):
    P = _PARAMS[curve]
    x = F.from_fq2(Q[0], P["xi_a"])
    y = F.from_fq2(Q[1], P["xi_a"])
    w2 = [0, 0, 1] + [0] * 9
    w3 = [0, 0, 0, 1] + [0] * 8
    if P["mtwist"]:
        w2, w3 = F.inv(w2), F.inv(w3)
    return (F.mul(x, w2), F.mul(y, w3))


def _line(F, P1, P2, T):
    """Line through P1, P2 (E(Fq12), affine) evaluated at T."""
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    elif y1 == y2:
        m = F.mul(F.scalar(F.mul(x1, x1), 3), F.inv(F.scalar(y1, 2)))
    else:
        return F.sub(xt, x1)
    return F.sub(F.mul(m, F.sub(xt, x1)), F.sub(yt, y1))


def _add(F, P1, P2):
    if P1 is None:
        return P2
    if P2 is None:
        return P1
    x1, y1 = P1
    x2, y2 = P2
    if x1 == x2:
        if y1 != y2:
            return None
        m = F.mul(F.scalar(F.mul(x1, x1), 3), F.inv(F.scalar(y1, 2)))
    else:
        m = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(m, m), x1), x2)
    return (x3, F.sub(F.mul(m, F.sub(x1, x3)), y1))


def miller_loop(curve, Pt, Q):
    """f_{loop,Q}(P) before the final exponentiation; Pt in G1 (ints), Q in G2 (Fq2 pairs)."""
    P = _PARAMS[curve]
    F = _Fq12(curve)
    if Pt is None or Q is None:
        return F.one
    Qw = _untwist(curve, Q, F)
    Pw = (F.from_fq(Pt[0]), F.from_fq(Pt[1]))
    R = Qw
    f = F.one
    for i in range(P["loop"].bit_length() - 2, -1, -1):
        f = F.mul(F.mul(f, f), _line(F, R, R, Pw))
        R = _add(F, R, R)
        if (P["loop"] >> i) & 1:
            f = F.mul(f, _line(F, R, Qw, Pw))
            R = _add(F, R, Qw)
    if P["bn"]:
        q = P["q"]
        Q1 = (F.pow(Qw[0], q), F.pow(Qw[1], q))
        nQ2 = (F.pow(Q1[0], q), F.neg(F.pow(Q1[1], q)))
        f = F.mul(f, _line(F, R, Q1, Pw))
        R = _add(F, R, Q1)
        f = F.mul(f, _line(F, R, nQ2, Pw))
    return f


def final_exponentiation(curve, f):
    P = _PARAMS[curve]
    F = _Fq12(curve)
    return tuple(F.pow(f, (P["q"] ** 12 - 1) // P["r"]))


def pairing(curve, Pt, Q):
    return final_exponentiation(curve, miller_loop(curve, Pt, Q))


def multi_pairing(curve, Ps, Qs):
    F = _Fq12(curve)
    f = F.one
    for Pt, Q in zip(Ps, Qs):
        f = F.mul(f, miller_loop(curve, Pt, Q))
    return final_exponentiation(curve, f)
