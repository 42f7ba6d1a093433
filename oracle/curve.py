"""G1 / G2 group oracle on BN254 and BLS12-381 (pure Python ints; test infrastructure only).

Restates what /root/reference/src/bn254/curve.rs (bls12_381 twin: src/bls12_381/curve.rs) asks
arkworks (ark-ec 0.4.2, ark-serialize 0.4.2; not vendored) to do:
  PointG1/PointG2 constructors and getters  curve.rs:28-56, :197-232 (x,y = 0 for infinity)
  + - neg * == is_zero                      curve.rs:74-118, :249-292
  to_bytes / from_bytes (compressed)        curve.rs:120-145, :298-323
  multiscalar_mul_g1 / _g2                  curve.rs:356-392 (length mismatch -> ValueError)
  batch_multi_scalar_g1 / _g2               curve.rs:326-354
Points are None (infinity) or (x, y); G1 coordinates are ints, G2 coordinates are (c0, c1) pairs over
Fq[u]/(u^2+1).  Short-Weierstrass, a = 0 on both curves.
"""
from .fields import PARAMS, BN254, BLS12_381


class Fq:
    def __init__(self, q):
        self.q = q
        self.zero, self.one = 0, 1

    def add(self, a, b): return (a + b) % self.q
    def sub(self, a, b): return (a - b) % self.q
    def mul(self, a, b): return a * b % self.q
    def neg(self, a): return (-a) % self.q
    def inv(self, a): return pow(a, -1, self.q)
    def is_zero(self, a): return a % self.q == 0
    def small(self, k): return k % self.q

    def sqrt(self, a):
        q = self.q
        assert q % 4 == 3
        s = pow(a, (q + 1) // 4, q)
        return s if s * s % q == a % q else None

    def lex_larger(self, y):
        """y > -y as integers."""
        return y > (self.q - y) % self.q


class Fq2:
    """Fq[u]/(u^2+1) -- both curves use non-residue -1."""

    def __init__(self, q):
        self.q = q
        self.zero, self.one = (0, 0), (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.q, (a[1] + b[1]) % self.q)
    def sub(self, a, b): return ((a[0] - b[0]) % self.q, (a[1] - b[1]) % self.q)
    def neg(self, a): return ((-a[0]) % self.q, (-a[1]) % self.q)
    def is_zero(self, a): return a[0] % self.q == 0 and a[1] % self.q == 0
    def small(self, k): return (k % self.q, 0)

    def mul(self, a, b):
        q = self.q
        return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def inv(self, a):
        q = self.q
        d = pow(a[0] * a[0] + a[1] * a[1], -1, q)
        return (a[0] * d % q, (-a[1]) * d % q)

    def sqrt(self, a):
        """Square root in Fq2 (q = 3 mod 4) via the norm; returns one root or None."""
        q = self.q
        if self.is_zero(a):
            return (0, 0)
        a0, a1 = a[0] % q, a[1] % q
        if a1 == 0:
            s = Fq(q).sqrt(a0)
            if s is not None:
                return (s, 0)
            s = Fq(q).sqrt((-a0) % q)
            return (0, s) if s is not None else None
        n = Fq(q).sqrt((a0 * a0 + a1 * a1) % q)
        if n is None:
            return None
        inv2 = pow(2, -1, q)
        for cand in ((a0 + n) * inv2 % q, (a0 - n) * inv2 % q):
            x0 = Fq(q).sqrt(cand)
            if x0 is not None and x0 != 0:
                x1 = a1 * pow(2 * x0, -1, q) % q
                if self.mul((x0, x1), (x0, x1)) == (a0, a1):
                    return (x0, x1)
        return None


class Group:
    """One of the four groups; all formulas textbook affine / Jacobian for y^2 = x^3 + b."""

    def __init__(self, curve, g2):
        P = PARAMS[curve]
        self.curve, self.is_g2, self.P = curve, g2, P
        self.F = Fq2(P.q) if g2 else Fq(P.q)
        self.b = P.b_g2 if g2 else P.b_g1
        self.gen = P.g2 if g2 else P.g1
        self.r = P.r

    # ---- affine ----
    def on_curve(self, p):
        if p is None:
            return True
        F = self.F
        x, y = p
        return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), self.b)

    def neg(self, p):
        return None if p is None else (p[0], self.F.neg(p[1]))

    def add(self, p, q):
        F = self.F
        if p is None:
            return q
        if q is None:
            return p
        if p[0] == q[0]:
            if p[1] == q[1] and not F.is_zero(p[1]):
                lam = F.mul(F.mul(F.small(3), F.mul(p[0], p[0])), F.inv(F.add(p[1], p[1])))
            else:
                return None
        else:
            lam = F.mul(F.sub(q[1], p[1]), F.inv(F.sub(q[0], p[0])))
        x3 = F.sub(F.sub(F.mul(lam, lam), p[0]), q[0])
        return (x3, F.sub(F.mul(lam, F.sub(p[0], x3)), p[1]))

    # ---- Jacobian (for speed in scalar mul) ----
    def _jdbl(self, p):
        F = self.F
        X, Y, Z = p
        if F.is_zero(Z):
            return p
        A = F.mul(X, X); B = F.mul(Y, Y); C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C); D = F.add(D, D)
        E = F.add(F.add(A, A), A); Fq_ = F.mul(E, E)
        X3 = F.sub(Fq_, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd_affine(self, p, q):
        """Jacobian p + affine q (q not infinity)."""
        F = self.F
        X1, Y1, Z1 = p
        if F.is_zero(Z1):
            return (q[0], q[1], F.one)
        Z1Z1 = F.mul(Z1, Z1)
        U2 = F.mul(q[0], Z1Z1)
        S2 = F.mul(F.mul(q[1], Z1), Z1Z1)
        H = F.sub(U2, X1)
        R = F.sub(S2, Y1)
        if F.is_zero(H):
            if F.is_zero(R):
                return self._jdbl(p)
            return (F.one, F.one, F.zero)
        HH = F.mul(H, H); HHH = F.mul(H, HH); V = F.mul(X1, HH)
        X3 = F.sub(F.sub(F.mul(R, R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(Y1, HHH))
        Z3 = F.mul(Z1, H)
        return (X3, Y3, Z3)

    def _to_affine(self, p):
        F = self.F
        if F.is_zero(p[2]):
            return None
        zi = F.inv(p[2]); zi2 = F.mul(zi, zi)
        return (F.mul(p[0], zi2), F.mul(p[1], F.mul(zi2, zi)))

    def mul(self, p, k):
        """p * Fr::from(k) -- scalar reduced mod r as at curve.rs:103-108."""
        k %= self.r
        if p is None or k == 0:
            return None
        F = self.F
        acc = (F.one, F.one, F.zero)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == "1":
                acc = self._jadd_affine(acc, p)
        return self._to_affine(acc)

    def msm(self, points, scalars):
        """multiscalar_mul_g{1,2}: curve.rs:356-392 -- the definition, term by term."""
        if len(points) != len(scalars):
            raise ValueError("Number of points and scalars mismatch")
        F = self.F
        acc = None
        for p, s in zip(points, scalars):
            acc = self.add(acc, self.mul(p, s))
        return acc

    def batch_mul(self, points, scalars):
        """batch_multi_scalar_g{1,2}: curve.rs:326-354 (zip -> shortest length)."""
        return [self.mul(p, s) for p, s in zip(points, scalars)]

    # ---- serialisation (ark-serialize compressed) ----
    def to_bytes(self, p):
        """BN254: ark SW default -- LE x (Fq2 = c0||c1), flags in the top two bits of the LAST byte
        (bit7 = y is the lexicographically larger root, bit6 = infinity).
        BLS12-381: ark-bls12-381 0.4.0 uses the Zcash/IETF format -- BE x (G2: c1||c0), flags in the top
        three bits of byte 0 (bit7 compressed, bit6 infinity, bit5 y larger)."""
        nb = self.P.fq_bytes
        ncoord = 2 if self.is_g2 else 1
        if self.curve == BN254:
            if p is None:
                out = bytearray(nb * ncoord)
                out[-1] |= 0x40
                return bytes(out)
            xs = p[0] if self.is_g2 else (p[0],)
            out = bytearray(b"".join(c.to_bytes(nb, "little") for c in xs))
            if self._y_is_larger(p[1]):
                out[-1] |= 0x80
            return bytes(out)
        if p is None:
            out = bytearray(nb * ncoord)
            out[0] |= 0xC0
            return bytes(out)
        xs = (p[0][1], p[0][0]) if self.is_g2 else (p[0],)
        out = bytearray(b"".join(c.to_bytes(nb, "big") for c in xs))
        out[0] |= 0x80
        if self._y_is_larger(p[1]):
            out[0] |= 0x20
        return bytes(out)

    def _y_is_larger(self, y):
        q = self.P.q
        if not self.is_g2:
            return y > (q - y) % q
        ny = ((-y[0]) % q, (-y[1]) % q)
        return (y[1], y[0]) > (ny[1], ny[0])  # compare c1 first, then c0

    def from_bytes(self, b):
        b = bytes(b)
        nb = self.P.fq_bytes
        ncoord = 2 if self.is_g2 else 1
        if len(b) != nb * ncoord:
            raise ValueError("Cannot deserialize point: bad length")
        F = self.F
        if self.curve == BN254:
            flags = b[-1] & 0xC0
            raw = bytearray(b)
            raw[-1] &= 0x3F
            cs = [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(ncoord)]
            inf, larger = bool(flags & 0x40), bool(flags & 0x80)
            if inf and larger:
                raise ValueError("Cannot deserialize point: invalid flags")
        else:
            flags = b[0] & 0xE0
            if not flags & 0x80:
                raise ValueError("Cannot deserialize point: uncompressed encoding")
            raw = bytearray(b)
            raw[0] &= 0x1F
            cs = [int.from_bytes(raw[i * nb:(i + 1) * nb], "big") for i in range(ncoord)]
            cs.reverse()
            inf, larger = bool(flags & 0x40), bool(flags & 0x20)
        if any(c >= self.P.q for c in cs):
            raise ValueError("Cannot deserialize point: coordinate not in field")
        if inf:
            if any(cs):
                raise ValueError("Cannot deserialize point: non-zero infinity")
            return None
        x = tuple(cs) if self.is_g2 else cs[0]
        y = F.sqrt(F.add(F.mul(F.mul(x, x), x), self.b))
        if y is None:
            raise ValueError("Cannot deserialize point: not on curve")
        if self._y_is_larger(y) != larger:
            y = F.neg(y)
        p = (x, y)
        if self.mul_raw(p, self.r) is not None:
            raise ValueError("Cannot deserialize point: not in the prime-order subgroup")
        return p

    def mul_raw(self, p, k):
        """Scalar mul without reducing k (subgroup check needs k = r)."""
        F = self.F
        acc = (F.one, F.one, F.zero)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == "1":
                acc = self._jadd_affine(acc, p)
        return self._to_affine(acc)


_groups = {}


def group(curve, g2=False):
    key = (curve, bool(g2))
    if key not in _groups:
        _groups[key] = Group(curve, g2)
    return _groups[key]
