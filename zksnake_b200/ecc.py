"""`EllipticCurve` facade over the `_algebra` ec modules -- the host-side counterpart of the reference's
python/zksnake/ecc.py:47-142 (G1/G2 generators, pairing, batch_mul, multiexp, from_hex), same method names and edge-case
behaviour; the arithmetic behind multiexp / batch_mul runs in libzkb200.so."""
from ._algebra import ec_bls12_381, ec_bn254
from ._algebra._ec import PointVector

CURVE_MODULES = {"BN128": ec_bn254, "BN254": ec_bn254, "ALT_BN128": ec_bn254, "BLS12_381": ec_bls12_381}
POINT_SIZE = {"BN128": 32, "BN254": 32, "ALT_BN128": 32, "BLS12_381": 48}   # CurvePointSize, ecc.py:40-44


def ispointG1(x):
    return isinstance(x, (ec_bn254.PointG1, ec_bls12_381.PointG1))


def ispointG2(x):
    return isinstance(x, (ec_bn254.PointG2, ec_bls12_381.PointG2))


class EllipticCurve:
    def __init__(self, curve="BN254"):
        if curve not in CURVE_MODULES:
            raise ValueError(f"Unsupported curve: {curve}")
        self.name = curve
        self.curve = CURVE_MODULES[curve]
        self.order = self.curve.ORDER
        self.field_modulus = self.curve.FIELD_MODULUS

    def G1(self):
        return self.curve.g1()

    def G2(self):
        return self.curve.g2()

    def pairing(self, a, b):
        return self.curve.pairing(a, b)

    def multi_pairing(self, a, b):
        assert len(a) == len(b), "Length of a and b must be equal"
        return self.curve.multi_pairing(a, b)

    def batch_mul(self, g, s):
        """ecc.py:88-105: s[i] * g[i], or s[i] * g for a single point."""
        if not isinstance(g, list):
            g = [g] * len(s)
        if len(g) == 0:
            return []
        if isinstance(g[0], self.curve.PointG1):
            return self.curve.batch_multi_scalar_g1(g, s)
        if isinstance(g[0], self.curve.PointG2):
            return self.curve.batch_multi_scalar_g2(g, s)
        raise TypeError(f"Invalid curve type: {g[0]}")

    def batch_mul_device(self, g, s, group=1):
        """The same, left on the GPU as a PointVector (what setup() uses so that a 2^20-point SRS never becomes a Python list)."""
        return self.curve.batch_mul_device(g, s, group)

    def multiexp(self, g, s):
        """ecc.py:107-126.  g: list of points or a device-resident PointVector; s: list of ints.  An empty scalar list gives
        the identity; more points than scalars are trimmed."""
        assert len(g) > 0
        if isinstance(g, PointVector):
            cls = self.curve.PointG1 if g.group == 1 else self.curve.PointG2
            if len(s) == 0:
                return cls.identity()
            if len(s) > len(g):
                raise ValueError("Number of points and scalars mismatch")
            fn = self.curve.multiscalar_mul_g1 if g.group == 1 else self.curve.multiscalar_mul_g2
            return fn(g.prefix(len(s)), s)
        if len(s) == 0:
            return g[0] * 0
        if len(s) < len(g):
            g = g[:len(s)]
        if isinstance(g[0], self.curve.PointG1):
            return self.curve.multiscalar_mul_g1(g, s)
        if isinstance(g[0], self.curve.PointG2):
            return self.curve.multiscalar_mul_g2(g, s)
        raise TypeError(f"Invalid curve type: {type(g[0])}")

    def from_hex(self, hexstring):
        b = bytes.fromhex(hexstring)
        n = POINT_SIZE[self.name] * 2
        if len(hexstring) == n:
            return self.curve.PointG1.from_bytes(b)
        if len(hexstring) == n * 2:
            return self.curve.PointG2.from_bytes(b)
        raise ValueError(f"Hexstring size of {n} or {n * 2} expected, got {len(hexstring)}")
