"""`EllipticCurve`: curve-name front door to the `_algebra` ec modules, with the method names the reference's
python/zksnake/ecc.py:47-142 offers (G1, G2, pairing, multi_pairing, batch_mul, multiexp, from_hex) so that prover code reads
the same.  Point vectors may be Python lists (the reference's representation) or device-resident `PointVector`s -- the form the
B200 provers keep keys and SRS in; multiexp / batch_mul then never build a Python list of points.
"""
from ._algebra import ec_bls12_381, ec_bn254
from ._algebra._ec import PointVector

# name -> (module, compressed G1 size in bytes)
_REGISTRY = {"BN128": (ec_bn254, 32), "BN254": (ec_bn254, 32), "ALT_BN128": (ec_bn254, 32), "BLS12_381": (ec_bls12_381, 48)}
CURVE_MODULES = {k: v[0] for k, v in _REGISTRY.items()}
POINT_SIZE = {k: v[1] for k, v in _REGISTRY.items()}     # CurvePointSize, ecc.py:40-44


def point_group(x):
    """1 for a G1 point, 2 for a G2 point (either curve), 0 for anything else."""
    for mod in (ec_bn254, ec_bls12_381):
        if isinstance(x, mod.PointG1):
            return 1
        if isinstance(x, mod.PointG2):
            return 2
    return 0


def ispointG1(x):
    return point_group(x) == 1


def ispointG2(x):
    return point_group(x) == 2


class EllipticCurve:
    def __init__(self, curve="BN254"):
        if curve not in _REGISTRY:
            raise ValueError(f"Unsupported curve: {curve}")
        self.name = curve
        self.curve, self._g1_bytes = _REGISTRY[curve]
        self.order, self.field_modulus = self.curve.ORDER, self.curve.FIELD_MODULUS
        self._classes = {1: self.curve.PointG1, 2: self.curve.PointG2}
        self._msm = {1: self.curve.multiscalar_mul_g1, 2: self.curve.multiscalar_mul_g2}
        self._batch = {1: self.curve.batch_multi_scalar_g1, 2: self.curve.batch_multi_scalar_g2}

    def _group(self, x, what):
        for grp, cls in self._classes.items():
            if isinstance(x, cls):
                return grp
        raise TypeError(f"Invalid curve type: {what}")

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
        """[s_i * g_i] (g a list) or [s_i * g] (g one point) as a list of points -- ecc.py:88-105."""
        bases = g if isinstance(g, list) else [g] * len(s)
        if not bases:
            return []
        return self._batch[self._group(bases[0], bases[0])](bases, s)

    def batch_mul_device(self, g, s, group=1):
        """The same products left in HBM as a PointVector: how setup() builds a 2^20-point SRS without 2^20 Python objects."""
        return self.curve.batch_mul_device(g, s, group)

    def multiexp(self, g, s):
        """sum_i s_i * g_i -- ecc.py:107-126: no scalars -> the identity; surplus points are ignored."""
        assert len(g) > 0
        on_device = isinstance(g, PointVector)
        grp = g.group if on_device else self._group(g[0], type(g[0]))
        if len(s) == 0:
            return self._classes[grp].identity()
        if on_device:
            if len(s) > len(g):
                raise ValueError("Number of points and scalars mismatch")
            return self._msm[grp](g.prefix(len(s)), s)
        return self._msm[grp](g[:len(s)] if len(s) < len(g) else g, s)

    def from_hex(self, hexstring):
        """compressed point from hex; the length tells G1 from G2 (ecc.py:128-142)"""
        raw = bytes.fromhex(hexstring)
        if len(raw) == self._g1_bytes:
            return self.curve.PointG1.from_bytes(raw)
        if len(raw) == 2 * self._g1_bytes:
            return self.curve.PointG2.from_bytes(raw)
        raise ValueError(f"Hexstring size of {2 * self._g1_bytes} or {4 * self._g1_bytes} expected, got {len(hexstring)}")

    def __call__(self, x, y):
        if isinstance(x, (tuple, list)) and isinstance(y, (tuple, list)):
            return self.curve.PointG2(x[0], x[1], y[0], y[1])
        return self.curve.PointG1(x, y)
