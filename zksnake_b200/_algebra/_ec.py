"""Builder of the `ec_bn254` / `ec_bls12_381` modules: the Python-visible surface the reference registers in
/root/reference/src/lib.rs:6-68 (PointG1, PointG2, g1, g2, batch_multi_scalar_g1/g2, multiscalar_mul_g1/g2, pairing,
multi_pairing), re-implemented over libzkb200.so.

Points are host objects holding canonical affine coordinates (None = identity); the group operators go through the
host-side group code of the library (zkb_point_lincomb), vectors of points go to the GPU (zkb_msm, zkb_batch_mul_dev).
A `PointVector` keeps a vector of points resident in HBM (Montgomery form) so that a proving key is uploaded once;
multiscalar_mul_g1/g2 accept either a list of points (reference signature) or a PointVector.
"""
import ctypes
import hashlib

import numpy as np

from .. import _native as nat
from . import _pairing

_Q = {
    0: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    1: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
}
_R = {
    0: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    1: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
}
_G1 = {
    0: (1, 2),
    1: (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
        0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1),
}
_G2 = {
    0: ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    1: ((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
         0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
        (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
         0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)),
}
_B1 = {0: 3, 1: 4}


def _fq2_mul(a, b, q):
    return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


def _fq2_inv(a, q):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, q)
    return (a[0] * d % q, (-a[1]) * d % q)


def _b2(curve):
    q = _Q[curve]
    return _fq2_mul((3, 0), _fq2_inv((9, 1), q), q) if curve == 0 else (4, 4)


def _fq_sqrt(a, q):
    s = pow(a, (q + 1) // 4, q)  # q = 3 mod 4 on both curves
    return s if s * s % q == a % q else None


def _fq2_sqrt(a, q):
    a0, a1 = a[0] % q, a[1] % q
    if a0 == 0 and a1 == 0:
        return (0, 0)
    if a1 == 0:
        s = _fq_sqrt(a0, q)
        if s is not None:
            return (s, 0)
        s = _fq_sqrt((-a0) % q, q)
        return (0, s) if s is not None else None
    n = _fq_sqrt((a0 * a0 + a1 * a1) % q, q)
    if n is None:
        return None
    inv2 = pow(2, -1, q)
    for cand in ((a0 + n) * inv2 % q, (a0 - n) * inv2 % q):
        x0 = _fq_sqrt(cand, q)
        if x0:
            x1 = a1 * pow(2 * x0, -1, q) % q
            if _fq2_mul((x0, x1), (x0, x1), q) == (a0, a1):
                return (x0, x1)
    return None


class PointVector:
    """A vector of G1 or G2 points resident on the device (affine, Montgomery form)."""

    def __init__(self, curve, group, n, buf=None):
        self.curve, self.group, self.n = curve, group, n
        self.affine_bytes = nat.lib.zkb_affine_bytes(curve, group)
        self.buf = buf if buf is not None else nat.DeviceBuffer(max(n, 1) * self.affine_bytes)

    @property
    def ptr(self):
        return self.buf.ptr

    def __len__(self):
        return self.n

    def prefix(self, n):
        """The first n points as a view sharing this vector's device buffer (no copy)."""
        assert 0 <= n <= self.n
        return PointVector(self.curve, self.group, n, buf=self.buf)

    def to_bytes(self):
        """The n ark-serialize compressed encodings, concatenated (what the reference's key serialisers produce by looping
        PointG1/G2.to_bytes: groth16/serialization.py:131-159, plonk/serialization.py:255-262) -- one kernel launch."""
        size = nat.lib.zkb_compressed_bytes(self.curve, self.group)
        out = np.zeros(self.n * size, dtype=np.uint8)
        nat.check(nat.lib.zkb_points_compress(self.curve, self.group, self.ptr, self.n, nat.ptr(out)))
        return out.tobytes()

    @classmethod
    def from_bytes(cls, curve, group, data, validate=True):
        """Inverse of to_bytes with the checks of from_bytes per point (flags, x < q, on the curve, in the subgroup); raises
        ValueError("Cannot deserialize point: ...") naming the first offending index (ecc.py:128-142).  validate: True / 1 =
        subgroup membership by the endomorphism criteria in G2 (r * P in G1), 2 = r * P = infinity everywhere (the definition;
        same accept set, 3x slower in G2), False / 0 = no subgroup check."""
        nat.ensure_init()
        size = nat.lib.zkb_compressed_bytes(curve, group)
        if len(data) % size:
            raise ValueError("Cannot deserialize point: bad length")
        n = len(data) // size
        vec = cls(curve, group, n)
        if n:
            raw = np.frombuffer(bytes(data), dtype=np.uint8)
            nat.check(nat.lib.zkb_points_decompress(curve, group, raw.ctypes.data, n, int(validate), vec.ptr, None, None))
        return vec

    def download(self):
        """-> uint64 array (n, affine_bytes/8), canonical coordinates (all-zero row = identity)."""
        out = np.zeros((self.n, self.affine_bytes // 8), dtype=np.uint64)
        nat.check(nat.lib.zkb_points_download(self.curve, self.group, self.ptr, self.n, nat.ptr(out)))
        return out


def build(curve):
    """Return the namespace dict of the ec module for `curve` (0 = BN254, 1 = BLS12-381)."""
    q, r = _Q[curve], _R[curve]
    nb = 32 if curve == 0 else 48  # Fq bytes on the wire (little-endian limbs)

    def coords_bytes(coords):
        return b"".join(c.to_bytes(nb, "little") for c in coords)

    class _Point:
        __slots__ = ("_p",)
        _group = 1

        # ---- helpers ----
        @classmethod
        def _wrap(cls, p):
            o = cls.__new__(cls)
            o._p = p
            return o

        def _flat(self):
            raise NotImplementedError

        @classmethod
        def _from_flat(cls, arr, inf):
            raise NotImplementedError

        @classmethod
        def _lincomb(cls, pts, scalars):
            """sum k_i P_i (k_i None -> 1) through the library's host-side group code."""
            n = len(pts)
            pbuf = np.frombuffer(b"".join(p._flat() for p in pts), dtype=np.uint64).copy()
            infs = np.array([1 if p._p is None else 0 for p in pts], dtype=np.int32)
            sbuf = np.frombuffer(b"".join(((s or 0) % r).to_bytes(32, "little") for s in scalars), dtype=np.uint64).copy()
            has = np.array([0 if s is None else 1 for s in scalars], dtype=np.int32)
            out = np.zeros(nat.lib.zkb_affine_bytes(curve, cls._group) // 8, dtype=np.uint64)
            inf = ctypes.c_int(0)
            nat.check(nat.lib.zkb_point_lincomb(curve, cls._group, n, nat.ptr(pbuf), nat.ptr(infs), nat.ptr(sbuf), nat.ptr(has),
                                                nat.ptr(out), ctypes.byref(inf)))
            return cls._from_flat(out, inf.value)

        # ---- reference API (curve.rs:58-118, :233-292) ----
        @property
        def generator(self):
            return type(self)._generator()

        def __add__(self, other):
            return self._lincomb([self, other], [None, None])

        __radd__ = __add__

        def __sub__(self, other):
            return self._lincomb([self, -other], [None, None])

        def __rsub__(self, other):
            return self.__sub__(other)  # the reference's __rsub__ also computes self - other (curve.rs:92-94)

        def __mul__(self, k):
            return self._lincomb([self], [int(k) % r])  # Fr::from(BigUint) reduces mod r

        __rmul__ = __mul__

        def __eq__(self, other):
            return isinstance(other, type(self)) and self._p == other._p

        def __hash__(self):
            return hash((type(self).__name__, self._p))

        def is_zero(self):
            return self._p is None

        def to_hex(self):
            return bytes(self.to_bytes()).hex()

        def __repr__(self):
            return self.__str__()

        @classmethod
        def identity(cls):
            return cls._wrap(None)

    class PointG1(_Point):
        __slots__ = ()
        _group = 1

        def __init__(self, x, y):  # curve.rs:28-33 (affine constructor, unchecked like G1Affine::new_unchecked is not: ark checks)
            x, y = int(x) % q, int(y) % q
            if (y * y - x * x * x - _B1[curve]) % q != 0:
                raise ValueError("point is not on the curve")
            self._p = (x, y)

        @classmethod
        def _generator(cls):
            return cls._wrap(_G1[curve])

        @property
        def x(self):
            return 0 if self._p is None else self._p[0]

        @property
        def y(self):
            return 0 if self._p is None else self._p[1]

        def _flat(self):
            return b"\0" * (2 * nb) if self._p is None else coords_bytes(self._p)

        @classmethod
        def _from_flat(cls, arr, inf):
            if inf:
                return cls._wrap(None)
            raw = np.ascontiguousarray(arr).tobytes()
            return cls._wrap((int.from_bytes(raw[:nb], "little"), int.from_bytes(raw[nb:2 * nb], "little")))

        def __neg__(self):
            return self._wrap(None if self._p is None else (self._p[0], (-self._p[1]) % q))

        def __str__(self):
            return "infinity" if self._p is None else f"({self._p[0]}, {self._p[1]})"

        def to_bytes(self):
            """ark-serialize compressed form as a list of byte values (curve.rs:127-132)."""
            p = self._p
            if curve == 0:
                if p is None:
                    out = bytearray(32)
                    out[-1] |= 0x40
                else:
                    out = bytearray(p[0].to_bytes(32, "little"))
                    if p[1] > (q - p[1]) % q:
                        out[-1] |= 0x80
            else:
                if p is None:
                    out = bytearray(48)
                    out[0] |= 0xC0
                else:
                    out = bytearray(p[0].to_bytes(48, "big"))
                    out[0] |= 0x80
                    if p[1] > (q - p[1]) % q:
                        out[0] |= 0x20
            return list(out)

        @classmethod
        def from_bytes(cls, data):
            b = bytes(data)
            try:
                if len(b) != (32 if curve == 0 else 48):
                    raise ValueError("bad length")
                if curve == 0:
                    inf, larger = bool(b[-1] & 0x40), bool(b[-1] & 0x80)
                    x = int.from_bytes(b[:-1] + bytes([b[-1] & 0x3F]), "little")
                    if inf and larger:
                        raise ValueError("invalid flags")
                else:
                    if not b[0] & 0x80:
                        raise ValueError("uncompressed encoding")
                    inf, larger = bool(b[0] & 0x40), bool(b[0] & 0x20)
                    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
                if x >= q:
                    raise ValueError("coordinate not in field")
                if inf:
                    if x:
                        raise ValueError("non-zero infinity")
                    return cls._wrap(None)
                y = _fq_sqrt((x * x * x + _B1[curve]) % q, q)
                if y is None:
                    raise ValueError("not on curve")
                if (y > (q - y) % q) != larger:
                    y = (q - y) % q
                pt = cls._wrap((x, y))
                if curve == 1 and not cls._lincomb([pt], [r - 1])._p == (-pt)._p:
                    raise ValueError("not in the prime-order subgroup")
                return pt
            except ValueError as exc:
                raise ValueError(f"Cannot deserialize point: {exc}") from None

        @classmethod
        def hash_to_field(cls, dst, data):
            raise NotImplementedError("hash_to_field is outside the proving hot path (SURVEY.md section 2.1 N4)")

        @classmethod
        def hash_to_curve(cls, dst, data):
            raise NotImplementedError("hash_to_curve is outside the proving hot path (SURVEY.md section 2.1 N4)")

        @classmethod
        def from_x(cls, x):
            x = int(x) % q
            y = _fq_sqrt((x * x * x + _B1[curve]) % q, q)
            if y is None:
                raise ValueError("Cannot found point")
            if y < (q - y) % q:  # ark's get_point_from_x_unchecked(x, greatest=true)
                y = (q - y) % q
            return cls._wrap((x, y))

    class PointG2(_Point):
        __slots__ = ()
        _group = 2

        def __init__(self, x1, x2, y1, y2):  # curve.rs:197-204: (x.c0, x.c1, y.c0, y.c1)
            x = (int(x1) % q, int(x2) % q)
            y = (int(y1) % q, int(y2) % q)
            lhs = _fq2_mul(y, y, q)
            x3 = _fq2_mul(_fq2_mul(x, x, q), x, q)
            b = _b2(curve)
            if lhs != ((x3[0] + b[0]) % q, (x3[1] + b[1]) % q):
                raise ValueError("point is not on the curve")
            self._p = (x, y)

        @classmethod
        def _generator(cls):
            return cls._wrap(_G2[curve])

        @property
        def x(self):
            return [0, 0] if self._p is None else list(self._p[0])

        @property
        def y(self):
            return [0, 0] if self._p is None else list(self._p[1])

        def _flat(self):
            if self._p is None:
                return b"\0" * (4 * nb)
            return coords_bytes((self._p[0][0], self._p[0][1], self._p[1][0], self._p[1][1]))

        @classmethod
        def _from_flat(cls, arr, inf):
            if inf:
                return cls._wrap(None)
            raw = np.ascontiguousarray(arr).tobytes()
            c = [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(4)]
            return cls._wrap(((c[0], c[1]), (c[2], c[3])))

        def __neg__(self):
            if self._p is None:
                return self._wrap(None)
            (x, y) = self._p
            return self._wrap((x, ((-y[0]) % q, (-y[1]) % q)))

        def __str__(self):
            return f"({self.x}, {self.y})"

        @staticmethod
        def _y_larger(y):
            ny = ((-y[0]) % q, (-y[1]) % q)
            return (y[1], y[0]) > (ny[1], ny[0])

        def to_bytes(self):
            p = self._p
            if curve == 0:
                if p is None:
                    out = bytearray(64)
                    out[-1] |= 0x40
                else:
                    out = bytearray(p[0][0].to_bytes(32, "little") + p[0][1].to_bytes(32, "little"))
                    if self._y_larger(p[1]):
                        out[-1] |= 0x80
            else:
                if p is None:
                    out = bytearray(96)
                    out[0] |= 0xC0
                else:
                    out = bytearray(p[0][1].to_bytes(48, "big") + p[0][0].to_bytes(48, "big"))
                    out[0] |= 0x80
                    if self._y_larger(p[1]):
                        out[0] |= 0x20
            return list(out)

        @classmethod
        def from_bytes(cls, data):
            b = bytes(data)
            try:
                if len(b) != (64 if curve == 0 else 96):
                    raise ValueError("bad length")
                if curve == 0:
                    inf, larger = bool(b[-1] & 0x40), bool(b[-1] & 0x80)
                    raw = b[:-1] + bytes([b[-1] & 0x3F])
                    x = (int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:], "little"))
                    if inf and larger:
                        raise ValueError("invalid flags")
                else:
                    if not b[0] & 0x80:
                        raise ValueError("uncompressed encoding")
                    inf, larger = bool(b[0] & 0x40), bool(b[0] & 0x20)
                    raw = bytes([b[0] & 0x1F]) + b[1:]
                    x = (int.from_bytes(raw[48:], "big"), int.from_bytes(raw[:48], "big"))
                if x[0] >= q or x[1] >= q:
                    raise ValueError("coordinate not in field")
                if inf:
                    if x != (0, 0):
                        raise ValueError("non-zero infinity")
                    return cls._wrap(None)
                x3 = _fq2_mul(_fq2_mul(x, x, q), x, q)
                bb = _b2(curve)
                y = _fq2_sqrt(((x3[0] + bb[0]) % q, (x3[1] + bb[1]) % q), q)
                if y is None:
                    raise ValueError("not on curve")
                if cls._y_larger(y) != larger:
                    y = ((-y[0]) % q, (-y[1]) % q)
                pt = cls._wrap((x, y))
                if not cls._lincomb([pt], [r - 1])._p == (-pt)._p:
                    raise ValueError("not in the prime-order subgroup")
                return pt
            except ValueError as exc:
                raise ValueError(f"Cannot deserialize point: {exc}") from None

    class PointG12:
        """Pairing output (curve.rs:394-415): supports == and printing."""

        def __init__(self, value):
            self.value = value

        def __eq__(self, other):
            return isinstance(other, PointG12) and self.value == other.value

        def __str__(self):
            return hashlib.sha256(repr(self.value).encode()).hexdigest()

        __repr__ = __str__

    # ---- vectors of points ------------------------------------------------------------------------------------
    def _pack_points(points, cls):
        return np.frombuffer(b"".join(p._flat() for p in points), dtype=np.uint64).copy()

    def _pack_scalars(scalars):
        # arbitrary non-negative ints; the library reduces mod r (Fr::from(BigUint)); clamp to 256 bits first
        return nat.ints_to_limbs([int(s) if 0 <= int(s) < (1 << 256) else int(s) % r for s in scalars], 32)

    def upload_points(points, group):
        """list[PointG1|PointG2] -> PointVector (device-resident, Montgomery form)."""
        nat.ensure_init()
        cls = PointG1 if group == 1 else PointG2
        pv = PointVector(curve, group, len(points))
        if points:
            arr = _pack_points(points, cls)
            nat.check(nat.lib.zkb_points_upload(curve, group, nat.ptr(arr), len(points), pv.ptr))
        return pv

    def _msm(points, scalars, group, cls):
        nat.ensure_init()
        if len(points) != len(scalars):
            raise ValueError("Number of points and scalars mismatch")
        n = len(scalars)
        out = np.zeros(nat.lib.zkb_affine_bytes(curve, group) // 8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        s = _pack_scalars(scalars) if n else np.zeros((1, 4), dtype=np.uint64)
        if isinstance(points, PointVector):
            d_s = nat.DeviceBuffer(max(n, 1) * 32).upload(s)
            nat.check(nat.lib.zkb_fr_reduce_dev(curve, n, d_s.ptr))
            nat.check(nat.lib.zkb_msm_dev(curve, group, points.ptr, d_s.ptr, n, nat.ptr(out), ctypes.byref(inf)))
            d_s.free()
        else:
            p = _pack_points(points, cls) if n else np.zeros(1, dtype=np.uint64)
            nat.check(nat.lib.zkb_msm(curve, group, nat.ptr(p), n, nat.ptr(s), n, nat.ptr(out), ctypes.byref(inf)))
        return cls._from_flat(out, inf.value)

    def multiscalar_mul_g1(points, scalars):
        """curve.rs:356-373."""
        return _msm(points, scalars, 1, PointG1)

    def multiscalar_mul_g2(points, scalars):
        """curve.rs:375-392."""
        return _msm(points, scalars, 2, PointG2)

    def batch_mul_device(points, scalars, group):
        """scalars[i] * points[i] (or * points if a single point) -> PointVector; the GPU fixed-base path of setup."""
        nat.ensure_init()
        cls = PointG1 if group == 1 else PointG2
        single = isinstance(points, cls)
        n = len(scalars)
        if single:
            bases = upload_points([points], group)
        elif isinstance(points, PointVector):
            bases = points
        else:
            n = min(n, len(points))  # zip semantics of curve.rs:331
            bases = upload_points(points[:n], group)
        out = PointVector(curve, group, n)
        if n:
            d_s = nat.DeviceBuffer(n * 32).upload(_pack_scalars(scalars[:n]))
            nat.check(nat.lib.zkb_fr_reduce_dev(curve, n, d_s.ptr))
            nat.check(nat.lib.zkb_batch_mul_dev(curve, group, bases.ptr, 1 if single else 0, d_s.ptr, n, out.ptr))
            nat.check(nat.lib.zkb_sync())
            d_s.free()
        return out

    def _unpack_vector(pv, cls):
        arr = pv.download()
        res = []
        for row in arr:
            res.append(cls._from_flat(row, 0 if row.any() else 1))
        return res

    def batch_multi_scalar_g1(points, scalars):
        """curve.rs:326-339."""
        return _unpack_vector(batch_mul_device(points, scalars, 1), PointG1)

    def batch_multi_scalar_g2(points, scalars):
        """curve.rs:341-354."""
        return _unpack_vector(batch_mul_device(points, scalars, 2), PointG2)

    def pairing(a, b):
        """curve.rs:417-422 -- CPU support code for verify(); not on the proving path."""
        return PointG12(_pairing.pairing(curve, a._p, b._p))

    def multi_pairing(a_list, b_list):
        """curve.rs:424-437."""
        return PointG12(_pairing.multi_pairing(curve, [a._p for a in a_list], [b._p for b in b_list]))

    def g1():
        return PointG1._generator()

    def g2():
        return PointG2._generator()

    return {
        "PointG1": PointG1, "PointG2": PointG2, "PointG12": PointG12, "PointVector": PointVector,
        "g1": g1, "g2": g2,
        "batch_multi_scalar_g1": batch_multi_scalar_g1, "batch_multi_scalar_g2": batch_multi_scalar_g2,
        "multiscalar_mul_g1": multiscalar_mul_g1, "multiscalar_mul_g2": multiscalar_mul_g2,
        "pairing": pairing, "multi_pairing": multi_pairing,
        "upload_points": upload_points, "batch_mul_device": batch_mul_device,
        "CURVE_ID": curve, "ORDER": r, "FIELD_MODULUS": q,
    }
