"""KZG polynomial commitments on the GPU: commit / open / verify with the call surface of the reference's
python/zksnake/commitment/polynomial/kzg.py:10-58 (the multi-opening batching of :60-259 is outside the proving hot path).

B200 layout: the SRS [tau^i]G1 is produced by the fixed-base scalar-multiplication kernel and never leaves HBM; a fixed-base
MSM table is built over it once (zkb_msm_table_create); `Polynomial`s are device-resident coefficient vectors
(`_algebra/_poly.py`), so a commitment is ONE table MSM over scalars that are already on the device, and an opening is the
device division by (X - z) followed by the same MSM -- no coefficient ever becomes a Python int on the way.
"""
import ctypes
import random

import numpy as np

from . import _native as nat
from .ecc import EllipticCurve
from .frvec import FrVec
from .polynomial import Polynomial


def get_random_int(n_max):
    return random.SystemRandom().randint(1, n_max)


class KZG:
    def __init__(self, max_degree, group="BN254"):
        self.name = "KZG"
        self.degree, self.group = max_degree, group
        self.E = EllipticCurve(group)
        self.order = self.E.order
        self._cid = self.E.curve.CURVE_ID
        self.G1_tau = None        # PointVector, degree + 1 points
        self.G2_tau = None
        self._table = None
        self.is_setup = False

    def setup(self):
        tau = get_random_int(self.order)      # toxic waste: a local
        n = self.degree + 1
        powers = FrVec.powers(self._cid, n, tau % self.order)
        gen = self.E.curve.upload_points([self.E.G1()], 1)
        self.G1_tau = self.E.curve.PointVector(self._cid, 1, n)
        nat.check(nat.lib.zkb_batch_mul_dev(self._cid, 1, gen.ptr, 1, powers.ptr, n, self.G1_tau.ptr))
        self.G2_tau = self.E.G2() * tau
        self._drop_table()
        tab = ctypes.c_void_p()
        nat.check(nat.lib.zkb_msm_table_create(self._cid, 1, self.G1_tau.ptr, n, 0, 1, ctypes.byref(tab)))
        self._table = tab
        self.is_setup = True

    def _drop_table(self):
        if self._table:
            nat.lib.zkb_msm_table_free(self._table)
            self._table = None

    def __del__(self):
        try:
            self._drop_table()
        except Exception:
            pass

    def zero_commitment(self):
        return self.E.curve.PointG1.identity()

    def _msm(self, polynomial):
        """[P(tau)]G1 from the polynomial's device vector"""
        vec, n = polynomial.device_vector()
        P = self.E.curve.PointG1
        if n == 0:
            return P.identity()
        if n > len(self.G1_tau):
            raise ValueError("Number of points and scalars mismatch")
        xy = np.zeros(nat.lib.zkb_affine_bytes(self._cid, 1) // 8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        nat.check(nat.lib.zkb_msm_table_dev(self._table, vec.ptr, n, 0, 1, nat.ptr(xy), ctypes.byref(inf)))
        return P._from_flat(xy, inf.value)

    def commit(self, polynomial):
        assert self.is_setup, "Trusted setup has not been run"
        return self._msm(polynomial)

    def open(self, polynomial, point):
        """(proof, evaluation): the proof commits to (P - P(z)) / (X - z)"""
        assert self.is_setup, "Trusted setup has not been run"
        value = polynomial(point)
        quotient, remainder = (polynomial - value) / Polynomial([-point % self.order, 1], self.order)
        if not remainder.is_zero():
            raise ValueError("Given polynomial is not divided to zero")
        return self._msm(quotient), value

    def verify(self, commitment, proof, point, evaluation, transcript=None):
        """e(proof, [tau - z]G2) == e(C - [v]G1, G2) -- host pairing; not a proving-path operation"""
        assert self.is_setup, "Trusted setup has not been run"
        E = self.E
        return E.pairing(proof, self.G2_tau - E.G2() * point) == E.pairing(commitment - E.G1() * evaluation, E.G2())
