"""Bulk point codec at key-file sizes, for `ncu --metrics gpu__time_duration.sum -k regex:compress` (kernel durations) and for
wall-clock timing: one to_bytes + from_bytes per (curve, group)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200.ecc import EllipticCurve  # noqa: E402
from zksnake_b200.frvec import FrVec  # noqa: E402
from zksnake_b200 import _native as nat  # noqa: E402

nat.ensure_init()
for curve, grp, log_n in (("BN254", 1, 20), ("BN254", 2, 18), ("BLS12_381", 1, 18), ("BLS12_381", 2, 16)):
    E = EllipticCurve(curve)
    cid = E.curve.CURVE_ID
    n = 1 << log_n
    gen = E.curve.upload_points([E.G1() if grp == 1 else E.G2()], grp)
    ks = FrVec.powers(cid, n, 3, 7)
    vec = E.curve.PointVector(cid, grp, n)
    nat.check(nat.lib.zkb_batch_mul_dev(cid, grp, gen.ptr, 1, ks.ptr, n, vec.ptr))
    nat.check(nat.lib.zkb_sync())
    raw = vec.to_bytes()
    back = E.curve.PointVector.from_bytes(cid, grp, raw)
    t0 = time.perf_counter()
    raw = vec.to_bytes()
    t1 = time.perf_counter()
    back = E.curve.PointVector.from_bytes(cid, grp, raw)
    t2 = time.perf_counter()
    assert back.to_bytes() == raw
    print(f"codec {curve} G{grp} n=2^{log_n}: to_bytes {1e3 * (t1 - t0):.2f} ms, from_bytes (validated) {1e3 * (t2 - t1):.2f} ms", flush=True)
