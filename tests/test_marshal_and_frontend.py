"""No-GPU checks of the host side of the binding layer: the C-level list[int] <-> limb marshaller (csrc/pymarshal.cpp) against
Python's own int.to_bytes / from_bytes, and the circuit front-end stand-in (`_algebra/circuit.py`) through the REFERENCE's
unmodified R1CS / Plonkish classes and its own test_symbolic.py (when the reference Python package is available: the build
container, or baseline/_ref on the GPU box)."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q381 = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def test_ints_to_limbs_matches_to_bytes(native):
    rnd = random.Random(3)
    vals = [rnd.randrange(R) for _ in range(70000)] + [0, 1, (1 << 64) - 1, 1 << 64, (1 << 256) - 1, 1 << 30, (1 << 60) + 3, True]
    arr = native.ints_to_limbs(vals)
    assert arr.shape == (len(vals), 4) and arr.tobytes() == b"".join(int(v).to_bytes(32, "little") for v in vals)
    assert native.limbs_to_ints(arr) == [int(v) for v in vals]
    # 6-limb base-field elements
    v6 = [rnd.randrange(Q381) for _ in range(1000)] + [0, Q381 - 1, (1 << 384) - 1]
    a6 = native.ints_to_limbs(v6, 48)
    assert a6.tobytes() == b"".join(v.to_bytes(48, "little") for v in v6) and native.limbs_to_ints(a6, 48) == v6
    assert native.ints_to_limbs([]).shape == (0, 4) and native.limbs_to_ints(np.zeros((0, 4), np.uint64)) == []


def test_ints_to_limbs_reduction_and_errors(native):
    vals = [-1, -R, R + 5, (1 << 300) + 7, 5]
    arr = native.ints_to_limbs(vals, 32, modulus=R)
    assert native.limbs_to_ints(arr) == [R - 1, 0, R + 5, ((1 << 300) + 7) % R, 5]     # in-range values pass through raw
    with pytest.raises(OverflowError):
        native.ints_to_limbs([1, -1])
    with pytest.raises(OverflowError):
        native.ints_to_limbs([1 << 256])
    with pytest.raises(OverflowError, match="negative"):
        native.ints_to_limbs([1 << 300, -1], 32, modulus=R, allow_negative=False)
    with pytest.raises(TypeError):
        native.ints_to_limbs([1, "a"])
    with pytest.raises(TypeError):
        native.ints_to_limbs([1.5])
    assert native.ints_to_limbs([np.int64(7), np.uint32(9)])[:, 0].tolist() == [7, 9]
    # (coeff, terms) tuples of the reference's Polynomial factory
    terms = [(c, [(0, 0)]) for c in (3, 0, R + 1)]
    assert native.limbs_to_ints(native.ints_to_limbs(terms, 32, modulus=R, item=0)) == [3, 0, R + 1]
    # into a caller-provided buffer slice
    buf = np.zeros((6, 4), dtype=np.uint64)
    native.ints_to_limbs([11, 12], out=buf[1:3])
    assert buf[:, 0].tolist() == [0, 11, 12, 0, 0, 0]


def _reference_dir():
    sys.path.insert(0, ROOT)
    from zksnake_b200 import dropin
    return dropin.reference_python_dir()


@pytest.mark.skipif(_reference_dir() is None, reason="reference Python package not present (no /root/reference, no baseline/_ref)")
def test_circuit_stand_in_through_the_reference_classes():
    """README circuit (SURVEY.md section 8c KATs: w = [1, 35, 3, 9], A.w = [3, 9], B.w = [3, 3], C.w = [9, 27]) and the reference's
    PlonK test circuit, built with the stand-in front end and lowered / checked by the reference's own R1CS and Plonkish."""
    code = r"""
import zksnake_b200.dropin as d
assert d.install() and d.installed()
from zksnake.arithmetization import Var, ConstraintSystem, R1CS
from zksnake.arithmetization.plonkish import Plonkish
from zksnake.constant import BN254_SCALAR_FIELD as P
x, y, v1 = Var("x"), Var("y"), Var("v1")
cs = ConstraintSystem(["x"], ["y"], P)
cs.add_constraint(v1 == x * x); cs.add_constraint(y - 5 - x == v1 * x); cs.set_public(y)
r = R1CS(cs); r.compile()
pub, priv = r.generate_witness(r.solve({"x": 3}))
assert (pub, priv) == ([1, 35], [3, 9]) and r.is_sat(pub, priv) and not r.is_sat(pub, [3, 10])
w = pub + priv
assert (r.A.dot(w), r.B.dot(w), r.C.dot(w)) == ([3, 9], [3, 3], [9, 27])
z, v0, v2, v3, v4, v5, v6 = (Var(n) for n in ("z", "v0", "v2", "v3", "v4", "v5", "v6"))
cs = ConstraintSystem(["x"], ["y"], P)
for eq in (z == x, v0 == z * z, v1 == z * z, v2 == v1 * x, v3 == v0 * 2 * 3, v4 == 2 * v1 * v2 * 3, v5 == 2 * v3 - v4,
           v6 == 2 + v5 + 3, y == v6 + v4 + 1337):
    cs.add_constraint(eq)
cs.set_public(y); cs.set_public(z)
pk = Plonkish(cs); pk.compile()
pub, priv = pk.generate_witness(pk.solve({"x": 3}))
assert pk.is_sat(pub, priv)
priv[4] += 1
assert not pk.is_sat(pub, priv)
print("ok")
"""
    res = subprocess.run([sys.executable, "-c", code], cwd="/tmp", env=dict(os.environ, PYTHONPATH=ROOT), stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout


@pytest.mark.skipif(_reference_dir() is None, reason="reference Python package not present")
def test_reference_test_symbolic_passes_over_the_stand_in():
    ref = _reference_dir()
    base = os.path.dirname(ref) if os.path.basename(ref) == "python" else ref
    res = subprocess.run([sys.executable, "-m", "pytest", "-p", "zksnake_b200.dropin", "-p", "no:cacheprovider", "-q", "--rootdir", base,
                          os.path.join(base, "tests", "test_symbolic.py")], cwd=base,
                         env=dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, ref])), stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "3 passed" in res.stdout, res.stdout
