"""CPU-only checks of the host side of libzkb200.so against the oracle: the even/odd-accumulator Montgomery multiplier
(the same algorithm text the GPU runs, with the PTX carry flag emulated), and the XYZZ group law + 64-bit host field used
for MSM window recombination and proof assembly."""
import ctypes
import random

import numpy as np
import pytest

from oracle.curve import group
from oracle.fields import BN254, BLS12_381, PARAMS

FIELDS = [PARAMS[BN254].r, PARAMS[BN254].q, PARAMS[BLS12_381].r, PARAMS[BLS12_381].q]


def _pack(vals, nl):
    return np.frombuffer(b"".join(v.to_bytes(nl * 4, "little") for v in vals), dtype=np.uint32).copy()


@pytest.mark.parametrize("field", range(4))
def test_field_ops_host(native, field):
    p = FIELDS[field]
    nl = (p.bit_length() + 31) // 32
    rnd = random.Random(100 + field)
    n = 300
    A = [rnd.randrange(p) for _ in range(n)]
    B = [rnd.randrange(p) for _ in range(n)]
    A[:8] = [0, 1, p - 1, p - 1, 0, 2, 1, p - 2]
    B[:8] = [0, p - 1, p - 1, 1, 5, (p + 1) // 2, 1, p - 2]
    a, b = _pack(A, nl), _pack(B, nl)
    ops = [lambda x, y: x * y % p, lambda x, y: (x + y) % p, lambda x, y: (x - y) % p,
           lambda x, y: pow(x, p - 2, p), lambda x, y: (-x) % p,
           lambda x, y: (x * y + (x + y) * (x - y)) % p, lambda x, y: x * y % p]   # 5, 6: mont_dot2 (lazy Fp2 half product)
    for op, fn in enumerate(ops):
        out = np.zeros_like(a)
        rc = native.lib.zkb_test_field_op_host(field, op, n, native.ptr(a), native.ptr(b), native.ptr(out))
        assert rc == 0
        got = [int.from_bytes(out[i * nl:(i + 1) * nl].tobytes(), "little") for i in range(n)]
        assert got == [fn(x, y) for x, y in zip(A, B)], f"field {field} op {op}"


def _flat_point(G, pt):
    nb = G.P.fq_bytes
    nb8 = (nb + 7) // 8 * 8
    if pt is None:
        return b"\0" * (nb8 * (4 if G.is_g2 else 2))
    coords = (pt[0][0], pt[0][1], pt[1][0], pt[1][1]) if G.is_g2 else (pt[0], pt[1])
    return b"".join(c.to_bytes(nb8, "little") for c in coords)


def _unflat_point(G, raw, inf):
    if inf:
        return None
    nb8 = (G.P.fq_bytes + 7) // 8 * 8
    cs = [int.from_bytes(raw[i * nb8:(i + 1) * nb8], "little") for i in range(4 if G.is_g2 else 2)]
    return ((cs[0], cs[1]), (cs[2], cs[3])) if G.is_g2 else (cs[0], cs[1])


@pytest.mark.parametrize("curve", [BN254, BLS12_381])
@pytest.mark.parametrize("g2", [False, True])
def test_host_lincomb(native, curve, g2):
    G = group(curve, g2)
    rnd = random.Random(7 + curve * 2 + g2)
    gen = G.gen
    cases = []
    # random terms, identity operands, P + P, P - P, zero scalars, scalar r-1
    P1, P2 = G.mul(gen, rnd.randrange(G.r)), G.mul(gen, rnd.randrange(G.r))
    cases.append(([P1, P2, gen], [rnd.randrange(G.r), None, rnd.randrange(G.r)]))
    cases.append(([P1, P1], [None, None]))
    cases.append(([P1, G.neg(P1)], [None, None]))
    cases.append(([None, P2, None], [5, None, None]))
    cases.append(([P1, P2], [0, 0]))
    cases.append(([P1, P2], [G.r - 1, 1]))
    cases.append(([gen], [2]))
    for pts, scs in cases:
        n = len(pts)
        pbuf = np.frombuffer(b"".join(_flat_point(G, p) for p in pts), dtype=np.uint64).copy()
        infs = np.array([1 if p is None else 0 for p in pts], dtype=np.int32)
        sbuf = np.frombuffer(b"".join((s or 0).to_bytes(32, "little") for s in scs), dtype=np.uint64).copy()
        has = np.array([0 if s is None else 1 for s in scs], dtype=np.int32)
        out = np.zeros(len(_flat_point(G, None)) // 8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        rc = native.lib.zkb_point_lincomb(curve, 2 if g2 else 1, n, native.ptr(pbuf), native.ptr(infs), native.ptr(sbuf),
                                              native.ptr(has), native.ptr(out), ctypes.byref(inf))
        assert rc == 0
        exp = None
        for p, s in zip(pts, scs):
            exp = G.add(exp, p if s is None else G.mul(p, s))
        assert _unflat_point(G, out.tobytes(), inf.value) == exp


def test_library_exports_every_declared_symbol(native):
    """Every function declared in include/zkb200.h is exported by the built library (and bound in _native)."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"zkb_groth16_pk"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(native.lib, name), f"{name} declared in zkb200.h but not exported"
    assert declared == set(native.EXPORTED)


def test_compute_fails_loudly_without_gpu(native):
    if native.gpu_available():
        pytest.skip("GPU present")
    out = np.zeros(4, dtype=np.uint64)
    rc = native.lib.zkb_ntt(0, 0, 0, 0, native.ptr(out), 1, native.ptr(out))
    assert rc != 0
    with pytest.raises(Exception):
        native.ensure_init()


def test_circom_r1cs_file_round_trip(tmp_path):
    """zksnake_b200.r1cs.read_r1cs_file: the sections /root/reference/python/zksnake/parser.py:37-90 reads, into prover triplets."""
    from zksnake_b200 import r1cs as rm
    for curve in ("BN254", "BLS12_381"):
        for circuit in (rm.readme_circuit(curve), rm.chain_circuit(9, curve), rm.dense_random_circuit(6, curve)):
            c, pub, priv = circuit
            path = tmp_path / f"c_{curve}.r1cs"
            n_out = min(1, c.n_public - 1)
            rm.write_r1cs_file(path, c, n_pub_out=n_out)
            back, header = rm.read_r1cs_file(path)
            for x, y in ((c.A, back.A), (c.B, back.B), (c.C, back.C)):
                assert sorted(x.triplets) == sorted(y.triplets)
                assert y.n_col == x.n_col
            assert back.n_public == c.n_public and back.p == c.p
            assert header["n_wires"] == c.A.n_col and header["n_pub_out"] == n_out and header["n_pub_in"] == c.n_public - 1 - n_out
            assert header["fs"] == 32 and list(header["wire_labels"]) == list(range(c.A.n_col))
            assert back.is_sat(list(pub) + list(priv))
    raw = bytearray(path.read_bytes())
    bad = tmp_path / "bad.r1cs"
    bad.write_bytes(b"r2cs" + bytes(raw[4:]))
    import pytest
    with pytest.raises(AssertionError, match="Invalid magic bytes"):
        rm.read_r1cs_file(bad)
    bad.write_bytes(bytes(raw[:4]) + (2).to_bytes(4, "little") + bytes(raw[8:]))
    with pytest.raises(AssertionError, match="Unsupported r1cs file version: 2"):
        rm.read_r1cs_file(bad)
    bad.write_bytes(bytes(raw[:-3]))
    with pytest.raises(AssertionError):
        rm.read_r1cs_file(bad)
