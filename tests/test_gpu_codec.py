"""GPU parity of the bulk point wire format: PointVector.to_bytes / from_bytes
(zkb_points_compress / zkb_points_decompress) against the oracle's per-point ark-serialize restatement (oracle/curve.py:197-271)
and the Groth16 key files of /root/reference/python/zksnake/groth16/serialization.py:68-220 (reference tests:
tests/test_groth16.py:186-210)."""
import random
import time

import numpy as np
import pytest

from oracle.curve import group
from oracle.fields import PARAMS, curve_id

pytestmark = pytest.mark.gpu

CASES = [("BN254", 1), ("BN254", 2), ("BLS12_381", 1), ("BLS12_381", 2)]


def _ec(curve_name):
    from zksnake_b200 import ecc
    return ecc.EllipticCurve(curve_name)


def _vector(curve_name, grp, scalars):
    E = _ec(curve_name)
    gen = E.G1() if grp == 1 else E.G2()
    return E.curve.batch_mul_device(gen, scalars, grp)


def _scalars(cid, count, seed):
    r = PARAMS[cid].r
    rnd = random.Random(seed)
    return [0, 1, 2, r - 1, 0] + [rnd.randint(1, r - 1) for _ in range(count - 5)]


def _encode_x(cid, g2, x, larger=False, inf=False, compressed=True):
    """hand-made encoding of an arbitrary x (not necessarily a valid point)"""
    nb = PARAMS[cid].fq_bytes
    xs = x if g2 else (x,)
    if cid == 0:
        out = bytearray(b"".join(c.to_bytes(nb, "little") for c in xs))
        out[-1] |= (0x80 if larger else 0) | (0x40 if inf else 0)
    else:
        out = bytearray(b"".join(c.to_bytes(nb, "big") for c in reversed(xs)))
        out[0] |= (0x80 if compressed else 0) | (0x40 if inf else 0) | (0x20 if larger else 0)
    return bytes(out)


@pytest.mark.parametrize("curve_name,grp", CASES)
def test_compress_matches_oracle(gpu, curve_name, grp):
    cid = curve_id(curve_name)
    G = group(cid, grp == 2)
    ks = _scalars(cid, 40, 7)
    vec = _vector(curve_name, grp, ks)
    want = b"".join(G.to_bytes(G.mul(G.gen, k)) for k in ks)
    assert vec.to_bytes() == want


@pytest.mark.parametrize("curve_name,grp", CASES)
def test_decompress_round_trip(gpu, curve_name, grp):
    cid = curve_id(curve_name)
    E = _ec(curve_name)
    ks = _scalars(cid, 70, 8)
    vec = _vector(curve_name, grp, ks)
    raw = vec.to_bytes()
    back = E.curve.PointVector.from_bytes(cid, grp, raw)
    assert len(back) == len(ks)
    assert np.array_equal(back.download(), vec.download())
    assert back.to_bytes() == raw
    # the host per-point decoder (the mirror of PointG1/G2.from_bytes) agrees on a few of them
    size = len(raw) // len(ks)
    cls = E.curve.PointG1 if grp == 1 else E.curve.PointG2
    for i in (0, 1, 3, 17):
        assert bytes(cls.from_bytes(raw[i * size:(i + 1) * size]).to_bytes()) == raw[i * size:(i + 1) * size]
    assert len(E.curve.PointVector.from_bytes(cid, grp, b"")) == 0


def _bad_encodings(cid, g2):
    """{reason: encoding} built from x values the ORACLE rejects for that reason"""
    G = group(cid, g2)
    q = PARAMS[cid].q
    F = G.F
    mk = (lambda k: (k, 1)) if g2 else (lambda k: k)
    bad = {}
    bad["coordinate not in field"] = _encode_x(cid, g2, (q, 0) if g2 else q)
    bad["non-zero infinity"] = _encode_x(cid, g2, mk(1), inf=True)
    if cid == 0:
        bad["invalid flags"] = _encode_x(cid, g2, mk(0) if not g2 else (0, 0), larger=True, inf=True)
    else:
        bad["uncompressed encoding"] = _encode_x(cid, g2, mk(5), compressed=False)
    k = 1
    while "not on curve" not in bad or ("not in the prime-order subgroup" not in bad and not (cid == 0 and not g2)):
        x = mk(k)
        k += 1
        y = F.sqrt(F.add(F.mul(F.mul(x, x), x), G.b))
        if y is None:
            bad.setdefault("not on curve", _encode_x(cid, g2, x))
        elif G.mul_raw((x, y), G.r) is not None:
            bad.setdefault("not in the prime-order subgroup", _encode_x(cid, g2, x, larger=G._y_is_larger(y)))
    for why, enc in bad.items():   # the oracle rejects each for exactly that reason
        with pytest.raises(ValueError, match=why):
            G.from_bytes(enc)
    return bad


@pytest.mark.parametrize("curve_name,grp", CASES)
def test_decompress_rejects_what_the_oracle_rejects(gpu, curve_name, grp):
    cid = curve_id(curve_name)
    E = _ec(curve_name)
    ks = _scalars(cid, 24, 9)
    good = _vector(curve_name, grp, ks).to_bytes()
    size = len(good) // len(ks)
    bad = _bad_encodings(cid, grp == 2)
    reason = {"uncompressed encoding": "invalid flags"}    # (one code for both curves' flag errors)
    cls = E.curve.PointG1 if grp == 1 else E.curve.PointG2
    for why, enc in bad.items():
        assert len(enc) == size
        with pytest.raises(ValueError, match=why):        # the host per-point mirror
            cls.from_bytes(enc)
        blob = good[:11 * size] + enc + good[12 * size:]
        with pytest.raises(ValueError, match=rf"Cannot deserialize point: {reason.get(why, why)} \(index 11\)"):
            E.curve.PointVector.from_bytes(cid, grp, blob)
    # the FIRST offending index is the one reported
    encs = list(bad.values())
    blob = good[:5 * size] + encs[0] + good[6 * size:19 * size] + encs[1] + good[20 * size:]
    with pytest.raises(ValueError, match=r"\(index 5\)"):
        E.curve.PointVector.from_bytes(cid, grp, blob)
    # without validation the subgroup check (only) is skipped
    if "not in the prime-order subgroup" in bad:
        blob = good[:3 * size] + bad["not in the prime-order subgroup"] + good[4 * size:]
        vec = E.curve.PointVector.from_bytes(cid, grp, blob, validate=False)
        assert vec.to_bytes() == blob
    with pytest.raises(ValueError, match="bad length"):
        E.curve.PointVector.from_bytes(cid, grp, good[:-1])


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_groth16_keys_round_trip_and_prove(gpu, curve_name):
    """tests/test_groth16.py:186-210 of the reference, plus: a prover given the re-read key produces the same proof bytes."""
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    from .test_gpu_groth16 import make, prove_seeded
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    circuit = rm.chain_circuit(37, curve_name)
    g, st, pub, priv = make(gm, rm, circuit, curve_name, seed=21)
    pk_bytes = g.proving_key.to_bytes()
    vk_bytes = g.verifying_key.to_bytes()
    # layout against the oracle's encodings of the closed-form key elements
    G1, G2 = group(cid), group(cid, True)
    tau, alpha, beta, gamma, delta = st.tau, st.alpha, st.beta, st.gamma, st.delta
    n = g.n
    head = (G1.to_bytes(G1.mul(G1.gen, alpha)) + G2.to_bytes(G2.mul(G2.gen, beta)) + G2.to_bytes(G2.mul(G2.gen, delta))
            + G1.to_bytes(G1.mul(G1.gen, beta)) + G1.to_bytes(G1.mul(G1.gen, delta)))
    assert pk_bytes[:len(head)] == head
    assert pk_bytes[len(head):len(head) + 8] == n.to_bytes(8, "little")
    size = PARAMS[cid].fq_bytes
    first_taus = b"".join(G1.to_bytes(G1.mul(G1.gen, pow(tau, i, r))) for i in range(4))
    assert pk_bytes[len(head) + 8:len(head) + 8 + 4 * size] == first_taus
    pk2 = gm.ProvingKey.from_bytes(pk_bytes, crv=curve_name)
    assert pk2.to_bytes() == pk_bytes
    vk2 = gm.VerifyingKey.from_bytes(vk_bytes, crv=curve_name)
    assert vk2.to_bytes() == vk_bytes
    rr, ss = random.Random(5).randint(1, r - 1), random.Random(6).randint(1, r - 1)
    proof1 = prove_seeded(gm, g, pub, priv, rr, ss)
    c2 = rm.chain_circuit(37, curve_name)[0]
    g2 = gm.Groth16(c2, curve_name)
    g2.proving_key, g2.verifying_key = pk2, vk2
    proof2 = prove_seeded(gm, g2, pub, priv, rr, ss)
    assert proof2.to_bytes() == proof1.to_bytes()
    assert g2.verify(proof2, pub)
    with pytest.raises(AssertionError):
        gm.ProvingKey.from_bytes(pk_bytes[:-5], crv=curve_name)


@pytest.mark.parametrize("curve_name,grp,log_n", [("BN254", 1, 18), ("BLS12_381", 1, 16), ("BN254", 2, 16)])
def test_bulk_sizes(gpu, curve_name, grp, log_n, capsys):
    """round trip at key-file sizes; prints the throughput (one from_hex call per point in the reference)"""
    cid = curve_id(curve_name)
    E = _ec(curve_name)
    r = PARAMS[cid].r
    n = 1 << log_n
    rnd = random.Random(11)
    ks = [rnd.randint(0, r - 1) for _ in range(n)]
    vec = _vector(curve_name, grp, ks)
    t0 = time.perf_counter()
    raw = vec.to_bytes()
    t1 = time.perf_counter()
    back = E.curve.PointVector.from_bytes(cid, grp, raw)
    t2 = time.perf_counter()
    assert np.array_equal(back.download(), vec.download())
    with capsys.disabled():
        print(f"\n[codec] {curve_name} G{grp} n=2^{log_n}: compress {1e3 * (t1 - t0):.1f} ms, decompress+validate {1e3 * (t2 - t1):.1f} ms")


@pytest.mark.parametrize("curve_name,n_gates", [("BN254", 13), ("BLS12_381", 8), ("BN254", 64)])
def test_plonk_keys_round_trip_and_prove(gpu, curve_name, n_gates):
    """plonk/serialization.py:157-353: the list-based prover's key and the device prover's key serialise to the SAME bytes;
    both provers, given the re-read keys, reproduce the oracle's proof bytes."""
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    from oracle import plonk as op
    from .test_gpu_plonk import seeded
    cs, pub, priv = chain_gates(n_gates, curve_name)
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(900 + n_gates)
    tau = rnd.randint(1, r - 1)
    blind = [rnd.randint(1, r - 1) for _ in range(11)]
    want, _ = op.prove(op.Circuit(cid, cs.qL, cs.qR, cs.qO, cs.qM, cs.qC, cs.permutation), tau, pub, priv, blind)
    host, dev = pm.Plonk(cs, curve_name), DevicePlonk(cs, curve_name)
    for prover in (host, dev):
        old = seeded(pm, [tau])
        try:
            prover.setup()
        finally:
            pm.get_random_int = old
    pk_bytes = host.proving_key.to_bytes()
    vk_bytes = host.verifying_key.to_bytes()
    assert dev.reference_proving_key().to_bytes() == pk_bytes
    assert dev.verifying_key.to_bytes() == vk_bytes
    with pytest.raises(AssertionError):
        dev.proving_key.to_bytes()           # the device prover's working key has no derived fields
    # layout: u64 SRS length, then [tau^i]G1 compressed (oracle encodings)
    G1 = group(cid)
    assert pk_bytes[:8] == (cs.length + 6).to_bytes(8, "little")
    size = PARAMS[cid].fq_bytes
    assert pk_bytes[8:8 + 3 * size] == b"".join(G1.to_bytes(G1.mul(G1.gen, pow(tau, i, r))) for i in range(3))
    assert vk_bytes[:8] == cs.length.to_bytes(8, "little")
    # re-read: reference-style (lists) and device-style (FrVecs)
    pk_host = pm.ProvingKey.from_bytes(pk_bytes, curve_name)
    assert pk_host.to_bytes() == pk_bytes and pk_host.n == cs.length
    pk_dev = pm.ProvingKey.from_bytes(pk_bytes, curve_name, device=True)
    assert pk_dev.to_bytes() == pk_bytes
    vk2 = pm.VerifyingKey.from_bytes(vk_bytes, curve_name)
    assert vk2.to_bytes() == vk_bytes
    host2 = pm.Plonk(cs, curve_name)
    host2.proving_key, host2.verifying_key = pk_host, vk2
    dev2 = DevicePlonk(cs, curve_name)
    dev2.load_keys(pk_dev, vk2)
    for prover in (host2, dev2):
        old = seeded(pm, blind)
        try:
            proof = prover.prove(pub, priv)
        finally:
            pm.get_random_int = old
        assert proof.to_bytes() == want
        assert prover.verify(proof, pub)
    with pytest.raises(AssertionError):
        pm.ProvingKey.from_bytes(pk_bytes[:-40], curve_name)


def test_groth16_from_circom_file(gpu, tmp_path):
    """a circuit written to and read back from a circom .r1cs file proves to the same bytes"""
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    from .test_gpu_groth16 import make, prove_seeded
    r = PARAMS[0].r
    circuit = rm.chain_circuit(21, "BN254")
    path = tmp_path / "chain.r1cs"
    rm.write_r1cs_file(path, circuit[0], n_pub_out=1)
    back, _ = rm.read_r1cs_file(path)
    proofs = []
    for c in (circuit[0], back):
        g, st, pub, priv = make(gm, rm, (c, circuit[1], circuit[2]), "BN254", seed=33)
        proofs.append(prove_seeded(gm, g, pub, priv, 1234567 % r, 7654321 % r).to_bytes())
        assert g.verify(gm.Proof.from_bytes(proofs[-1]), pub)
    assert proofs[0] == proofs[1]


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_g2_subgroup_criteria_agree_with_the_definition(gpu, curve_name):
    """validate=1 (endomorphism criterion) and validate=2 (r * P = infinity) against the oracle's r * P on members, random
    non-members of E'(Fq2), pure cofactor-torsion points and member + torsion sums."""
    cid = curve_id(curve_name)
    E = _ec(curve_name)
    G = group(cid, True)
    F, q, r = G.F, PARAMS[cid].q, PARAMS[cid].r
    rnd = random.Random(4242)
    pts, member = [], []
    while len(pts) < 36:
        x = (rnd.randint(0, q - 1), rnd.randint(0, q - 1))
        y = F.sqrt(F.add(F.mul(F.mul(x, x), x), G.b))
        if y is None:
            continue
        p0 = (x, y)
        t = G.mul_raw(p0, r)                         # the cofactor-torsion part of a random curve point
        cands = [p0, t, G.add(t, G.mul(G.gen, rnd.randint(1, r - 1))) if t is not None else None,
                 G.mul(G.gen, rnd.randint(1, r - 1))]
        for c in cands:
            if c is None:
                continue
            pts.append(c)
            member.append(G.mul_raw(c, r) is None)
    assert sum(member) >= 8 and sum(1 for m in member if not m) >= 20
    encs = [G.to_bytes(p) for p in pts]
    size = len(encs[0])
    for mode in (1, 2):
        # one at a time: the kernel's verdict per point
        for enc, ok in zip(encs, member):
            if ok:
                assert E.curve.PointVector.from_bytes(cid, 2, enc, validate=mode).to_bytes() == enc
            else:
                with pytest.raises(ValueError, match="not in the prime-order subgroup"):
                    E.curve.PointVector.from_bytes(cid, 2, enc, validate=mode)
        # all members in one vector pass; the first non-member of the whole list is the one reported
        good = b"".join(e for e, ok in zip(encs, member) if ok)
        assert E.curve.PointVector.from_bytes(cid, 2, good, validate=mode).to_bytes() == good
        first_bad = member.index(False)
        with pytest.raises(ValueError, match=rf"\(index {first_bad}\)"):
            E.curve.PointVector.from_bytes(cid, 2, b"".join(encs), validate=mode)
    assert E.curve.PointVector.from_bytes(cid, 2, b"".join(encs), validate=0).to_bytes() == b"".join(encs)
    assert len(encs) * size == len(b"".join(encs))
