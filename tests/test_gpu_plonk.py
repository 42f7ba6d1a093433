"""GPU parity of Plonk.prove (python/zksnake/plonk/protocol.py:157-484 in the reference) and KZG commit/open
(commitment/polynomial/kzg.py:32-51): proof bytes against the oracle's independent route (coefficient arithmetic + closed-form
commitments) with tau and the 11 blinding scalars fixed from a seed, then verify() through the pairing."""
import random

import pytest

from oracle import plonk as op
from oracle.curve import group
from oracle.fields import PARAMS, curve_id

pytestmark = pytest.mark.gpu


def seeded(mod, values):
    seq = iter(values)
    old = mod.get_random_int
    mod.get_random_int = lambda n_max: next(seq)
    return old


@pytest.mark.parametrize("curve_name,n_gates", [("BN254", 4), ("BN254", 8), ("BN254", 13), ("BLS12_381", 8), ("BN254", 64)])
def test_prove_matches_oracle(gpu, curve_name, n_gates):
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(n_gates, curve_name)
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(100 + n_gates)
    tau = rnd.randint(1, r - 1)
    blind = [rnd.randint(1, r - 1) for _ in range(11)]
    want, aux = op.prove(op.Circuit(cid, cs.qL, cs.qR, cs.qO, cs.qM, cs.qC, cs.permutation), tau, pub, priv, blind)
    plonk = pm.Plonk(cs, curve_name)
    old = seeded(pm, [tau] + blind)
    try:
        plonk.setup()
        proof = plonk.prove(pub, priv)
    finally:
        pm.get_random_int = old
    assert proof.to_bytes() == want
    assert plonk.verify(proof, pub)
    again = pm.Proof.from_bytes(proof.to_bytes(), curve_name)
    assert again.to_bytes() == want and plonk.verify(again, pub)
    again.zeta_a = (again.zeta_a + 1) % r
    assert not plonk.verify(again, pub)


def test_prove_larger_circuit_verifies(gpu):
    """2^10 gates: every MSM and transform takes the multi-pass GPU paths; unseeded randomness; verify() must pass and a
    wrong public input must fail."""
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(1 << 10, "BN254")
    plonk = pm.Plonk(cs, "BN254")
    plonk.setup()
    proof = plonk.prove(pub, priv)
    assert len(proof.to_bytes()) == 9 * 32 + 192
    assert plonk.verify(proof, pub)
    k = next(iter(pub))
    assert not plonk.verify(proof, {k: (pub[k] + 5) % plonk.order})


def test_unsatisfied_copy_constraint_is_rejected(gpu):
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(8, "BN254")
    plonk = pm.Plonk(cs, "BN254")
    plonk.setup()
    bad = list(priv)
    bad[4] = (bad[4] + 1) % plonk.order    # b wire of gate 1 no longer equals `inp`
    with pytest.raises(AssertionError):
        plonk.prove(pub, bad)


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_kzg_commit_open_verify(gpu, curve_name):
    from zksnake_b200 import kzg as km
    from zksnake_b200.polynomial import Polynomial
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    G1 = group(cid, False)
    rnd = random.Random(9)
    tau = rnd.randint(1, r - 1)
    coeffs = [rnd.randint(0, r - 1) for _ in range(300)] + [0, 0]     # trailing zeros are stripped like ark's DensePolynomial
    k = km.KZG(400, curve_name)
    old = seeded(km, [tau])
    try:
        k.setup()
    finally:
        km.get_random_int = old
    poly = Polynomial(coeffs, r)
    c = k.commit(poly)
    want = G1.mul(G1.gen, op.peval(coeffs, tau, r))
    assert (c.x, c.y) == want
    z = rnd.randint(0, r - 1)
    proof, value = k.open(poly, z)
    assert value == op.peval(coeffs, z, r)
    q, rem = op.pdiv_linear(op.psub(op.strip(coeffs), [value], r), z, r)
    assert rem == 0 and (proof.x, proof.y) == G1.mul(G1.gen, op.peval(q, tau, r))
    assert k.verify(c, proof, z, value)
    assert not k.verify(c, proof, z, (value + 1) % r)
    zero = k.commit(Polynomial([0], r))
    assert zero.is_zero() and zero == k.zero_commitment()


@pytest.mark.parametrize("curve_name,n_gates", [("BN254", 4), ("BN254", 8), ("BN254", 13), ("BLS12_381", 8), ("BN254", 64),
                                                 ("BLS12_381", 50)])
def test_device_prover_matches_oracle(gpu, curve_name, n_gates):
    """DevicePlonk (every vector resident in HBM, scans / batch inversion / linear division as kernels) must produce the same
    proof bytes as the oracle and as the list-based prover."""
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(n_gates, curve_name)
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(500 + n_gates)
    tau = rnd.randint(1, r - 1)
    blind = [rnd.randint(1, r - 1) for _ in range(11)]
    want, aux = op.prove(op.Circuit(cid, cs.qL, cs.qR, cs.qO, cs.qM, cs.qC, cs.permutation), tau, pub, priv, blind)
    plonk = DevicePlonk(cs, curve_name)
    old = seeded(pm, [tau] + blind)
    try:
        plonk.setup()
        proof = plonk.prove(pub, priv)
    finally:
        pm.get_random_int = old
    assert proof.to_bytes() == want
    assert plonk.verify(proof, pub)
    assert {"round1", "round2", "round3", "round4", "round5"} <= set(plonk.timings)


def test_device_prover_large_and_bad_witness(gpu):
    """2^14 gates through the multi-pass kernels; the list-based prover with the same randomness gives the same bytes."""
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    n = 1 << 12
    cs, pub, priv = chain_gates(n, "BN254")
    r = PARAMS[0].r
    rnd = random.Random(77)
    vals = [rnd.randint(1, r - 1) for _ in range(12)]
    dev, host = DevicePlonk(cs, "BN254"), pm.Plonk(cs, "BN254")
    proofs = []
    for prover in (dev, host):
        old = seeded(pm, vals)
        try:
            prover.setup()
            proofs.append(prover.prove(pub, priv))
        finally:
            pm.get_random_int = old
    assert proofs[0].to_bytes() == proofs[1].to_bytes()
    assert dev.verify(proofs[0], pub) and host.verify(proofs[0], pub)
    bad = list(priv)
    bad[4] = (bad[4] + 1) % r
    with pytest.raises(AssertionError):
        dev.prove(pub, bad)


def test_device_prover_2p16_verifies(gpu):
    """2^16 gates: transforms on the 2^16 / 2^18 / 2^19 domains, MSMs of 65 542 points over the SRS table; verify() must pass."""
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(1 << 16, "BN254")
    plonk = DevicePlonk(cs, "BN254")
    plonk.setup()
    proof = plonk.prove(pub, priv)
    assert plonk.verify(proof, pub)
    k = next(iter(pub))
    assert not plonk.verify(proof, {k: (pub[k] + 1) % plonk.order})


def test_device_prover_2p20_commitments_match_closed_form(gpu):
    """BASELINE.json configs[3] at full size: 2^20 gates on BN254.  With tau known, every one of the nine commitments must be
    [P(tau)]G1 for the polynomial P the prover committed to -- P is downloaded from the device and Horner-evaluated in Python
    ints, the point comes from the oracle's scalar multiplication: this pins the nine 2^20-point MSMs over the SRS table.  The
    proof must verify, the six opening values must be the evaluations of those polynomials, and the window-sharded MSMs (two
    ranks, emulated) must add up to the same commitments."""
    import ctypes

    import numpy as np

    from zksnake_b200 import _native as nat
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    n = 1 << 20
    cs, pub, priv = chain_gates(n, "BN254")
    r = PARAMS[0].r
    G1 = group(0, False)
    rnd = random.Random(2020)
    tau = rnd.randint(1, r - 1)
    blind = [rnd.randint(1, r - 1) for _ in range(11)]
    plonk = DevicePlonk(cs, "BN254")
    plonk.keep_polys = True
    old = seeded(pm, [tau] + blind)
    try:
        plonk.setup()
        proof = plonk.prove(pub, priv)
    finally:
        pm.get_random_int = old
    assert plonk.verify(proof, pub)

    def horner(coeffs, x):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % r
        return acc

    names = {"a": "tau_a", "b": "tau_b", "c": "tau_c", "z": "tau_z", "t_lo": "tau_t_lo", "t_mid": "tau_t_mid", "t_hi": "tau_t_hi",
             "w_zeta": "tau_W_zeta", "w_zeta_omega": "tau_W_zeta_omega"}
    coeffs = {}
    for key, attr in names.items():
        coeffs[key] = plonk.last_polys[key].to_ints()
        pt = getattr(proof, attr)
        assert (pt.x, pt.y) == G1.mul(G1.gen, horner(coeffs[key], tau)), key
    # the opening values are evaluations of the committed polynomials at the transcript's zeta
    beta, gamma, alpha, zeta, v, u = plonk._challenges(proof, pub)
    assert proof.zeta_a == horner(coeffs["a"], zeta) and proof.zeta_b == horner(coeffs["b"], zeta)
    assert proof.zeta_c == horner(coeffs["c"], zeta)
    assert proof.zeta_omega == horner(coeffs["z"], zeta * plonk.omega % r)
    # two-rank window shards of the first-round batch add up to the same three commitments
    vecs = [plonk.last_polys[k] for k in ("a", "b", "c")]
    limbs = nat.lib.zkb_affine_bytes(0, 1) // 8
    acc = [plonk.E.curve.PointG1.identity() for _ in vecs]
    for rank in range(2):
        ptrs = (ctypes.c_void_p * 3)(*[x.ptr for x in vecs])
        lens = (ctypes.c_size_t * 3)(*[x.n for x in vecs])
        xy = np.zeros((3, limbs), dtype=np.uint64)
        inf = (ctypes.c_int * 3)()
        nat.check(nat.lib.zkb_msm_table_batch_dev(plonk.table, 3, ptrs, lens, rank, 2, nat.ptr(xy), inf))
        for i in range(3):
            acc[i] = acc[i] + plonk.E.curve.PointG1._from_flat(xy[i], inf[i])
    assert acc == [proof.tau_a, proof.tau_b, proof.tau_c]


def test_commitment_window_shards_add_up(gpu):
    """DevicePlonk with shard=(r, W): each "rank" returns its window shard of every commitment; emulated here by calling the
    batch MSM for every shard and adding the partial points -- must equal the single-rank commitments."""
    import ctypes

    import numpy as np

    from zksnake_b200 import _native as nat
    from zksnake_b200.frvec import FrVec
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(300, "BN254")
    plonk = DevicePlonk(cs, "BN254")
    plonk.setup()
    r = plonk.order
    rnd = random.Random(8)
    vecs = [FrVec.from_ints(0, [rnd.randrange(r) for _ in range(k)]) for k in (512, 518, 100)]
    want = plonk._commit_many(vecs)
    limbs = nat.lib.zkb_affine_bytes(0, 1) // 8
    world = 3
    acc = [plonk.E.curve.PointG1.identity() for _ in vecs]
    for rank in range(world):
        ptrs = (ctypes.c_void_p * 3)(*[v.ptr for v in vecs])
        lens = (ctypes.c_size_t * 3)(*[v.n for v in vecs])
        xy = np.zeros((3, limbs), dtype=np.uint64)
        inf = (ctypes.c_int * 3)()
        nat.check(nat.lib.zkb_msm_table_batch_dev(plonk.table, 3, ptrs, lens, rank, world, nat.ptr(xy), inf))
        for i in range(3):
            acc[i] = acc[i] + plonk.E.curve.PointG1._from_flat(xy[i], inf[i])
    assert acc == want
