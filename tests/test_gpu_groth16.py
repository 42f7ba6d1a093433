"""GPU parity of Groth16.prove (python/zksnake/groth16/protocol.py:115-165 in the reference) through the C ABI: proof bytes,
H coefficients and every raw MSM point against the oracle, with toxic waste and prover randomness fixed from seeds."""
import random

import pytest

from oracle import groth16 as og
from oracle import poly
from oracle.curve import group
from oracle.fields import PARAMS, curve_id

pytestmark = pytest.mark.gpu


def make(groth_mod, r1cs_mod, circuit, curve_name, seed=1):
    c, pub, priv = circuit
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(seed)
    toxic = [rnd.randint(1, r - 1) for _ in range(5)]
    n_rows = max(t[0] for t in c.A.triplets) + 1
    st = og.Setup(cid, list(c.A.triplets), list(c.B.triplets), list(c.C.triplets), n_rows, c.A.n_col, c.n_public, tuple(toxic))
    g = groth_mod.Groth16(c, curve_name)
    seq = iter(toxic)
    old = groth_mod.get_random_int
    groth_mod.get_random_int = lambda n_max: next(seq)
    try:
        g.setup()
    finally:
        groth_mod.get_random_int = old
    return g, st, pub, priv


def prove_seeded(groth_mod, g, pub, priv, r_rand, s_rand):
    seq = iter([r_rand, s_rand])
    old = groth_mod.get_random_int
    groth_mod.get_random_int = lambda n_max: next(seq)
    try:
        return g.prove(pub, priv)
    finally:
        groth_mod.get_random_int = old


def as_oracle_point(pt):
    if pt.is_zero():
        return None
    return (tuple(pt.x), tuple(pt.y)) if isinstance(pt.x, list) else (pt.x, pt.y)


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_prove_matches_oracle_small(gpu, curve_name):
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    circuits = [rm.readme_circuit(curve_name), rm.chain_circuit(8, curve_name), rm.chain_circuit(37, curve_name),
                rm.dense_random_circuit(16, curve_name)]
    for k, circuit in enumerate(circuits):
        g, st, pub, priv = make(gm, rm, circuit, curve_name, seed=10 + k)
        rr, ss = random.Random(2).randint(1, r - 1), random.Random(3).randint(1, r - 1)
        proof = prove_seeded(gm, g, pub, priv, rr, ss)
        A, B, C, U, V, W, H = og.prove_literal(st, pub + priv, rr, ss)
        assert proof.to_bytes() == og.proof_bytes(cid, A, B, C), k
        assert (as_oracle_point(proof.A), as_oracle_point(proof.B), as_oracle_point(proof.C)) == (A, B, C)
        u, v, h = g.last_polys()
        assert (poly.strip(u), poly.strip(v), poly.strip(h)) == (U, V, H)
        # every raw MSM result is the discrete-log closed form
        G1, G2 = group(cid, False), group(cid, True)
        exps = og.msm_exponents(st, pub + priv)
        msms = g.last_msms()
        for i, (e, pt) in enumerate(zip(exps, msms)):
            G = G2 if i == 2 else G1
            assert as_oracle_point(pt) == G.mul(G.gen, e), (k, i)
        # byte round trip + verification equation (independent pairing implementation)
        again = gm.Proof.from_bytes(proof.to_bytes(), curve_name)
        assert again.to_bytes() == proof.to_bytes()
        if k < 2:
            assert g.verify(proof, pub)
            forged = list(pub)
            forged[-1] = (forged[-1] + 1) % r
            assert not g.verify(proof, forged)


@pytest.mark.parametrize("curve_name,n", [("BN254", 1 << 12), ("BLS12_381", 1 << 10), ("BN254", 1 << 16)])
def test_prove_matches_closed_form_large(gpu, curve_name, n):
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    g, st, pub, priv = make(gm, rm, rm.chain_circuit(n, curve_name), curve_name)
    rr, ss = random.Random(2).randint(1, r - 1), random.Random(3).randint(1, r - 1)
    proof = prove_seeded(gm, g, pub, priv, rr, ss)
    A, B, C = og.prove_closed_form(st, pub + priv, rr, ss)
    assert proof.to_bytes() == og.proof_bytes(cid, A, B, C)
    if n <= 1 << 12:
        assert g.verify(proof, pub)


def test_bad_witness_raises(gpu):
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    g, st, pub, priv = make(gm, rm, rm.chain_circuit(8, "BN254"), "BN254")
    bad = list(priv)
    bad[3] = (bad[3] + 1) % PARAMS[0].r
    with pytest.raises(ValueError, match="Failed to evaluate"):
        g.prove(pub, bad)
    with pytest.raises(AssertionError):
        g.prove(pub, priv[:-1])


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_algebra_module_surface(gpu, curve_name):
    """The polynomial KATs of /root/reference/tests/test_algebra.py:6-26 through the _algebra mirror, plus multiexp."""
    from zksnake_b200._algebra import ec_bls12_381, ec_bn254, polynomial_bls12_381, polynomial_bn254
    pm = polynomial_bn254 if curve_name == "BN254" else polynomial_bls12_381
    ec = ec_bn254 if curve_name == "BN254" else ec_bls12_381
    p = pm.MODULUS
    mk = lambda c: pm.Polynomial(1, [(x, [(0, 0)]) for x in c], len(c))  # noqa: E731
    a, b = mk([1, 2, 3]), mk([2, 3, 4])
    assert (a * b).coeffs() == [2, 7, 16, 17, 12]
    q, rem = (a * b) / b
    assert q == a and rem.is_zero()
    assert (a + b).coeffs() == [3, 5, 7] and (a - b).coeffs() == [p - 1, p - 1, p - 1]
    assert (a * 2).coeffs() == [2, 4, 6] and a(2) == 17
    # mul_over_fft the way polynomial.py:151-165 drives the module
    fa, fb = pm.fft([1, 2, 3, 0, 0, 0, 0], 7), pm.fft([2, 3, 4, 0, 0, 0, 0], 7)
    ab = pm.mul_over_evaluation_domain(len(fa), fa, fb)
    assert mk(pm.ifft(ab, len(ab))).coeffs() == [2, 7, 16, 17, 12]
    assert pm.ifft(pm.fft([5, 6, 7, 8], 4), 4) == [5, 6, 7, 8]
    f = pm.fft([5, 6, 7, 8], 4)
    assert pm.coset_fft([5, 6, 7, 8], 4) == f[1:] + f[:1]
    tau = 123456789
    lag = pm.evaluate_lagrange_coefficients(8, tau)
    assert lag == og.lagrange_coeffs(curve_id(curve_name), 8, tau)
    assert pm.evaluate_vanishing_polynomial(8, tau) == (pow(tau, 8, p) - 1) % p
    g1, g2 = ec.g1(), ec.g2()
    pts = [g1 * k for k in (3, 5, 7)]
    assert ec.multiscalar_mul_g1(pts, [2, 4, 6]) == g1 * (6 + 20 + 42)
    assert ec.multiscalar_mul_g2([g2 * 3, g2 * 5], [10, 1]) == g2 * 35
    assert ec.batch_multi_scalar_g1(g1, [1, 2, 3]) == [g1, g1 * 2, g1 * 3]
    with pytest.raises(ValueError, match="mismatch"):
        ec.multiscalar_mul_g1(pts, [1])
