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


@pytest.mark.parametrize("curve_name,n", [("BN254", 1 << 12), ("BLS12_381", 1 << 10), ("BN254", 1 << 16), ("BLS12_381", 1 << 16),
                                          ("BN254", 1 << 20), ("BLS12_381", 1 << 20)])   # the last two: BASELINE.json's full size
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


@pytest.mark.parametrize("curve_name,n", [("BN254", 1 << 12), ("BN254", 1 << 20)])
def test_prove_dense_random_matches_closed_form(gpu, curve_name, n):
    """SURVEY.md section 8d config 3, dense-random variant: A.w, B.w uniform in Fr, C.w their product, so U, V, H are full-size
    random scalars (the realistic MSM digit distribution) and the private witness has 3n columns.  Full size: n = 2^20."""
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    g, st, pub, priv = make(gm, rm, rm.dense_random_circuit(n, curve_name), curve_name, seed=7)
    rr, ss = random.Random(12).randint(1, r - 1), random.Random(13).randint(1, r - 1)
    proof = prove_seeded(gm, g, pub, priv, rr, ss)
    A, B, C = og.prove_closed_form(st, pub + priv, rr, ss)
    assert proof.to_bytes() == og.proof_bytes(cid, A, B, C)
    if n <= 1 << 12:
        assert g.verify(proof, pub)
        # every raw MSM against its discrete log
        G1, G2 = group(cid, False), group(cid, True)
        for i, (e, pt) in enumerate(zip(og.msm_exponents(st, pub + priv), g.last_msms())):
            G = G2 if i == 2 else G1
            assert as_oracle_point(pt) == G.mul(G.gen, e), i


def test_prove_accepts_unreduced_and_negative_witness_values(gpu):
    """Groth16.prove(list, list) marshals through csrc/pymarshal.cpp: values >= r go up raw and are reduced on the device, negative
    and > 256-bit values on the host -- the reference reduces them all in SparseArray.dot (array.py:43)."""
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    r = PARAMS[0].r
    g, st, pub, priv = make(gm, rm, rm.chain_circuit(37, "BN254"), "BN254")
    want = prove_seeded(gm, g, pub, priv, 5, 7).to_bytes()
    shifted = [v + r if i % 3 == 0 else (v - r if i % 3 == 1 else v + (r << 200)) for i, v in enumerate(priv)]
    assert prove_seeded(gm, g, [pub[0] + r, pub[1] - 2 * r], shifted, 5, 7).to_bytes() == want
    with pytest.raises(TypeError):
        g.prove(pub, [1.5] * len(priv))


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


@pytest.mark.parametrize("mode", ["windows", "points"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_partials_assemble_to_the_same_proof(gpu, mode, world):
    """Multi-GPU split emulated on one GPU: every "rank" runs zkb_groth16_partial over its shard (scalar
    windows or point ranges), the partial points are added and assembled -- the proof bytes must not depend on the split."""
    import ctypes

    import numpy as np

    from zksnake_b200 import _native as nat
    from zksnake_b200 import dist
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    curve_name, n = "BN254", 1 << 11
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rr, ss = random.Random(5).randint(1, r - 1), random.Random(6).randint(1, r - 1)
    g1, st, pub, priv = make(gm, rm, rm.chain_circuit(n, curve_name), curve_name)
    want = prove_seeded(gm, g1, pub, priv, rr, ss).to_bytes()
    A, B, C = og.prove_closed_form(st, pub + priv, rr, ss)
    assert want == og.proof_bytes(cid, A, B, C)
    toxic = [st.tau, st.alpha, st.beta, st.gamma, st.delta]    # the seeded toxic waste (the prover object keeps none)
    w = nat.ints_to_limbs([x % r for x in pub + priv])
    parts_xy, parts_inf, provers = [], [], []
    for rank in range(world):
        g = gm.Groth16(rm.chain_circuit(n, curve_name)[0], curve_name, shard=(rank, world), shard_mode=mode, emulate_shard=True)
        seq = iter(toxic)
        old = gm.get_random_int
        gm.get_random_int = lambda n_max: next(seq)
        try:
            g.setup()
        finally:
            gm.get_random_int = old
        xy = np.zeros((dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
        flags = np.zeros(dist.MSM_SLOTS, dtype=np.int32)
        nat.check(nat.lib.zkb_groth16_partial(g._pk_handle, g._r1cs_handle, nat.ptr(w), 0, g.n_public, nat.ptr(xy), nat.ptr(flags)))
        parts_xy.append(xy)
        parts_inf.append(flags)
        provers.append(g)
    sxy, sinf = dist.add_partials(cid, np.stack(parts_xy), np.stack(parts_inf))
    g = provers[0]
    g1b = nat.lib.zkb_affine_bytes(cid, 1) // 8
    g2b = nat.lib.zkb_affine_bytes(cid, 2) // 8
    oa, ob, oc = np.zeros(g1b, np.uint64), np.zeros(g2b, np.uint64), np.zeros(g1b, np.uint64)
    inf = (ctypes.c_int * 3)()
    nat.check(nat.lib.zkb_groth16_assemble(g._pk_handle, nat.ptr(sxy), nat.ptr(sinf), nat.ptr(nat.ints_to_limbs([rr])),
                                           nat.ptr(nat.ints_to_limbs([ss])), nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf))
    ec = g.ec
    got = gm.Proof(ec.PointG1._from_flat(oa, inf[0]), ec.PointG2._from_flat(ob, inf[1]), ec.PointG1._from_flat(oc, inf[2]))
    assert got.to_bytes() == want
    # the same with (r, s) handed to every rank beforehand (what the multi-GPU prover does): each rank multiplies ITS partial sums
    # of [U] and [V] by s and r and folds them into its HZ slot (bit 1 of that slot's flag), the assembly is additions only --
    # through both assembly routes (one C call over all ranks' slots; Python slot sums + zkb_groth16_assemble)
    lr, ls = nat.ints_to_limbs([rr]), nat.ints_to_limbs([ss])
    for route in ("partials", "python"):
        parts_xy, parts_inf = [], []
        for g in provers:
            nat.check(nat.lib.zkb_groth16_precompute(g._pk_handle, nat.ptr(lr), nat.ptr(ls)))
            xy = np.zeros((dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
            flags = np.zeros(dist.MSM_SLOTS, dtype=np.int32)
            nat.check(nat.lib.zkb_groth16_partial(g._pk_handle, g._r1cs_handle, nat.ptr(w), 0, g.n_public, nat.ptr(xy), nat.ptr(flags)))
            assert flags[3] & 2 and not any(int(f) & 2 for k, f in enumerate(flags) if k != 3)
            parts_xy.append(xy)
            parts_inf.append(flags)
        oa[:], ob[:], oc[:] = 0, 0, 0
        if route == "partials":
            axy, ainf = np.ascontiguousarray(np.stack(parts_xy)), np.ascontiguousarray(np.stack(parts_inf), dtype=np.int32)
            nat.check(nat.lib.zkb_groth16_assemble_partials(provers[0]._pk_handle, world, nat.ptr(axy), nat.ptr(ainf), nat.ptr(lr),
                                                            nat.ptr(ls), nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf))
        else:
            sxy, sinf = dist.add_partials(cid, np.stack(parts_xy), np.stack(parts_inf))
            assert sinf[3] & 2
            nat.check(nat.lib.zkb_groth16_assemble(provers[0]._pk_handle, nat.ptr(sxy), nat.ptr(sinf), nat.ptr(lr), nat.ptr(ls),
                                                   nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf))
        got = gm.Proof(ec.PointG1._from_flat(oa, inf[0]), ec.PointG2._from_flat(ob, inf[1]), ec.PointG1._from_flat(oc, inf[2]))
        assert got.to_bytes() == want, route
    # uneven shares of the [K w] MSM (what the chain-spreading prover does: ranks that transform nothing take more of its windows,
    # dist.kw_windows); explicit window ranges per rank, including empty ones, must add up to the same proof
    if mode == "windows":
        wins = ctypes.c_uint32()
        nat.check(nat.lib.zkb_groth16_pk_msm_info(provers[0]._pk_handle, 3, None, ctypes.byref(wins)))
        for table in ([dist.kw_windows(k, world, wins.value) for k in range(world)],
                      [(0, 0)] * (world - 1) + [(0, wins.value)],                       # one rank runs it all
                      [(0, 1)] + [(1, 0)] * (world - 2) + [(1, wins.value + 5)]):       # (counts past the last window are clipped)
            parts_xy, parts_inf = [], []
            for g, (first, count) in zip(provers, table):
                nat.check(nat.lib.zkb_groth16_pk_set_kw_windows(g._pk_handle, first, count, 1))
                nat.check(nat.lib.zkb_groth16_precompute(g._pk_handle, nat.ptr(lr), nat.ptr(ls)))
                xy = np.zeros((dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
                flags = np.zeros(dist.MSM_SLOTS, dtype=np.int32)
                nat.check(nat.lib.zkb_groth16_partial(g._pk_handle, g._r1cs_handle, nat.ptr(w), 0, g.n_public, nat.ptr(xy), nat.ptr(flags)))
                nat.check(nat.lib.zkb_groth16_pk_set_kw_windows(g._pk_handle, 0, 0, 0))
                parts_xy.append(xy)
                parts_inf.append(flags)
            axy, ainf = np.ascontiguousarray(np.stack(parts_xy)), np.ascontiguousarray(np.stack(parts_inf), dtype=np.int32)
            nat.check(nat.lib.zkb_groth16_assemble_partials(provers[0]._pk_handle, world, nat.ptr(axy), nat.ptr(ainf), nat.ptr(lr),
                                                            nat.ptr(ls), nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf))
            got = gm.Proof(ec.PointG1._from_flat(oa, inf[0]), ec.PointG2._from_flat(ob, inf[1]), ec.PointG1._from_flat(oc, inf[2]))
            assert got.to_bytes() == want, table
    # a mix of folded and plain slots is refused
    ainf[0, 3] &= 1
    rc = nat.lib.zkb_groth16_assemble_partials(provers[0]._pk_handle, world, nat.ptr(axy), nat.ptr(ainf), nat.ptr(lr), nat.ptr(ls),
                                               nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf)
    assert rc != 0
