"""CPU-only pins of the PlonK oracle (oracle/plonk.py): its proofs satisfy the protocol's verification equation, checked by the
host-side verifier of zksnake_b200.plonk (group arithmetic through the host path of the C ABI + the Python pairing): no GPU
compute is involved, so this runs in the `-m "not gpu"` suite."""
import random

import pytest

from oracle import plonk as op
from oracle.curve import group
from oracle.fields import PARAMS, curve_id


def to_point(E, pt):
    return E.curve.PointG1.identity() if pt is None else E.curve.PointG1(pt[0], pt[1])


def oracle_case(curve_name, n_gates, seed):
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(n_gates, curve_name)
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(seed)
    tau = rnd.randint(1, r - 1)
    blind = [rnd.randint(1, r - 1) for _ in range(11)]
    circ = op.Circuit(cid, cs.qL, cs.qR, cs.qO, cs.qM, cs.qC, cs.permutation)
    blob, aux = op.prove(circ, tau, pub, priv, blind)
    return cs, pub, priv, tau, blind, circ, blob, aux


def oracle_verifying_key(plonk_mod, E, circ, tau):
    """selector / permutation commitments in closed form, as VerifyingKey"""
    P = PARAMS[circ.curve]
    r, n = P.r, circ.n
    G1 = group(circ.curve, False)
    omega = P.omega(n.bit_length() - 1)
    roots = [pow(omega, i, r) for i in range(n)]
    ids = roots + [2 * w % r for w in roots] + [3 * w % r for w in roots]
    cm = lambda evals: to_point(E, G1.mul(G1.gen, op.peval(op.interpolate(evals, omega, r), tau, r)))  # noqa: E731
    tau_sel = {k: cm(v) for k, v in circ.q.items()}
    tau_perm = [cm([ids[circ.permutation[i + k * n]] for i in range(n)]) for k in range(3)]
    return plonk_mod.VerifyingKey(n, E.G2() * tau, tau_sel, tau_perm, E.name)


@pytest.mark.parametrize("curve_name,n_gates", [("BN254", 4), ("BN254", 7), ("BLS12_381", 8)])
def test_oracle_proof_satisfies_the_verification_equation(native, curve_name, n_gates):
    from zksnake_b200 import plonk as pm
    cs, pub, priv, tau, blind, circ, blob, aux = oracle_case(curve_name, n_gates, seed=n_gates)
    assert cs.is_sat(pub, priv)
    assert len(blob) == 9 * (32 if curve_name == "BN254" else 48) + 192
    v = pm.Plonk(cs, curve_name)
    v.verifying_key = oracle_verifying_key(pm, v.E, circ, tau)
    proof = pm.Proof.from_bytes(blob, curve_name)
    assert proof.to_bytes() == blob
    assert v.verify(proof, pub)
    # any single tampered opening value breaks the equation
    bad = pm.Proof.from_bytes(blob, curve_name)
    bad.zeta_b = (bad.zeta_b + 1) % v.order
    assert not v.verify(bad, pub)
    # a different public input does not verify
    k = next(iter(pub))
    assert not v.verify(proof, {k: (pub[k] + 1) % v.order})


def test_oracle_rejects_bad_witness(native):
    cs, pub, priv, tau, blind, circ, blob, aux = oracle_case("BN254", 8, seed=3)
    bad = list(priv)
    bad[5] = (bad[5] + 1) % PARAMS[0].r       # c wire of gate 1: breaks a gate and a copy constraint
    assert not cs.is_sat(pub, bad)
    with pytest.raises(AssertionError):
        op.prove(circ, tau, pub, bad, blind)


def test_transcript_and_padding_conventions(native):
    """the byte conventions every challenge depends on (transcript.py:42-71) and the _pad_coeffs rule (polynomial.py:126-148)"""
    from zksnake_b200.polynomial import _pad_coeffs, next_power_of_two
    from zksnake_b200.transcript import FiatShamirTranscript
    import hashlib
    t = FiatShamirTranscript(field=PARAMS[0].r)
    t.append(5)            # bit_length 3 -> three bytes 00 00 05
    t.append(0)            # zero bytes
    t.append([256, 1])     # 9 bytes and 1 byte
    h = hashlib.blake2b(b"")
    h.update(b"\x00\x00\x05")
    h.update((256).to_bytes(9, "big"))
    h.update(b"\x01")
    d = h.digest()
    assert t.get_challenge_scalar() == int.from_bytes(d, "big") % PARAMS[0].r
    h2 = hashlib.blake2b(d)
    assert t.get_challenge() == h2.digest()    # re-seeded with the previous digest
    assert [next_power_of_two(k) for k in (0, 1, 2, 3, 5, -1)] == [2, 1, 2, 4, 8, 4]
    a, b = _pad_coeffs([1], [2])               # two degree-0 operands: 2 zeros each -> domain 4
    assert (len(a), len(b)) == (3, 3)
    a, b = _pad_coeffs([1] * 16, [1] * 16)     # full-degree U, V of length n -> 2n - ... -> domain 2n
    assert len(a) == len(b) == 16 + 16
    a, b = _pad_coeffs([1] * 28, [1] * 11)     # PlonK round 3 shape (3n+4, n+3 with n = 8): both to the same length
    assert len(a) == len(b) == 28 + 32
