"""The drop-in, dropped in: the reference's UNMODIFIED Python layer (python/zksnake: groth16/protocol.py:32-186,
plonk/protocol.py:39-647, polynomial.py, ecc.py, arithmetization/, commitment/polynomial/kzg.py) and its own test files run over
`zksnake_b200._algebra` registered as `zksnake._algebra` (src/lib.rs:178-185) -- every field and curve operation they make goes
through libzkb200.so on the GPU.

Where the reference package comes from: $ZKSNAKE_REF, /root/reference/python (build container) or baseline/_ref (the
git-ignored unmodified copy __graft_entry__.build() makes; it travels to the GPU box).  Nothing here reads /root/reference at
run time on the GPU box.

Checks: (1) proof bytes of the reference's Groth16 / Plonk classes over the mirror == zksnake_b200's own provers == the
oracle's closed form, with the randomness hook the reference's harnesses patch (`get_random_int`, protocol.py:11) seeded;
(2) the reference's own pytest files pass unmodified (`-p zksnake_b200.dropin`).
"""
import os
import random
import subprocess
import sys

import pytest

from oracle import groth16 as og
from oracle import plonk as op
from oracle.fields import PARAMS, curve_id

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref(gpu):
    from zksnake_b200 import dropin
    where = dropin.reference_python_dir()
    if where is None:
        pytest.fail("reference Python package not found: run __graft_entry__.build() where /root/reference exists (it fills "
                    "baseline/_ref) or set ZKSNAKE_REF")
    dropin.install()
    import zksnake._algebra as alg
    from zksnake_b200 import _algebra
    assert alg is _algebra and dropin.installed()
    return where


def seeded(mod, values):
    seq = iter(values)
    old = mod.get_random_int
    mod.get_random_int = lambda n_max: next(seq)
    return old


def reference_r1cs(circuit, curve_name):
    """zksnake_b200 R1CS triplets -> the reference's R1CS object (arithmetization/r1cs.py:9-41) through a pre-lowered
    ConstraintSystem: `compile()` fills A / B / C exactly as it does from the symbolic compiler's rows."""
    from zksnake.arithmetization.r1cs import R1CS
    from zksnake_b200._algebra.circuit import ConstraintSystem
    c, pub, priv = circuit
    n_rows = max(t[0] for t in c.A.triplets) + 1
    rows = [([], [], []) for _ in range(n_rows)]
    for which, arr in enumerate((c.A, c.B, c.C)):
        for t in arr.triplets:
            rows[t[0]][which].append(t)
    names = ["0"] + [f"w{i}" for i in range(1, c.A.n_col)]
    cs = ConstraintSystem.precompiled_r1cs(rows, names, names[1:c.n_public], c.p)
    r1cs = R1CS(cs, curve_name)
    r1cs.compile()
    assert r1cs.n_public == c.n_public and r1cs.A.n_col == c.A.n_col
    assert r1cs.is_sat(pub, priv)
    return r1cs


@pytest.mark.parametrize("curve_name", ["BN254", "BLS12_381"])
def test_reference_groth16_over_the_mirror(ref, curve_name):
    import zksnake.groth16.protocol as rp
    from zksnake.groth16 import Proof as RefProof
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    for k, make in enumerate((lambda: rm.readme_circuit(curve_name), lambda: rm.chain_circuit(37, curve_name),
                              lambda: rm.dense_random_circuit(16, curve_name), lambda: rm.chain_circuit(1 << 10, curve_name))):
        rnd = random.Random(40 + k)
        seeds = [rnd.randint(1, r - 1) for _ in range(7)]        # tau, alpha, beta, gamma, delta, then r, s
        # the reference's classes over the mirror
        circuit = make()
        _, pub, priv = circuit
        prover = rp.Groth16(reference_r1cs(circuit, curve_name), curve_name)
        old = seeded(rp, seeds)
        try:
            prover.setup()
            proof = prover.prove(pub, priv)
        finally:
            rp.get_random_int = old
        got = bytes(proof.to_bytes())
        assert prover.verify(proof, pub)
        assert prover.verify(RefProof.from_bytes(got, curve_name), pub)
        # zksnake_b200's own prover (device-resident key, one C-ABI call per proof), same randomness
        circuit2 = make()
        mine = gm.Groth16(circuit2[0], curve_name)
        old = seeded(gm, seeds)
        try:
            mine.setup()
            proof2 = mine.prove(pub, priv)
        finally:
            gm.get_random_int = old
        assert proof2.to_bytes() == got, (curve_name, k)
        # the oracle's closed form
        c = circuit2[0]
        n_rows = max(t[0] for t in c.A.triplets) + 1
        st = og.Setup(cid, list(c.A.triplets), list(c.B.triplets), list(c.C.triplets), n_rows, c.A.n_col, c.n_public, tuple(seeds[:5]))
        A, B, C = og.prove_closed_form(st, pub + priv, seeds[5], seeds[6])
        assert got == og.proof_bytes(cid, A, B, C), (curve_name, k)
        # keys written by the reference's serialiser are read back by the bulk codec and prove the same bytes
        if k == 1:
            pk_bytes = prover.proving_key.to_bytes()
            assert mine.proving_key.to_bytes() == pk_bytes
            assert mine.verifying_key.to_bytes() == prover.verifying_key.to_bytes()
        bad = list(priv)
        bad[-1] = (bad[-1] + 1) % r
        with pytest.raises(ValueError, match="Failed to evaluate"):
            prover.prove(pub, bad)


@pytest.mark.parametrize("curve_name,n_gates", [("BN254", 13), ("BLS12_381", 8), ("BN254", 256)])
def test_reference_plonk_over_the_mirror(ref, curve_name, n_gates):
    import zksnake.plonk.protocol as rp
    from zksnake.arithmetization.plonkish import Plonkish as RefPlonkish
    from zksnake.plonk import Proof as RefProof
    from zksnake_b200 import plonk as pm
    from zksnake_b200._algebra.circuit import ConstraintSystem
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates
    cs, pub, priv = chain_gates(n_gates, curve_name)
    cid = curve_id(curve_name)
    r = PARAMS[cid].r
    rnd = random.Random(900 + n_gates)
    seeds = [rnd.randint(1, r - 1) for _ in range(12)]           # tau, then the 11 blinding scalars in call order
    want, _ = op.prove(op.Circuit(cid, cs.qL, cs.qR, cs.qO, cs.qM, cs.qC, cs.permutation), seeds[0], pub, priv, seeds[1:])
    # the reference's Plonkish + Plonk over the mirror; the gate list goes in pre-lowered (compile() reads it as it reads the
    # symbolic compiler's output, arithmetization/plonkish.py:26-54)
    gates = [(cs.qL[i], cs.qR[i], cs.qO[i], cs.qM[i], cs.qC[i], ["", "", ""]) for i in range(cs.unpadded_length)]
    ref_cs = RefPlonkish(ConstraintSystem.precompiled_plonkish(gates, list(cs.permutation), [], cs.p), curve_name)
    ref_cs.compile()
    assert ref_cs.length == cs.length and ref_cs.is_sat(pub, list(priv))
    prover = rp.Plonk(ref_cs, curve_name)
    old = seeded(rp, seeds)
    try:
        prover.setup()
        proof = prover.prove(pub, list(priv))
    finally:
        rp.get_random_int = old
    got = bytes(proof.to_bytes())
    assert got == want
    assert prover.verify(RefProof.from_bytes(got, curve_name), pub)
    # zksnake_b200's device prover, same randomness
    dev = DevicePlonk(cs, curve_name)
    old = seeded(pm, seeds)
    try:
        dev.setup()
        proof2 = dev.prove(pub, priv)
    finally:
        pm.get_random_int = old
    assert proof2.to_bytes() == got
    assert prover.verify(RefProof.from_bytes(proof2.to_bytes(), curve_name), pub)      # cross-verification
    assert dev.verify(pm.Proof.from_bytes(got, curve_name), pub)


def test_reference_kzg_over_the_mirror(ref):
    from zksnake.commitment.polynomial import KZG
    from zksnake.polynomial import Polynomial
    kzg = KZG(16, "BN254")
    kzg.setup()
    poly = Polynomial([7, 0, 3, 1, 5, 0, 0, 9], kzg.order)
    c = kzg.commit(poly)
    proof, value = kzg.open(poly, 1234567)
    assert value == poly(1234567)
    assert kzg.verify(c, proof, 1234567, value)
    assert not kzg.verify(c, proof, 1234567, (value + 1) % kzg.order)


REFERENCE_TESTS = [
    # (file, pytest -k expression or None): everything the reference tests on the proving path; multivariate polynomials, IPA,
    # bulletproofs, GKR / sumcheck are outside it and need pieces of `_algebra` that are out of scope
    ("test_symbolic.py", None),
    ("test_r1cs_qap.py", None),
    ("test_algebra.py", "univariate"),
    ("test_groth16.py", None),
    ("test_plonk.py", None),
    ("test_commitment.py", "kzg"),
]


def test_reference_test_suite_over_the_mirror(ref):
    """The reference's own pytest files, unmodified, with the mirror plugged in by `-p zksnake_b200.dropin`."""
    base = os.path.dirname(ref) if os.path.basename(ref) == "python" else ref     # /root/reference or baseline/_ref
    tests = os.path.join(base, "tests")
    assert os.path.isdir(tests), tests
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, ref, os.environ.get("PYTHONPATH", "")]), ZKSNAKE_REF=ref)
    log_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(log_dir, exist_ok=True)
    summary = []
    failed = False
    for name, expr in REFERENCE_TESTS:
        cmd = [sys.executable, "-m", "pytest", "-p", "zksnake_b200.dropin", "-p", "no:cacheprovider", "-q", "-x", "--rootdir", base,
               os.path.join(tests, name)]
        if expr:
            cmd += ["-k", expr]
        res = subprocess.run(cmd, cwd=base, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1800)
        tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
        summary.append(f"{name}{' -k ' + expr if expr else ''}: rc={res.returncode} {tail}")
        if res.returncode != 0:
            failed = True
            summary.append(res.stdout[-4000:])
    text = "\n".join(summary)
    with open(os.path.join(log_dir, "dropin_reference_tests.log"), "w") as f:
        f.write(text + "\n")
    print(text)
    assert not failed, text
