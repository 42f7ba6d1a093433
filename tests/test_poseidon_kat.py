"""Known-answer check of Fr (BN254) arithmetic against vectors held by the reference's own test file
(/root/reference/tests/test_gadgets.py:19-50: five Poseidon hashes published by the Hades reference code and by circomlib -- two
implementations independent of this repository and of arkworks).  oracle/poseidon.py restates the permutation over a backend with
element-wise add / mul; here every field operation of it -- ~3 000 additions and multiplications per vector, chained, so one wrong
limb anywhere changes the result -- goes through

  * Python ints (the oracle's own arithmetic),
  * the HOST build of csrc/ff.cuh (zkb_test_field_op_host: the Montgomery multiplier text the GPU runs, carry flag emulated),
  * on the B200: the same text compiled for the device (zkb_test_field_op_dev) and the product's element-wise kernel
    (zkb_vec_op -> vec_op_kernel, the F4 row of SURVEY.md section 8: mul / add_over_evaluation_domain).
"""
import os
import random

import numpy as np
import pytest

from oracle import circom, poseidon
from oracle.fields import BN254, PARAMS

# byte copy of /root/reference/tests/stub/test_poseidon.r1cs (circomlib Poseidon(3), compiled by circom 2.1.6); see oracle/circom.py
CIRCOM_POSEIDON3 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "circom_poseidon3.r1cs")

R = PARAMS[BN254].r
FR_BN254 = 0          # field id of the self-test hooks (include/zkb200.h: 0 FrBN254)
NL = 8                # 32-bit limbs per element


def _pack32(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(NL * 4, "little") for v in vals), dtype=np.uint32).copy()


def _unpack32(arr, n):
    return [int.from_bytes(arr[i * NL:(i + 1) * NL].tobytes(), "little") for i in range(n)]


class HookBackend:
    """add / mul through zkb_test_field_op_host or _dev (op 0 = mul, 1 = add; canonical in, canonical out)."""

    def __init__(self, nat, fn):
        self.nat, self.fn = nat, fn

    def _op(self, op, xs, ys):
        n = len(xs)
        a, b = _pack32(xs), _pack32(ys)
        out = np.zeros(n * NL, dtype=np.uint32)
        self.nat.check(self.fn(FR_BN254, op, n, self.nat.ptr(a), self.nat.ptr(b), self.nat.ptr(out)))
        return _unpack32(out, n)

    def add(self, xs, ys):
        return self._op(1, xs, ys)

    def mul(self, xs, ys):
        return self._op(0, xs, ys)


class VecOpBackend:
    """add / mul through zkb_vec_op (host pointers in, vec_op_kernel on the device; op 0 = mul, 1 = add)."""

    def __init__(self, nat):
        self.nat = nat

    def _op(self, op, xs, ys):
        n = len(xs)
        a, b = self.nat.ints_to_limbs(list(xs)), self.nat.ints_to_limbs(list(ys))
        out = np.zeros((n, 4), dtype=np.uint64)
        self.nat.check(self.nat.lib.zkb_vec_op(BN254, op, n, self.nat.ptr(a), n, self.nat.ptr(b), n, self.nat.ptr(out)))
        return self.nat.limbs_to_ints(out)

    def add(self, xs, ys):
        return self._op(1, xs, ys)

    def mul(self, xs, ys):
        return self._op(0, xs, ys)


def test_parameters_are_the_published_ones():
    """First round constant and first matrix entry of the t = 3 instance as circomlib's poseidon_constants lists them; the matrix is
    a Cauchy matrix (every 2 x 2 minor non-zero is what makes it MDS -- checked for t = 3)."""
    constants, mds = poseidon.parameters(3)
    assert len(constants) == (8 + 57) * 3 and all(0 <= c < R for c in constants)
    assert constants[0] == 0x0EE9A592BA9A9518D05986D656F40C2114C4993C11BB29938D21D47304CD8E6E
    assert mds[0][0] == 0x109B7F411BA0E4C9B2B70CAF5C36A7B194BE7C11AD24378BFEDB68592BA8118B
    for i in range(3):
        for j in range(i + 1, 3):
            for k in range(3):
                for m in range(k + 1, 3):
                    assert (mds[i][k] * mds[j][m] - mds[i][m] * mds[j][k]) % R != 0


@pytest.mark.parametrize("inputs,expected", poseidon.REFERENCE_VECTORS)
def test_reference_vectors_python_ints(inputs, expected):
    assert poseidon.poseidon_hash(inputs, poseidon.IntBackend()) == expected


def test_full_permutation_output_of_the_hades_reference():
    """All three output words of poseidonperm_x5_254_3 on (0, 1, 2) (the hadeshash test vector the first reference entry is word 0 of)."""
    out = poseidon.permutation([0, 1, 2], poseidon.IntBackend())
    assert out == [0x115CC0F5E7D690413DF64C6B9662E9CF2A3617F2743245519E19607A4417189A,
                   0x0FCA49B798923AB0239DE1C9E7A4A9A2210312B6A2F616D18B5A87F9B628AE29,
                   0x0E7AE82E40091E63CBD4F16A6D16310B3729D4B6E138FCF54110E2867045A30C]


@pytest.mark.parametrize("inputs,expected", poseidon.REFERENCE_VECTORS)
def test_reference_vectors_host_build_of_the_gpu_multiplier(native, inputs, expected):
    assert poseidon.poseidon_hash(inputs, HookBackend(native, native.lib.zkb_test_field_op_host)) == expected


@pytest.mark.gpu
@pytest.mark.parametrize("inputs,expected", poseidon.REFERENCE_VECTORS)
def test_reference_vectors_on_the_device(gpu, inputs, expected):
    assert poseidon.poseidon_hash(inputs, HookBackend(gpu, gpu.lib.zkb_test_field_op_dev)) == expected
    assert poseidon.poseidon_hash(inputs, VecOpBackend(gpu)) == expected


# ---- the circom-compiled Poseidon(3) circuit of the reference's test fixtures -------------------------------------------------
def _circom_circuit_and_witness(inputs):
    from zksnake_b200 import r1cs as rm
    circuit, header = rm.read_r1cs_file(CIRCOM_POSEIDON3)
    first_input = 1 + header["n_pub_out"] + header["n_pub_in"]
    witness = circom.forward_witness(circuit.A.triplets, circuit.B.triplets, circuit.C.triplets, header["m_constraints"],
                                     header["n_wires"], {first_input + k: v for k, v in enumerate(inputs)}, R)
    return circuit, header, witness


def test_reader_on_a_real_circom_file_and_poseidon_through_its_constraints():
    """zksnake_b200.r1cs.read_r1cs_file on a file circom wrote (every other reader test uses files this repository wrote itself):
    header fields, wire order [1, h, a, b, c, intermediates], and -- solving the 261 rows forward from (a, b, c) -- wire h equals
    the Poseidon hash computed from the Grain-generated parameters (t = 4), i.e. the coefficients in the file are circomlib's
    constants and the rows mean <A,w> * <B,w> = <C,w>.  The product's own SparseArray.dot agrees row by row."""
    rnd = random.Random(7)
    for inputs in ([1, 2, 3], [0, 0, 0], [R - 1, 5, R - 2], [rnd.randrange(R) for _ in range(3)]):
        circuit, header, w = _circom_circuit_and_witness(inputs)
        assert (header["n_wires"], header["m_constraints"], header["n_pub_out"], header["n_pub_in"], header["n_priv_in"]) == (265, 261, 1, 0, 3)
        assert circuit.n_public == 2 and circuit.A.n_col == 265
        assert w[0] == 1 and w[2:5] == [v % R for v in inputs]
        assert w[1] == poseidon.poseidon_hash(inputs, poseidon.IntBackend())
        az, bz, cz = circuit.A.dot(w), circuit.B.dot(w), circuit.C.dot(w)
        assert all(x * y % R == z for x, y, z in zip(az, bz, cz))
    # (a, b, c) = (1, 2, 3): the value both routes agree on, for the record
    _, _, w = _circom_circuit_and_witness([1, 2, 3])
    assert w[1] == 6542985608222806190361240322586112750744169038454362455181422643027100751666
    # a wrong witness is caught by the same row check
    w[7] = (w[7] + 1) % R
    circuit, _, _ = _circom_circuit_and_witness([1, 2, 3])
    assert any(x * y % R != z for x, y, z in zip(circuit.A.dot(w), circuit.B.dot(w), circuit.C.dot(w)))


@pytest.mark.gpu
def test_groth16_over_the_circom_poseidon_circuit(gpu):
    """Groth16 setup / prove / verify on the B200 over the circom-compiled circuit with the forward-solved witness: device SpMV over
    a real circom constraint system, proof bytes == the oracle's closed form, verify() passes and rejects a wrong public output."""
    from oracle import groth16 as og
    from zksnake_b200 import groth16 as gm
    circuit, header, w = _circom_circuit_and_witness([1, 2, 3])
    pub, priv = w[:circuit.n_public], w[circuit.n_public:]
    rnd = random.Random(11)
    toxic = [rnd.randint(1, R - 1) for _ in range(5)]
    rr, ss = rnd.randint(1, R - 1), rnd.randint(1, R - 1)
    st = og.Setup(BN254, list(circuit.A.triplets), list(circuit.B.triplets), list(circuit.C.triplets), header["m_constraints"],
                  circuit.A.n_col, circuit.n_public, tuple(toxic))
    prover = gm.Groth16(circuit, "BN254")
    seq = iter(toxic + [rr, ss])
    old = gm.get_random_int
    gm.get_random_int = lambda n_max: next(seq)
    try:
        prover.setup()
        proof = prover.prove(pub, priv)
    finally:
        gm.get_random_int = old
    A, B, C = og.prove_closed_form(st, w, rr, ss)
    assert proof.to_bytes() == og.proof_bytes(BN254, A, B, C)
    assert prover.verify(proof, pub)
    assert not prover.verify(proof, [pub[0], (pub[1] + 1) % R])
