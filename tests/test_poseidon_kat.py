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
import numpy as np
import pytest

from oracle import poseidon
from oracle.fields import BN254, PARAMS

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
