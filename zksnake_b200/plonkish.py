"""PlonK arithmetisation container: selector vectors, copy permutation and witness layout that Plonk.setup / Plonk.prove
consume -- the fields of the reference's `Plonkish` (python/zksnake/arithmetization/plonkish.py:10-56, 98-126) without the
symbolic compiler behind it (`circuit::ConstraintSystem`, out of scope): circuits are given as gate lists.

Gate i:  qL a_i + qR b_i + qM a_i b_i + qO c_i + qC + PI_i = 0.   Wire positions: a -> i, b -> n + i, c -> 2n + i.
`permutation[pos]` is the next position of pos's copy class (a fixed point for an unconstrained wire).
Witness convention (plonkish.py:65-96): private_witness = [a_0, b_0, c_0, a_1, ...]; a public value sitting on wire c of gate i
is replaced by 0 there and enters as public_witness[i] = -value mod p."""
from .polynomial import BLS12_381_SCALAR_FIELD, BN254_SCALAR_FIELD, next_power_of_two

_FIELDS = {"BN128": BN254_SCALAR_FIELD, "BN254": BN254_SCALAR_FIELD, "ALT_BN128": BN254_SCALAR_FIELD,
           "BLS12_381": BLS12_381_SCALAR_FIELD}


class Plonkish:
    def __init__(self, gates, copy_classes, curve="BN254"):
        """gates: list of (qL, qR, qO, qM, qC); copy_classes: iterable of lists of (wire, gate) with wire in "abc"."""
        self.p = _FIELDS[curve]
        self.curve = curve
        self.unpadded_length = len(gates)
        self.length = n = next_power_of_two(len(gates))
        pad = [0] * (n - len(gates))
        p = self.p
        self.qL = [g[0] % p for g in gates] + pad
        self.qR = [g[1] % p for g in gates] + pad
        self.qO = [g[2] % p for g in gates] + pad
        self.qM = [g[3] % p for g in gates] + pad
        self.qC = [g[4] % p for g in gates] + pad
        perm = list(range(3 * n))
        base = {"a": 0, "b": n, "c": 2 * n}
        for cls in copy_classes:
            pos = [base[w] + i for w, i in cls]
            for j, src in enumerate(pos):
                perm[src] = pos[(j + 1) % len(pos)]
        self.permutation = perm

    def is_sat(self, public_witness, private_witness):
        """plonkish.py:98-126."""
        p, n = self.p, self.length
        a, b, c = list(private_witness[::3]), list(private_witness[1::3]), list(private_witness[2::3])
        for i in range(self.unpadded_length):
            pi = public_witness.get(i, 0)
            if (self.qL[i] * a[i] + self.qR[i] * b[i] + self.qM[i] * a[i] * b[i] + self.qO[i] * c[i] + self.qC[i] + pi) % p:
                return False
        flat = a + [0] * (n - len(a)) + b + [0] * (n - len(b)) + c + [0] * (n - len(c))
        return all(flat[src] == flat[dst] for src, dst in enumerate(self.permutation))


def chain_gates(n_constraints, curve="BN254", inp=2):
    """The multiplication chain of benchmarks/benchmark_plonk.py:12-25 (v0 = inp*inp, v_i = v_{i-1}*inp, out == v_{N-2}) as N
    gates: N-1 multiplication gates (qM = 1, qO = -1) and one output gate a - c + PI = 0 whose c wire carries the public value.
    Returns (Plonkish, public_witness, private_witness)."""
    N = n_constraints
    assert N >= 2
    p = _FIELDS[curve]
    gates = [(0, 0, -1, 1, 0)] * (N - 1) + [(1, 0, -1, 0, 0)]
    v, cur = [], inp % p
    for _ in range(N - 1):
        cur = cur * inp % p
        v.append(cur)
    witness = []
    for i in range(N - 1):
        witness += [inp % p if i == 0 else v[i - 1], inp % p, v[i]]
    witness += [v[N - 2], 0, 0]          # output gate: a = v_{N-2}; the public `out` on wire c is replaced by 0
    public = {N - 1: -v[N - 2] % p}
    classes = [[("a", 0)] + [("b", i) for i in range(N - 1)]]            # every use of `inp`
    classes += [[("c", i), ("a", i + 1)] for i in range(N - 1)]           # v_i feeds the next gate
    return Plonkish(gates, classes, curve), public, witness
