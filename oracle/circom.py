"""TEST INFRASTRUCTURE (only tests/ may import this): a forward witness solver for circom-compiled R1CS.

circom writes the constraints of a deterministic computation in an order in which almost every row introduces exactly one wire that
no earlier row has fixed -- the product's result in <C, w>, or (after its linear-constraint elimination) the one unknown of <A, w>
or <B, w>.  So the witness of such a circuit follows from its inputs by sweeping the rows until nothing changes, solving each row
that has exactly one unknown wire.  That is all this does (Python ints, no field library); it is NOT a general R1CS solver.

Used with tests/golden/circom_poseidon3.r1cs, a byte copy of the fixture /root/reference/tests/stub/test_poseidon.r1cs (the
reference's tests hold it next to test_poseidon.circom: circomlib's Poseidon(3) behind signals a, b, c -> h, compiled by circom
2.1.6): an artefact of a toolchain independent of this repository whose coefficients ARE circomlib's Poseidon constants.  Reading it
with zksnake_b200.r1cs.read_r1cs_file (which replaces /root/reference/python/zksnake/parser.py:37-90), solving it here and comparing
wire `h` with oracle/poseidon.py checks the reader on a real circom file, the row semantics <A,w> * <B,w> = <C,w> and the wire order
the prover assumes, against that toolchain (tests/test_poseidon_kat.py).
"""


def forward_witness(a_triplets, b_triplets, c_triplets, n_rows, n_wires, known, p):
    """known: {wire index: value} (wire 0 = 1 is added).  Returns the full witness list; raises if a wire stays undetermined or a
    fully determined row does not hold."""
    rows = []
    for trips in (a_triplets, b_triplets, c_triplets):
        side = [[] for _ in range(n_rows)]
        for i, j, v in trips:
            side[i].append((j, v % p))
        rows.append(side)
    w = [None] * n_wires
    w[0] = 1
    for j, v in known.items():
        w[j] = v % p

    def evaluate(row):
        acc, unknown = 0, []
        for j, v in row:
            if w[j] is None:
                unknown.append((j, v))
            else:
                acc = (acc + v * w[j]) % p
        return acc, unknown

    todo = set(range(n_rows))
    progress = True
    while todo and progress:
        progress = False
        for i in sorted(todo):
            (a, ua), (b, ub), (c, uc) = (evaluate(side[i]) for side in rows)
            n_unknown = len(ua) + len(ub) + len(uc)
            if n_unknown == 0:
                if a * b % p != c:
                    raise ValueError(f"row {i} does not hold")
            elif n_unknown != 1:
                continue
            elif uc:
                j, k = uc[0]
                w[j] = (a * b - c) * pow(k, p - 2, p) % p
            elif ua and b:
                j, k = ua[0]
                w[j] = (c * pow(b, p - 2, p) - a) * pow(k, p - 2, p) % p
            elif ub and a:
                j, k = ub[0]
                w[j] = (c * pow(a, p - 2, p) - b) * pow(k, p - 2, p) % p
            else:
                continue
            todo.discard(i)
            progress = True
    if todo or any(x is None for x in w):
        raise ValueError(f"{len(todo)} rows / {sum(x is None for x in w)} wires undetermined: not a forward-solvable circuit")
    return w
