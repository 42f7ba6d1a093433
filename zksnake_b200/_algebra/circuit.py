"""Host-side stand-in for the reference's `zksnake._algebra.circuit` submodule (/root/reference/src/lib.rs:168-175: `Field`,
`ConstraintSystem`), so that the reference's UNMODIFIED Python layer -- `zksnake.arithmetization.{R1CS, Plonkish}`,
`zksnake.groth16.Groth16`, `zksnake.plonk.Plonk` -- and its own test-suite import and run over `zksnake_b200._algebra`.

This is the circuit FRONT END: symbolic expressions, witness solving and the lowering to R1CS rows / PlonK gates.  It is
outside the proving hot path (SURVEY.md section 8 marks it out of scope for acceleration; the reference runs it on the CPU as
well, src/arithmetization/{symbolic,r1cs,plonkish}.rs) and touches no field arithmetic beyond Python ints, so nothing here
runs on the GPU.  It restates the observable behaviour of those three files -- the lowering rules decide which linear
combination lands in A, B or C and the witness column order [1, outputs, public inputs, private inputs, intermediates]
(r1cs.rs:133-167), which the prover's key layout depends on -- with an own structure: expressions are immutable tuples and the
walkers are plain functions.

Large synthetic circuits do not come through here (the symbolic lowering is quadratic in the reference too): they are given as
triplets / gate lists (`zksnake_b200.r1cs`, `zksnake_b200.plonkish`) or wrapped by `ConstraintSystem.precompiled_*`.
"""
from collections import deque

# expression nodes: ("in", name) | ("const", value) | ("neg", x) | (op, left, right) with op in + - * /
_BINARY = {"+": "({} + {})", "-": "({} - {})", "*": "{} * {}", "/": "{} / {}"}


def _text(node):
    kind = node[0]
    if kind == "in":
        return node[1]
    if kind == "const":
        return str(node[1])
    if kind == "neg":
        return f"-({_text(node[1])})"
    return _BINARY[kind].format(_text(node[1]), _text(node[2]))


def _names(node, out):
    kind = node[0]
    if kind == "in":
        out.append(node[1])
    elif kind == "neg":
        _names(node[1], out)
    elif kind != "const":
        _names(node[1], out)
        _names(node[2], out)
    return out


def _value(node, env, p):
    """symbolic.rs:63-111 (missing variable -> KeyError, the reference's 'Missing one or more variable on evaluation')"""
    kind = node[0]
    if kind == "in":
        return env[node[1]] % p
    if kind == "const":
        return node[1] % p
    if kind == "neg":
        return (-_value(node[1], env, p)) % p
    a, b = _value(node[1], env, p), _value(node[2], env, p)
    if kind == "+":
        return (a + b) % p
    if kind == "-":
        return (a - b) % p
    if kind == "*":
        return a * b % p
    return a * pow(b, -1, p) % p     # "/": ValueError when b is not invertible ('Modular inverse not found')


def _solve_for(node, target, rhs):
    """Rearrange `node == rhs` into `target == <expression>` by peeling one operator at a time (symbolic.rs:133-189); the
    target must occur once, linearly."""
    kind = node[0]
    if kind == "in" and node[1] == target:
        return rhs
    if kind in ("+", "-", "*"):
        left, right = node[1], node[2]
        in_left = target in _names(left, [])
        in_right = target in _names(right, [])
        if not (in_left or in_right):
            raise ValueError(f"Target term not found in {kind} gate")
        if kind == "+":
            return _solve_for(left, target, ("-", rhs, right)) if in_left else _solve_for(right, target, ("-", rhs, left))
        if kind == "-":
            # left - right = rhs:  left = rhs + right.  (For the right operand the reference moves `left` across as rhs - left,
            # symbolic.rs:158-161, i.e. it solves for -right; reproduced as is.)
            return _solve_for(left, target, ("+", rhs, right)) if in_left else _solve_for(right, target, ("-", rhs, left))
        return _solve_for(left, target, ("/", rhs, right)) if in_left else _solve_for(right, target, ("/", rhs, left))
    raise ValueError(f"Unable to rearrange non-linear equation: {_text(node)} = {_text(rhs)}")


def _as_node(value, op):
    if isinstance(value, Field):
        return value.inner
    if isinstance(value, int) and not isinstance(value, bool) and value >= 0:
        return ("const", value)
    raise TypeError(f"Unsupported operand type for {op}")


class Equation:
    """symbolic.rs:213-258"""

    def __init__(self, left, right):
        self.lhs = left.inner if isinstance(left, Field) else left
        self.rhs = right.inner if isinstance(right, Field) else right

    def evaluate(self, inputs, modulus):
        return _value(self.lhs, inputs, modulus), _value(self.rhs, inputs, modulus)

    def swap(self):
        self.lhs, self.rhs = self.rhs, self.lhs

    def __repr__(self):
        return f"{_text(self.lhs)} = {_text(self.rhs)}"


class Field:
    """A symbolic field element / expression (symbolic.rs:255-436).  Note that `==` builds an Equation, so a Field is not
    hashable by value."""
    __slots__ = ("inner",)

    def __init__(self, var):
        self.inner = ("in", str(var))

    @classmethod
    def _of(cls, node):
        f = cls.__new__(cls)
        f.inner = node
        return f

    def evaluate(self, inputs, modulus):
        return _value(self.inner, inputs, modulus)

    def __add__(self, other):
        return Field._of(("+", self.inner, _as_node(other, "+")))

    __radd__ = __add__       # (the reference's __radd__ keeps self on the left as well: symbolic.rs:290-304)

    def __sub__(self, other):
        return Field._of(("-", self.inner, _as_node(other, "-")))

    __rsub__ = __sub__       # sic: the reference computes self - other for `other - self` too (symbolic.rs:322-336)

    def __neg__(self):
        return Field._of(("neg", self.inner))

    def __mul__(self, other):
        return Field._of(("*", self.inner, _as_node(other, "*")))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return Field._of(("/", self.inner, _as_node(other, "/")))

    __rtruediv__ = __truediv__   # sic (symbolic.rs:392-406)

    def __eq__(self, other):
        return Equation(self.inner, _as_node(other, "=="))

    __hash__ = object.__hash__

    def __repr__(self):
        return _text(self.inner)


# ---- lowering to R1CS rows (r1cs.rs:8-131) ---------------------------------------------------------------------------------
def _linear_terms(row, node, columns, p, negate, out):
    kind = node[0]
    if kind == "const":
        out.append((row, 0, (p - node[1]) if negate else node[1]))
    elif kind == "in":
        out.append((row, columns[node[1]], (p - 1) if negate else 1))
    elif kind == "+":
        _linear_terms(row, node[1], columns, p, negate, out)
        _linear_terms(row, node[2], columns, p, negate, out)
    elif kind == "-":
        _linear_terms(row, node[1], columns, p, negate, out)
        _linear_terms(row, node[2], columns, p, True, out)      # (the reference forces the sign here: r1cs.rs:42-45)
    elif kind == "neg":
        _linear_terms(row, node[1], columns, p, True, out)
    elif kind == "*":
        a, b = node[1], node[2]
        if a[0] == "in" and b[0] == "const":
            out.append((row, columns[a[1]], (p - b[1]) if negate else b[1]))
        elif a[0] == "const" and b[0] == "in":
            out.append((row, columns[b[1]], (p - a[1]) if negate else a[1]))
        else:
            raise ValueError(f"Invalid R1CS: {_text(node)}")
    else:
        raise ValueError(f"Invalid R1CS: {_text(node)}")


def _r1cs_row(row, eq, columns, p):
    a, b, c = [], [], []
    lhs, rhs = eq.lhs, eq.rhs
    kind = rhs[0]
    if kind == "*":
        _linear_terms(row, rhs[1], columns, p, False, a)
        _linear_terms(row, rhs[2], columns, p, False, b)
        _linear_terms(row, lhs, columns, p, False, c)
    elif kind == "/":                       # lhs = x / y  <=>  lhs * y = x
        _linear_terms(row, rhs[1], columns, p, False, c)
        _linear_terms(row, rhs[2], columns, p, False, b)
        _linear_terms(row, lhs, columns, p, False, a)
    else:                                   # linear right-hand side: rhs * 1 = lhs
        _linear_terms(row, rhs, columns, p, kind in ("-", "neg"), a)
        b.append((row, 0, 1))
        _linear_terms(row, lhs, columns, p, False, c)
    return a, b, c


# ---- lowering to PlonK gates (plonkish.rs:6-306) ---------------------------------------------------------------------------
class _GateScan:
    __slots__ = ("q", "const", "touched")

    def __init__(self, q):
        self.q, self.const, self.touched = q, 0, []


def _scan(node, st, p):
    kind = node[0]
    if kind == "in":
        st.touched.append(node[1])
    elif kind in ("+", "-"):
        _scan(node[1], st, p)
        _scan(node[2], st, p)
    elif kind == "*":
        a, b = node[1], node[2]
        if b[0] == "const":
            _scan(a, st, p)
            st.q *= b[1]
        elif a[0] == "const":
            st.q *= a[1]
            _scan(b, st, p)
        else:
            _scan(a, st, p)
            _scan(b, st, p)
    elif kind == "neg":
        _scan(node[1], st, p)
        st.q = p - st.q
    elif kind == "const":
        st.const += node[1]
    else:
        raise ValueError(f"Invalid plonkish constraint: {_text(node)}")


def _plonk_gate(eq, public, p):
    ql = qr = qo = qm = qc = 0
    w = ["", "", ""]
    lhs, rhs = eq.lhs, eq.rhs
    if lhs[0] == "const":
        qc = p - lhs[1]
    elif lhs[0] == "in":
        if lhs[1] not in public:
            qo = p - 1
        w[2] = lhs[1]
    else:
        raise ValueError(f"Constraint {eq!r} not in the form of C=A*B")
    kind = rhs[0]
    if kind == "const":
        qc += rhs[1]
    elif kind == "in":
        ql, w[0] = 1, rhs[1]
    elif kind in ("+", "-"):
        left, right = _GateScan(1), _GateScan(1)
        _scan(rhs[1], left, p)
        _scan(rhs[2], right, p)
        qc += left.const + right.const
        touched = left.touched + right.touched
        if len(touched) == 0:
            ql = qr = 0
        elif len(touched) == 1:
            ql, qr, w[0] = left.q % p, 0, touched[0]
        elif len(touched) == 2:
            ql = left.q % p
            if kind == "+":
                qr = right.q % p
            else:
                qr = 0 if touched[1] in public else p - (right.q % p)
            w[0], w[1] = touched
        else:
            raise ValueError(f"More than two variables in single gate: {eq!r}")
    elif kind == "*":
        st = _GateScan(1)
        _scan(rhs, st, p)
        if len(st.touched) == 0:
            qc = st.const
        elif len(st.touched) == 1:
            ql, w[0] = st.q % p, st.touched[0]
        elif len(st.touched) == 2:
            w[0], w[1] = st.touched
            qm = st.q % p
        else:
            raise ValueError(f"More than two variables in single gate: {eq!r}")
    elif kind == "neg":
        st = _GateScan(1)
        _scan(rhs[1], st, p)
        qc += st.const
        if st.touched:
            ql = 0 if st.touched[0] in public else p - (st.q % p)
            w[0] = st.touched[0]
    else:
        raise ValueError("Division operation is not supported")
    if not w[0] and w[1]:
        w[0], w[1] = w[1], w[0]
    return ql, qr, qo, qm, qc, w


def _copy_permutation(n_gates, wires):
    """plonkish.rs:255-283: column-major positions (a | b | c, each padded to the power of two); every occurrence of a wire
    name is swapped with the next one."""
    padded = 1 if n_gates <= 1 else 1 << (n_gates - 1).bit_length()
    flat = list(wires) + [""] * (3 * padded - len(wires))
    cols = [flat[k::3] for k in range(3)]
    names = cols[0] + cols[1] + cols[2]
    size = len(wires)
    perm = list(range(3 * padded))
    following = {}
    swaps = []
    for i in range(size - 1, -1, -1):           # next occurrence of the same name after i, within the first `size` positions
        name = names[i]
        if name:
            if name in following:
                swaps.append((i, following[name]))
            following[name] = i
    for i, j in reversed(swaps):
        perm[i], perm[j] = perm[j], perm[i]
    return perm


class ConstraintSystem:
    """symbolic.rs:438-832: constraints over named variables, witness solving, and the two lowerings."""

    def __init__(self, inputs, outputs, modulus):
        self.inputs, self.outputs = [str(x) for x in inputs], [str(x) for x in outputs]
        self.modulus = int(modulus)
        self._constraints = []
        self.vars = {}                       # name -> value, in first-seen order (the reference's HashMap order is arbitrary)
        self._public = []
        self._sequence = []                  # ("eq", Equation) | ("set", name, node) | ("hint", name, func, args)
        self._assigned = set(self.inputs)
        self._precompiled = None

    # ---- construction ----
    @property
    def constraints(self):
        return list(self._constraints)

    @property
    def public_vars(self):
        return list(self._public)

    def num_constraints(self):
        if self._precompiled:
            return self._precompiled["n_constraints"]
        return len(self._constraints)

    def num_witness(self):
        return len(self.vars)

    def _note_vars(self, node):
        for name in _names(node, []):
            self.vars.setdefault(name, 0)

    def add_variable(self, var):
        self._note_vars(var.inner)

    def set_public(self, var):
        items = var if isinstance(var, (list, tuple)) else [var]
        for item in items:
            if isinstance(item, str):
                self._public.append(item)
            elif isinstance(item, Field) and item.inner[0] == "in":
                self._public.append(item.inner[1])
            else:
                raise TypeError("Invalid expression")

    def add_constraint(self, constraint):
        eq = Equation(constraint.lhs, constraint.rhs)
        if eq.rhs[0] in ("in", "const") and eq.lhs[0] != "in":
            eq.swap()
        if eq.lhs[0] == "in":
            name = eq.lhs[1]
            if name not in self._assigned:
                self._assigned.add(name)
                self._sequence.append(("set", name, eq.rhs))
        else:
            found = _names(eq.lhs, [])
            if found and found[0] not in self._assigned:
                self._assigned.add(found[0])
                self._sequence.append(("set", found[0], _solve_for(eq.lhs, found[0], eq.rhs)))
        self._note_vars(eq.lhs)
        self._note_vars(eq.rhs)
        self._constraints.append(eq)
        self._sequence.append(("eq", eq))

    def unsafe_assign(self, target, func, args):
        if not isinstance(target, Field) or target.inner[0] != "in":
            raise TypeError("Invalid assignment expression")
        self._sequence.append(("hint", target.inner[1], func, [str(a) for a in args]))

    # ---- witness ----
    def evaluate(self, inputs):
        p = self.modulus
        known = set()
        for key in self.inputs:
            if key not in inputs:
                raise KeyError(f"All inputs and outputs variable must present: {key} is missing")
            if key in self.vars:
                self.vars[key] = int(inputs[key])
            known.add(key)
        queue = deque(self._sequence)
        budget = len(self._sequence) * 256
        while queue:
            step = queue.popleft()
            if step[0] == "eq":
                eq = step[1]
                left_names, right_names = _names(eq.lhs, []), _names(eq.rhs, [])
                unknown = [v for v in left_names + right_names if v not in known]
                if not unknown:
                    a, b = _value(eq.lhs, self.vars, p), _value(eq.rhs, self.vars, p)
                    assert a == b, f"{_text(eq.lhs)} != {_text(eq.rhs)}"
                else:
                    if len(unknown) == 1:
                        name = unknown[0]
                        expr = _solve_for(eq.lhs, name, eq.rhs) if name in left_names else _solve_for(eq.rhs, name, eq.lhs)
                        try:
                            self.vars[name] = _value(expr, self.vars, p)
                            known.add(name)
                        except (KeyError, ValueError):
                            pass
                    queue.append(step)
            elif step[0] == "set":
                _, name, node = step
                if all(v in known for v in _names(node, [])):
                    self.vars[name] = _value(node, self.vars, p)
                    known.add(name)
                else:
                    queue.append(step)
            else:
                _, name, func, args = step
                if all(a in known for a in args):
                    result = func(**{a: self.vars[a] for a in args})
                    if not isinstance(result, int):
                        raise TypeError("Non deterministic result must be Integer")
                    if name in self.vars:
                        self.vars[name] = result
                    known.add(name)
                else:
                    queue.append(step)
            budget -= 1
            if budget < 0:
                raise RuntimeError("Evaluation timeout: unique solution might not exist for the given constraints")

    def solve(self, inputs):
        if self._precompiled and "assignment" in self._precompiled:
            return dict(self._precompiled["assignment"])
        self.evaluate(inputs)
        return dict(self.vars)

    def get_witness_vector(self):
        """r1cs.rs:133-167: ["0", outputs..., public inputs..., private inputs..., intermediates...]"""
        if self._precompiled and "witness_vector" in self._precompiled:
            return list(self._precompiled["witness_vector"])
        pub_in, priv_in, rest = [], [], []
        for v in self.vars:
            if v in self.inputs:
                (pub_in if v in self._public else priv_in).append(v)
            elif v not in self.outputs:
                rest.append(v)
        return ["0"] + self.outputs + pub_in + priv_in + rest

    # ---- lowerings ----
    def compile_to_r1cs(self):
        if self._precompiled and "r1cs_rows" in self._precompiled:
            return self._precompiled["r1cs_rows"]
        columns = {name: i for i, name in enumerate(self.get_witness_vector())}
        return [_r1cs_row(i, eq, columns, self.modulus) for i, eq in enumerate(self._constraints)]

    def compile_to_plonkish(self):
        if self._precompiled and "gates" in self._precompiled:
            return self._precompiled["gates"], self._precompiled["permutation"]
        gates = [_plonk_gate(eq, self._public, self.modulus) for eq in self._constraints]
        wires = [name for g in gates for name in g[5]]
        return [(g[0], g[1], g[2], g[3], g[4], list(g[5])) for g in gates], _copy_permutation(len(gates), wires)

    # ---- large synthetic circuits, already lowered ---------------------------------------------------------------------------
    @classmethod
    def precompiled_r1cs(cls, rows, witness_vector, public_vars, modulus, assignment=None):
        """A constraint system whose R1CS lowering is given: rows[i] = (A_i, B_i, C_i), each a list of (row, column, value)
        triplets over the columns named by `witness_vector` (["0", ...]).  The benchmark circuits are synthesised this way
        (SURVEY.md section 8d): the symbolic lowering is O(n^2) in the reference as well."""
        cs = cls([], [], modulus)
        for name in witness_vector[1:]:
            cs.vars[name] = 0
        cs._public = list(public_vars)
        cs._precompiled = {"r1cs_rows": rows, "witness_vector": list(witness_vector), "n_constraints": len(rows)}
        if assignment is not None:
            cs._precompiled["assignment"] = assignment
        return cs

    @classmethod
    def precompiled_plonkish(cls, gates, permutation, public_vars, modulus, assignment=None):
        """The same for PlonK: gates[i] = (qL, qR, qO, qM, qC, [a, b, c] wire names), permutation over 3 * padded positions."""
        cs = cls([], [], modulus)
        for g in gates:
            for name in g[5]:
                if name:
                    cs.vars.setdefault(name, 0)
        cs._public = list(public_vars)
        cs._precompiled = {"gates": gates, "permutation": permutation, "n_constraints": len(gates)}
        if assignment is not None:
            cs._precompiled["assignment"] = assignment
        return cs
