"""R1CS containers with the layout Groth16 consumes in the reference, plus the synthetic circuits of the benchmark configs.

`SparseArray` mirrors /root/reference/python/zksnake/array.py:4-43 (triplets, triplets_map, n_row, n_col, dot); `R1CS` carries
what /root/reference/python/zksnake/arithmetization/r1cs.py:9-64 exposes to the prover: A, B, C, n_public and the witness
column order [1, outputs..., public inputs..., private inputs..., intermediates...] (src/arithmetization/r1cs.rs:133-167).
The symbolic circuit front end that normally produces these is out of scope; circuits are given as
triplets.
"""
import random

import numpy as np

from . import _native as nat
from ._algebra._poly import _R


class SparseArray:
    def __init__(self, matrix, n_row, n_col, p):
        self.p = p
        self.n_row = n_row
        self.n_col = n_col
        self.triplets_map = {}
        self.triplets = []
        for i, row in enumerate(matrix):
            for j, v in enumerate(row):
                if v != 0:
                    self.triplets.append((i, j, v))

    def append(self, triplets):
        for row, col, value in triplets:
            if value != 0:
                self.triplets_map.setdefault(row, []).append((col, value))
                self.triplets.append((row, col, value))

    def dot(self, vector):
        """array.py:36-43 -- kept for small circuits and as documentation; the prover uses the device SpMV."""
        result = [0] * self.n_row
        for row, col, value in self.triplets:
            result[row] += vector[col] * value
        return [x % self.p for x in result]

    def _arrays(self):
        """(rows, cols int64 arrays, values as (nnz, 4) uint64 limbs mod p) of the triplet list, without per-element Python big-int
        work when the values are small (the common case: R1CS coefficients are mostly +-1 and small constants)."""
        trip = self.triplets
        nnz = len(trip)
        cached = getattr(self, "_arrays_cache", None)
        if cached is not None and cached[0] == nnz:
            return cached[1]
        rows = np.fromiter((t[0] for t in trip), dtype=np.int64, count=nnz)
        cols = np.fromiter((t[1] for t in trip), dtype=np.int64, count=nnz)
        vals = np.zeros((nnz, 4), dtype=np.uint64)
        p = self.p
        # one-limb fast path only when EVERY value is in [0, 2^64) -- decided explicitly, not by numpy's overflow behaviour
        # (numpy 1.x wraps a -1 to 2^64 - 1 instead of raising)
        if all(0 <= t[2] < (1 << 64) for t in trip):
            vals[:, 0] = np.fromiter((t[2] for t in trip), dtype=np.uint64, count=nnz)
        elif nnz:
            vals = nat.ints_to_limbs([t[2] % p for t in trip], 32)
        self._arrays_cache = (nnz, (rows, cols, vals))     # (append() only ever grows the list, so the length is the version)
        return rows, cols, vals

    def _csr(self, major, minor, vals, n_major):
        order = np.argsort(major, kind="stable")
        ptr = np.zeros(n_major + 1, dtype=np.uint64)
        np.cumsum(np.bincount(major, minlength=n_major), out=ptr[1:])
        return ptr, minor[order].astype(np.uint32), np.ascontiguousarray(vals[order])

    def to_csr(self, n_rows):
        """(row_ptr uint64[n_rows+1], col uint32[nnz], val uint64[nnz,4]) sorted by row."""
        rows, cols, vals = self._arrays()
        return self._csr(rows, cols, vals, n_rows)

    def to_csr_transposed(self, n_cols):
        """CSR of the transposed matrix (one row per COLUMN of this one): (ptr uint64[n_cols+1], row uint32[nnz], val)."""
        rows, cols, vals = self._arrays()
        return self._csr(cols, rows, vals, n_cols)


class R1CS:
    """Compiled R1CS: three SparseArrays, the number of public columns (constant 1 included) and the column count."""

    def __init__(self, A, B, C, n_public, p):
        self.A, self.B, self.C = A, B, C
        self.n_public = n_public
        self.p = p

    @classmethod
    def from_triplets(cls, a, b, c, n_row, n_col, n_public, curve="BN254"):
        p = _R[0 if curve in ("BN254", "BN128", "ALT_BN128") else 1]
        arrays = []
        for trip in (a, b, c):
            s = SparseArray([], n_row, n_col, p)
            s.append(trip)
            arrays.append(s)
        return cls(arrays[0], arrays[1], arrays[2], n_public, p)

    def is_sat(self, witness):
        a, b, c = self.A.dot(witness), self.B.dot(witness), self.C.dot(witness)
        return all(x * y % self.p == z for x, y, z in zip(a, b, c))


def readme_circuit(curve="BN254", x=3):
    """y == x^3 + x + 5 (/root/reference/README.md:18-35) as the compiler lays it out: columns [1, y, x, v1], rows
    v1 = x*x ; y - x - 5 = v1*x   (A.w=[3,9], B.w=[3,3], C.w=[9,27] for x=3)."""
    p = _R[0 if curve in ("BN254", "BN128", "ALT_BN128") else 1]
    a = [(0, 2, 1), (1, 3, 1)]
    b = [(0, 2, 1), (1, 2, 1)]
    c = [(0, 3, 1), (1, 1, 1), (1, 2, p - 1), (1, 0, p - 5)]
    r1cs = R1CS.from_triplets(a, b, c, 2, 4, 2, curve)
    y = (x ** 3 + x + 5) % p
    return r1cs, [1, y], [x % p, x * x % p]


def chain_circuit(n_constraints, curve="BN254", inp=2):
    """/root/reference/benchmarks/benchmark_groth16.py:11-24: v0 = inp*inp, v_i = v_{i-1}*inp, out == v_{N-2}.
    Columns [1, out, inp, v0..v_{N-2}], n_public = 2.  Returns (r1cs, public, private)."""
    p = _R[0 if curve in ("BN254", "BN128", "ALT_BN128") else 1]
    N = n_constraints
    assert N >= 2
    a, b, c = [(0, 2, 1)], [(0, 2, 1)], [(0, 3, 1)]
    for i in range(1, N - 1):
        a.append((i, 3 + i - 1, 1))
        b.append((i, 2, 1))
        c.append((i, 3 + i, 1))
    a.append((N - 1, 3 + N - 2, 1))
    b.append((N - 1, 0, 1))
    c.append((N - 1, 1, 1))
    r1cs = R1CS.from_triplets(a, b, c, N, N + 2, 2, curve)
    v, cur = [], inp % p
    for _ in range(N - 1):
        cur = cur * inp % p
        v.append(cur)
    return r1cs, [1, v[-1]], [inp % p] + v


def dense_random_circuit(n_constraints, curve="BN254", seed=20):
    """Dense-random variant of SURVEY.md section 8d config 3: row i is  w_{1+i} * w_{1+N+i} = w_{1+2N+i}  with uniform random
    factors, so that A.w, B.w (hence U, V, H) are full-size random field elements -- the realistic MSM scalar distribution.
    Columns [1, x_0..x_{N-1}, y_0..y_{N-1}, z_0..z_{N-1}], n_public = 1."""
    p = _R[0 if curve in ("BN254", "BN128", "ALT_BN128") else 1]
    N = n_constraints
    rnd = random.Random(seed)
    xs = [rnd.randrange(p) for _ in range(N)]
    ys = [rnd.randrange(p) for _ in range(N)]
    zs = [x * y % p for x, y in zip(xs, ys)]
    a = [(i, 1 + i, 1) for i in range(N)]
    b = [(i, 1 + N + i, 1) for i in range(N)]
    c = [(i, 1 + 2 * N + i, 1) for i in range(N)]
    r1cs = R1CS.from_triplets(a, b, c, N, 3 * N + 1, 1, curve)
    return r1cs, [1], xs + ys + zs


# ---------------------------------------------------------------------------------------------------------- circom .r1cs files
R1CS_FILE_VERSIONS = (1,)   # parser.py:8


def read_r1cs_file(path):
    """iden3/circom binary `.r1cs` -> (R1CS, header).  The sections /root/reference/python/zksnake/parser.py:37-90 reads (magic,
    version, header = type 1, constraints = type 2, wire-to-label = type 3, in any order), but straight into the A, B, C triplets
    the prover consumes instead of symbolic equations for the circuit compiler (which is outside the proving path, SURVEY.md
    section 8): row i holds <A_i, w> * <B_i, w> = <C_i, w> over the file's own wire order
    [1, outputs, public inputs, private inputs, intermediates] -- the order of a circom witness file -- and
    n_public = 1 + n_pub_out + n_pub_in.  The curve is recognised from the header's prime."""
    with open(path, "rb") as f:
        data = f.read()
    magic = data[:4]
    assert magic == b"r1cs", f"Invalid magic bytes: {magic}"
    version = int.from_bytes(data[4:8], "little")
    assert version in R1CS_FILE_VERSIONS, f"Unsupported r1cs file version: {version}"
    n_section = int.from_bytes(data[8:12], "little")
    pos, header, raw_constraints, labels = 12, None, [], None
    for _ in range(n_section):
        kind = int.from_bytes(data[pos:pos + 4], "little")
        size = int.from_bytes(data[pos + 4:pos + 12], "little")
        body = memoryview(data)[pos + 12:pos + 12 + size]
        assert len(body) == size, "Truncated r1cs file"
        pos += 12 + size
        if kind == 1:
            fs = int.from_bytes(body[:4], "little")
            vals = np.frombuffer(body[4 + fs:4 + fs + 16], dtype="<u4")
            header = {"fs": fs, "prime": int.from_bytes(body[4:4 + fs], "little"), "n_wires": int(vals[0]), "n_pub_out": int(vals[1]),
                      "n_pub_in": int(vals[2]), "n_priv_in": int(vals[3]),
                      "n_labels": int.from_bytes(body[20 + fs:28 + fs], "little"),
                      "m_constraints": int.from_bytes(body[28 + fs:32 + fs], "little")}
        elif kind == 2:
            raw_constraints.append(body)
        elif kind == 3:
            labels = np.frombuffer(body, dtype="<u8")
    assert header is not None, "r1cs file without a header section"
    curves = {_R[0]: "BN254", _R[1]: "BLS12_381"}
    assert header["prime"] in curves, "r1cs file over an unsupported field"
    p, fs = header["prime"], header["fs"]
    trips = ([], [], [])
    row = 0
    for body in raw_constraints:
        off, end = 0, len(body)
        while off < end:
            for which in range(3):
                count = int.from_bytes(body[off:off + 4], "little")
                off += 4
                for _ in range(count):
                    wire = int.from_bytes(body[off:off + 4], "little")
                    value = int.from_bytes(body[off + 4:off + 4 + fs], "little") % p
                    off += 4 + fs
                    assert wire < header["n_wires"], "wire index out of range"
                    trips[which].append((row, wire, value))
            row += 1
    assert row == header["m_constraints"], "constraint count does not match the header"
    header["wire_labels"] = labels
    n_public = 1 + header["n_pub_out"] + header["n_pub_in"]
    return R1CS.from_triplets(trips[0], trips[1], trips[2], row, header["n_wires"], n_public, curves[p]), header


def write_r1cs_file(path, r1cs, n_pub_out=0, n_priv_in=0):
    """The inverse of read_r1cs_file (version 1, sections header / constraints / wire-to-label), for tests and for handing a
    synthetic circuit to other tools."""
    fs = 32
    n_rows = max((t[0] for arr in (r1cs.A, r1cs.B, r1cs.C) for t in arr.triplets), default=-1) + 1
    n_wires = r1cs.A.n_col
    rows = [([], [], []) for _ in range(n_rows)]
    for which, arr in enumerate((r1cs.A, r1cs.B, r1cs.C)):
        for i, j, v in arr.triplets:
            rows[i][which].append((j, v % r1cs.p))
    body = bytearray()
    for lcs in rows:
        for lc in lcs:
            body += len(lc).to_bytes(4, "little")
            for j, v in lc:
                body += j.to_bytes(4, "little") + v.to_bytes(fs, "little")
    n_pub_in = r1cs.n_public - 1 - n_pub_out
    head = (fs.to_bytes(4, "little") + r1cs.p.to_bytes(fs, "little") + n_wires.to_bytes(4, "little")
            + n_pub_out.to_bytes(4, "little") + n_pub_in.to_bytes(4, "little") + n_priv_in.to_bytes(4, "little")
            + n_wires.to_bytes(8, "little") + n_rows.to_bytes(4, "little"))
    labels = np.arange(n_wires, dtype="<u8").tobytes()
    with open(path, "wb") as f:
        f.write(b"r1cs" + (1).to_bytes(4, "little") + (3).to_bytes(4, "little"))
        for kind, content in ((2, bytes(body)), (1, head), (3, labels)):   # (constraints before the header: order is free)
            f.write(kind.to_bytes(4, "little") + len(content).to_bytes(8, "little") + content)
