"""Groth16 over libzkb200.so -- same public surface as the reference's `zksnake.groth16`
(/root/reference/python/zksnake/groth16/protocol.py: Groth16(r1cs, curve).setup() / .prove(public, private) / .verify(proof,
public); serialization.py: Proof, ProvingKey, VerifyingKey).

What differs is where the work happens: the proving key lives in HBM (four device-resident point vectors), the R1CS lives
there too (CSR), and prove() is ONE call into the C ABI (zkb_groth16_prove_witness): witness up, three points down.  The
randomness hook is the module-level `get_random_int`, the name the reference's harnesses monkeypatch (protocol.py:11).
"""
import ctypes
import os
import random
import time

import numpy as np

from . import _native as nat
from . import dist
from ._algebra import ec_bls12_381, ec_bn254, polynomial_bls12_381, polynomial_bn254
from .frvec import FrVec
from .r1cs import R1CS

_CURVES = {"BN128": 0, "BN254": 0, "ALT_BN128": 0, "BLS12_381": 1}
_EC = {0: ec_bn254, 1: ec_bls12_381}
_POLY = {0: polynomial_bn254, 1: polynomial_bls12_381}
POINT_SIZE = {0: 32, 1: 48}  # CurvePointSize, /root/reference/python/zksnake/ecc.py:40-44


def get_random_int(n_max):
    """utils.py:6-9"""
    return random.SystemRandom().randint(1, n_max)


def next_power_of_two(n):
    """utils.py:26-28"""
    return 1 << (n - 1).bit_length()


class Proof:
    """serialization.py:5-42: A (G1) || B (G2) || C (G1), compressed."""

    def __init__(self, A, B, C):
        self.A, self.B, self.C = A, B, C

    def __str__(self):
        return f"A = {self.A}\nB = {self.B}\nC = {self.C}"

    __repr__ = __str__

    def to_bytes(self):
        return bytes(self.A.to_bytes() + self.B.to_bytes() + self.C.to_bytes())

    @classmethod
    def from_bytes(cls, s, crv="BN254"):
        ec = _EC[_CURVES[crv]]
        n = POINT_SIZE[_CURVES[crv]]
        assert len(s) == n * 4, f"Length of the Proof must equal {n * 4} bytes"
        return cls(ec.PointG1.from_bytes(s[:n]), ec.PointG2.from_bytes(s[n:3 * n]), ec.PointG1.from_bytes(s[3 * n:]))


def _take(buf, pos, size):
    if pos + size > len(buf):
        raise AssertionError("Invalid key length")
    return buf[pos:pos + size], pos + size


class ProvingKey:
    """serialization.py:45-159.  The four vectors are device-resident PointVectors (zksnake_b200._algebra._ec.PointVector);
    to_bytes / from_bytes produce and read the reference's byte layout with ONE bulk (de)compression kernel per vector instead
    of one from_hex call per point."""

    def __init__(self, alpha_G1, beta_G1, beta_G2, delta_G1, delta_G2, tau_G1, tau_G2, target_G1, k_delta_G1):
        self.alpha_1, self.beta_1, self.beta_2 = alpha_G1, beta_G1, beta_G2
        self.delta_1, self.delta_2 = delta_G1, delta_G2
        self.tau_1, self.tau_2, self.target_1, self.kdelta_1 = tau_G1, tau_G2, target_G1, k_delta_G1

    def to_bytes(self) -> bytes:
        """serialization.py:131-159: alpha_1 | beta_2 | delta_2 | beta_1 | delta_1, then each vector as u64 length + points."""
        s = bytes(self.alpha_1.to_bytes() + self.beta_2.to_bytes() + self.delta_2.to_bytes() + self.beta_1.to_bytes()
                  + self.delta_1.to_bytes())
        for vec in (self.tau_1, self.tau_2, self.target_1, self.kdelta_1):
            s += len(vec).to_bytes(8, "little") + vec.to_bytes()
        return s

    @classmethod
    def from_bytes(cls, b: bytes, crv="BN254"):
        """serialization.py:68-129"""
        cid = _CURVES[crv]
        ec = _EC[cid]
        n = POINT_SIZE[cid]
        b = bytes(b)
        assert len(b) >= 7 * n, "Invalid proving key length"
        alpha_1 = ec.PointG1.from_bytes(b[:n])
        beta_2 = ec.PointG2.from_bytes(b[n:3 * n])
        delta_2 = ec.PointG2.from_bytes(b[3 * n:5 * n])
        beta_1 = ec.PointG1.from_bytes(b[5 * n:6 * n])
        delta_1 = ec.PointG1.from_bytes(b[6 * n:7 * n])
        pos = 7 * n
        vecs = []
        for group in (1, 2, 1, 1):
            head, pos = _take(b, pos, 8)
            raw, pos = _take(b, pos, int.from_bytes(head, "little") * n * group)
            vecs.append(ec.PointVector.from_bytes(cid, group, raw))
        return cls(alpha_1, beta_1, beta_2, delta_1, delta_2, *vecs)


class VerifyingKey:
    """serialization.py:162-220."""

    def __init__(self, alpha_G1, beta_G2, gamma_G2, delta_G2, ic):
        self.alpha_1, self.beta_2, self.gamma_2, self.delta_2, self.ic = alpha_G1, beta_G2, gamma_G2, delta_G2, ic

    def to_bytes(self) -> bytes:
        s = bytes(self.alpha_1.to_bytes() + self.beta_2.to_bytes() + self.gamma_2.to_bytes() + self.delta_2.to_bytes())
        s += len(self.ic).to_bytes(8, "little")
        for pt in self.ic:
            s += bytes(pt.to_bytes())
        return s

    @classmethod
    def from_bytes(cls, s: bytes, crv="BN254"):
        cid = _CURVES[crv]
        ec = _EC[cid]
        n = POINT_SIZE[cid]
        s = bytes(s)
        assert len(s) >= n * 7, "Invalid verifying key length"
        alpha_1 = ec.PointG1.from_bytes(s[:n])
        beta_2, gamma_2, delta_2 = (ec.PointG2.from_bytes(s[n + 2 * n * k:3 * n + 2 * n * k]) for k in range(3))
        rest = s[7 * n + 8:]   # (the reference skips the length header and reads to the end of the buffer)
        ic = [ec.PointG1.from_bytes(rest[i:i + n]) for i in range(0, len(rest), n)]
        return cls(alpha_1, beta_2, gamma_2, delta_2, ic)


class Groth16:
    def __init__(self, r1cs: R1CS, curve: str = "BN254", shard=None, shard_mode="windows", tables=None, emulate_shard=False):
        """shard = (rank, world): this process proves cooperatively with the other ranks (zksnake_b200/dist.py).  Default: the
        torch.distributed world, else (0, 1).  shard_mode "windows" (default): every rank holds the whole key and runs the
        scalar windows [W*rank/world, W*(rank+1)/world) of every MSM -- sort, accumulation and bucket reduction all shrink by
        1/world.  "points": every rank keeps only its contiguous 1/world slice of the four key vectors (1/world of the memory;
        only the accumulation shrinks).  An explicit shard must match the torch.distributed world (dist.require_world);
        emulate_shard=True is for tests that drive zkb_groth16_partial / zkb_groth16_assemble rank by rank in ONE process."""
        self.rank, self.world = shard if shard is not None else dist.world()
        self._emulate = bool(emulate_shard)
        dist.require_world(self.rank, self.world, self._emulate)
        assert shard_mode in ("windows", "points")
        self.shard_mode = shard_mode
        # fixed-base tables for the four key vectors (zkb_msm_table_create): ~14x the key's memory, ~20 % faster MSMs.
        # Default: on, unless ZKB_MSM_TABLES=0.
        self.tables = (os.environ.get("ZKB_MSM_TABLES", "1") != "0") if tables is None else bool(tables)
        self.curve_name = curve
        self.curve = _CURVES[curve]
        self.ec = _EC[self.curve]
        self.poly = _POLY[self.curve]
        self.order = self.ec.ORDER
        assert r1cs.A is not None, "R1CS is not compiled"
        self.r1cs = r1cs
        self.n_public = r1cs.n_public
        # QAP.from_r1cs (qap.py:32-40) rounds the row count up to the NTT domain, in place
        n = next_power_of_two(r1cs.A.n_row)
        r1cs.A.n_row = r1cs.B.n_row = r1cs.C.n_row = n
        self.n = n
        self.log_n = n.bit_length() - 1
        self.m = r1cs.A.n_col
        self.proving_key = None
        self.verifying_key = None
        self._pk_handle = None
        self._r1cs_handle = None
        self._bound_key = None
        self._staging = None
        self.phase_ms = {}
        self._kw_windows = None
        self._staging_ptr = None

    # ------------------------------------------------------------------------------------------------ setup
    def setup(self):
        """protocol.py:32-113 with every batch_mul on the GPU fixed-base kernel and the keys left resident."""
        nat.ensure_init()
        o = self.order
        ec = self.ec
        G1, G2 = ec.g1(), ec.g2()
        # toxic waste: locals only, as in the reference (protocol.py:38-43) -- never stored on the object.  All ranks need the
        # SAME values: rank 0 draws, everybody receives (dist.shared_draws)
        tau, alpha, beta, gamma, delta = self._draws(o - 1, 5)
        inv_gamma, inv_delta = pow(gamma, -1, o), pow(delta, -1, o)
        n, m = self.n, self.m
        # L_i(tau) on the device: one inverse transform of the powers of tau (polynomial.rs:646-652)
        lagrange = FrVec.powers(self.curve, n, tau).intt()
        # L = A^T lambda, R = B^T lambda, O = C^T lambda (protocol.py:64-77) as a device SpMV over the transposed matrices, then
        # K = beta L + alpha R + O -- no Python loop over the triplets
        csr_t = [arr.to_csr_transposed(m) for arr in (self.r1cs.A, self.r1cs.B, self.r1cs.C)]
        vp3 = ctypes.c_void_p * 3
        ht = ctypes.c_void_p()
        nat.check(nat.lib.zkb_r1cs_create(self.curve, m, n, vp3(*[c[0].ctypes.data for c in csr_t]),
                                          vp3(*[c[1].ctypes.data if len(c[1]) else None for c in csr_t]),
                                          vp3(*[c[2].ctypes.data if len(c[2]) else None for c in csr_t]), ctypes.byref(ht)))
        L, R, O = (FrVec(self.curve, m) for _ in range(3))
        nat.check(nat.lib.zkb_r1cs_eval_dev(ht, lagrange.ptr, m, L.ptr, R.ptr, O.ptr))
        K = L.axpy(beta, R.axpy(alpha, O))
        nat.check(nat.lib.zkb_sync())
        nat.lib.zkb_r1cs_free(ht)
        del csr_t, L, R, O, lagrange
        t = self.poly.evaluate_vanishing_polynomial(n, tau)

        by_points = self.shard_mode == "points"
        lo, hi = dist.shard_range(n, self.rank, self.world) if by_points else (0, n)
        self._slice = (lo, hi)

        def powers(scale):  # scale * tau^i for i in this rank's slice
            d = nat.DeviceBuffer(max(hi - lo, 1) * 32)
            first = scale * pow(tau, lo, o) % o
            nat.check(nat.lib.zkb_fr_powers_dev(self.curve, nat.ptr(nat.ints_to_limbs([tau])), nat.ptr(nat.ints_to_limbs([first])),
                                                hi - lo, d.ptr))
            return d

        def batch(base, group, d_scalars, count):
            bases = ec.upload_points([base], group)
            out = ec.PointVector(self.curve, group, count)
            nat.check(nat.lib.zkb_batch_mul_dev(self.curve, group, bases.ptr, 1, d_scalars.ptr, count, out.ptr))
            nat.check(nat.lib.zkb_sync())
            return out

        d_pow = powers(1)
        tau_G1 = batch(G1, 1, d_pow, hi - lo)
        tau_G2 = batch(G2, 2, d_pow, hi - lo)
        d_tgt = powers(t * inv_delta % o)
        target_G1 = batch(G1, 1, d_tgt, hi - lo)
        n_priv = m - self.n_public
        klo, khi = dist.shard_range(n_priv, self.rank, self.world) if by_points else (0, n_priv)
        self._kslice = (klo, khi)
        k_scaled = K.scale(inv_delta)                          # K_j / delta for the private columns of this rank's slice
        d_k = k_scaled.copy(self.n_public + klo, self.n_public + khi, n=max(khi - klo, 1))
        k_delta_G1 = batch(G1, 1, d_k, khi - klo)
        k_gamma_G1 = [G1 * (k * inv_gamma % o) for k in K.to_ints(self.n_public)]
        for d in (d_pow, d_tgt):
            d.free()
        del k_scaled, d_k, K
        self.proving_key = ProvingKey(G1 * alpha, G1 * beta, G2 * beta, G1 * delta, G2 * delta, tau_G1, tau_G2, target_G1,
                                      k_delta_G1)
        self.verifying_key = VerifyingKey(G1 * alpha, G2 * beta, G2 * gamma, G2 * delta, k_gamma_G1)
        self._bind()

    def _draws(self, n_max, count):
        if self.world == 1 or self._emulate:
            return [get_random_int(n_max) for _ in range(count)]
        return dist.shared_draws(get_random_int, n_max, count, self.world)

    def _ensure_bound(self):
        """A key assigned from outside (`prover.proving_key = ProvingKey.from_bytes(...)`, the reference's way of reusing a
        stored key) is bound to the device prover on first use."""
        if self._bound_key is self.proving_key and self._pk_handle is not None:
            return
        pk = self.proving_key
        assert pk is not None, "ProvingKey has not been generated"
        assert self.world == 1 or self.shard_mode == "windows", "a loaded key holds whole vectors: use shard_mode='windows'"
        assert len(pk.tau_1) == self.n and len(pk.tau_2) == self.n and len(pk.target_1) == self.n, \
            "ProvingKey does not match the constraint system"
        assert len(pk.kdelta_1) == self.m - self.n_public, "Length of kdelta_1 and private_witness must be equal"
        self._slice = (0, self.n)
        self._kslice = (0, self.m - self.n_public)
        self._bind()

    def _bind(self):
        """Create the device-side prover objects (proving-key handle + CSR R1CS); whatever was bound before (an earlier setup(),
        another key) is released first -- the handle owns the work buffer, the fixed-base tables and the device CSR."""
        self._release()
        pk = self.proving_key
        flat = lambda pt: np.frombuffer(pt._flat(), dtype=np.uint64).copy()  # noqa: E731
        singles = [flat(pk.alpha_1), flat(pk.beta_1), flat(pk.beta_2), flat(pk.delta_1), flat(pk.delta_2)]
        h = ctypes.c_void_p()
        lo, hi = self._slice
        klo, khi = self._kslice
        nat.check(nat.lib.zkb_groth16_pk_create_sharded(self.curve, self.log_n, pk.tau_1.ptr, pk.tau_2.ptr, pk.target_1.ptr, lo,
                                                        hi - lo, pk.kdelta_1.ptr, self.m - self.n_public, klo, khi - klo,
                                                        *[nat.ptr(s) for s in singles], ctypes.byref(h)))
        self._pk_handle = h
        self._kw_windows = None
        self._singles = singles
        if self.shard_mode == "windows" and self.world > 1:
            nat.check(nat.lib.zkb_groth16_pk_set_window_shard(h, self.rank, self.world))
        if self.tables:
            nat.check(nat.lib.zkb_groth16_pk_build_tables(h, self.world if self.shard_mode == "windows" else 1))
        n_rows = max((int(arr._arrays()[0].max()) for arr in (self.r1cs.A, self.r1cs.B, self.r1cs.C) if arr.triplets), default=-1) + 1
        csr = [arr.to_csr(n_rows) for arr in (self.r1cs.A, self.r1cs.B, self.r1cs.C)]
        self._csr = csr
        vp3 = ctypes.c_void_p * 3
        rp = vp3(*[c[0].ctypes.data for c in csr])
        col = vp3(*[c[1].ctypes.data if len(c[1]) else None for c in csr])
        val = vp3(*[c[2].ctypes.data if len(c[2]) else None for c in csr])
        hr = ctypes.c_void_p()
        nat.check(nat.lib.zkb_r1cs_create(self.curve, n_rows, self.m, rp, col, val, ctypes.byref(hr)))
        self._r1cs_handle = hr
        self._bound_key = pk

    # ------------------------------------------------------------------------------------------------ prove
    def prove(self, public_witness: list, private_witness: list) -> Proof:
        """protocol.py:115-165."""
        assert self.proving_key, "ProvingKey has not been generated"
        assert self.m - self.n_public == len(private_witness), \
            "Length of kdelta_1 and private_witness must be equal"
        r, s = self._draws(self.order - 1, 2)
        # list[int] -> limbs straight into a pinned staging buffer (csrc/pymarshal.cpp: host threads over the PyLong digits; the
        # reference pays one BigUint conversion per element under the GIL, src/bn254/polynomial.rs:537-540).  Values in
        # [0, 2^256) go up as they are and are reduced mod r on the device; negatives (legal here: SparseArray.dot reduces at
        # the end, array.py:43) and wider ints are reduced on the host.
        w = self._witness_staging()
        n_pub = len(public_witness)
        if n_pub + len(private_witness) != self.m:
            raise ValueError(f"witness has {n_pub + len(private_witness)} entries, the constraint system has {self.m} columns")
        nat.ints_to_limbs(public_witness, 32, modulus=self.order, out=w[:n_pub])
        nat.ints_to_limbs(private_witness, 32, modulus=self.order, out=w[n_pub:])
        try:
            return self.prove_packed(w, r, s)
        except ValueError as exc:
            raise ValueError("Failed to evaluate with the given witness") from exc

    def _witness_staging(self):
        """(m, 4) uint64 view of a pinned host buffer owned by this prover (full-rate H2D, no per-proof allocation)."""
        if self._staging is None:
            p = ctypes.c_void_p()
            nat.check(nat.lib.zkb_host_alloc(max(self.m, 1) * 32, ctypes.byref(p)))
            self._staging_ptr = p
            self._staging = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(self.m, 4))
        return self._staging

    def prove_packed(self, witness_limbs, r, s):
        """witness as a (m, 4) uint64 array (host; pinned for full H2D speed) or a DeviceBuffer holding the same bytes -- the
        zero-marshalling entry used by bench.py."""
        self._ensure_bound()
        g1b = nat.lib.zkb_affine_bytes(self.curve, 1) // 8
        g2b = nat.lib.zkb_affine_bytes(self.curve, 2) // 8
        oa, ob, oc = np.zeros(g1b, np.uint64), np.zeros(g2b, np.uint64), np.zeros(g1b, np.uint64)
        inf = (ctypes.c_int * 3)()
        rr, ss = nat.ints_to_limbs([r]), nat.ints_to_limbs([s])
        on_dev = isinstance(witness_limbs, nat.DeviceBuffer)
        wptr = witness_limbs.ptr if on_dev else nat.ptr(witness_limbs)
        if self.world == 1:
            fn = nat.lib.zkb_groth16_prove_witness_dev if on_dev else nat.lib.zkb_groth16_prove_witness
            nat.check(fn(self._pk_handle, self._r1cs_handle, wptr, self.n_public, nat.ptr(rr), nat.ptr(ss), nat.ptr(oa),
                         nat.ptr(ob), nat.ptr(oc), inf))
        else:
            keep = None
            t_ph = [time.perf_counter()]
            if not on_dev and dist.nccl_ready() and os.environ.get("ZKB_WITNESS_SHARDED", "1") != "0":
                # each rank uploads 1/world of the witness; NVLink all-gather instead of world x the same PCIe transfer
                keep = dist.upload_sharded(witness_limbs)
                wptr, on_dev = keep.data_ptr(), True
            # every rank: witness polynomials + the five MSMs over its key slice; one all-gather; identical assembly everywhere
            nat.check(nat.lib.zkb_groth16_precompute(self._pk_handle, nat.ptr(rr), nat.ptr(ss)))   # host threads, under the GPU work
            xy = np.zeros((dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
            flags = np.zeros(dist.MSM_SLOTS, dtype=np.int32)
            spread = self._spread_chains()
            self._set_kw_windows(spread)
            if spread:
                # the three transform chains on three ranks (two: 2 + 1), the quotient's last step on one; dist.exchange_chains
                coeffs, evals, hbuf = dist.spread_buffers(self.n)
                pc, pe, ph = (ctypes.c_void_p(t.data_ptr()) for t in (coeffs, evals, hbuf))
                dist.spread_trace_mark("start")
                nat.check(nat.lib.zkb_groth16_spread_begin(self._pk_handle, self._r1cs_handle, wptr, int(on_dev), self.n_public,
                                                           dist.chain_mask(self.rank, self.world), pc, pe))
                h_ready = dist.exchange_chains(
                    coeffs, evals, hbuf, self.n,
                    lambda: nat.check(nat.lib.zkb_groth16_spread_quotient(self._pk_handle, pe, ph)))
                nat.check(nat.lib.zkb_groth16_spread_finish(self._pk_handle, self._r1cs_handle, self.n_public, pc, ph,
                                                            ctypes.c_void_p(h_ready), nat.ptr(xy), nat.ptr(flags)))
                dist.spread_trace_mark("end")
                dist.spread_trace_collect()
            else:
                nat.check(nat.lib.zkb_groth16_partial(self._pk_handle, self._r1cs_handle, wptr, int(on_dev), self.n_public,
                                                      nat.ptr(xy), nat.ptr(flags)))
            t_ph.append(time.perf_counter())
            all_xy, all_inf = dist.all_gather_partials(xy, flags)
            t_ph.append(time.perf_counter())
            all_xy = np.ascontiguousarray(all_xy)
            all_inf = np.ascontiguousarray(all_inf, dtype=np.int32)
            nat.check(nat.lib.zkb_groth16_assemble_partials(self._pk_handle, self.world, nat.ptr(all_xy), nat.ptr(all_inf), nat.ptr(rr),
                                                            nat.ptr(ss), nat.ptr(oa), nat.ptr(ob), nat.ptr(oc), inf))
            t_ph.append(time.perf_counter())
            # host-side phases of the sharded proof (ms): this rank's work up to its partial sums | waiting for the slowest rank in
            # the exchange | the assembly after it; bench.py reports their means as `host_phases_ms`
            for k, name in enumerate(("partial", "exchange", "assemble")):
                self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (t_ph[k + 1] - t_ph[k]) * 1e3
            self.phase_ms["proofs"] = self.phase_ms.get("proofs", 0) + 1
        ec = self.ec
        return Proof(ec.PointG1._from_flat(oa, inf[0]), ec.PointG2._from_flat(ob, inf[1]), ec.PointG1._from_flat(oc, inf[2]))

    def _set_kw_windows(self, spread):
        """With the chains spread, the ranks that transform nothing run more windows of the [K w] MSM than the ones that do
        (dist.kw_windows); otherwise every MSM is cut evenly.  Needs the fixed-base tables (they fix the window count)."""
        # Measured on the B200 boxes (profiles/R3b_*, R3c_*): 2 GPUs 14.45 -> 14.08 ms; 4 GPUs 8.51 -> 8.92 and 8 GPUs 6.57 -> 6.71 --
        # there the ranks without a chain are not idle for as long as the model says, so only two ranks use it (ZKB_KW_UNEVEN=1 / 0
        # forces it on / off).
        mode = os.environ.get("ZKB_KW_UNEVEN", "auto")
        uneven = spread and (mode == "1" or (mode == "auto" and self.world == 2))
        want = None
        if uneven and self.tables and self.m - self.n_public > 0:
            wins = ctypes.c_uint32()
            nat.check(nat.lib.zkb_groth16_pk_msm_info(self._pk_handle, 3, None, ctypes.byref(wins)))
            want = dist.kw_windows(self.rank, self.world, wins.value)
        if want != self._kw_windows:
            nat.check(nat.lib.zkb_groth16_pk_set_kw_windows(self._pk_handle, *(want or (0, 0)), int(want is not None)))
            self._kw_windows = want

    def _spread_chains(self):
        """Run the quotient's three transform chains on different ranks?  Needs the NCCL world this prover shards over, whole key
        vectors on every rank (window sharding) and a domain large enough for a broadcast to be cheaper than a transform."""
        if os.environ.get("ZKB_NTT_SPREAD", "1") == "0" or self._emulate or self.shard_mode != "windows":
            return False
        return self.world > 1 and dist.nccl_ready() and self.n >= (1 << 16)

    def last_polys(self):
        """U, V, H coefficient lists of the last prove (n entries each, unstripped) for parity tests."""
        out = []
        for which in range(3):
            buf = np.zeros((self.n, 4), dtype=np.uint64)
            nat.check(nat.lib.zkb_groth16_last_poly(self._pk_handle, which, nat.ptr(buf)))
            out.append(nat.limbs_to_ints(buf))
        return out

    def last_msms(self):
        """The five raw MSM results (A, B1, B2, HZ, sum_delta_witness) of the last prove as points."""
        ec = self.ec
        res = []
        for which in range(5):
            grp = 2 if which == 2 else 1
            buf = np.zeros(nat.lib.zkb_affine_bytes(self.curve, grp) // 8, dtype=np.uint64)
            inf = ctypes.c_int()
            nat.check(nat.lib.zkb_groth16_last_msm(self._pk_handle, which, nat.ptr(buf), ctypes.byref(inf)))
            res.append((ec.PointG2 if grp == 2 else ec.PointG1)._from_flat(buf, inf.value))
        return res

    # ------------------------------------------------------------------------------------------------ verify
    def verify(self, proof: Proof, public_witness: list) -> bool:
        """protocol.py:167-186 (pairing on the host; not a proving-path operation)."""
        assert self.verifying_key, "VerifyingKey has not been generated"
        vk = self.verifying_key
        assert len(vk.ic) == len(public_witness), "Length of IC and public_witness must be equal"
        ec = self.ec
        acc = ec.PointG1.identity()
        for pt, w in zip(vk.ic, public_witness):   # tiny MSM (n_public terms) -- host group code
            acc = acc + pt * (int(w) % self.order)
        return ec.pairing(proof.A, proof.B) == ec.multi_pairing([vk.alpha_1, acc, proof.C],
                                                               [vk.beta_2, vk.gamma_2, vk.delta_2])

    def _release(self):
        if self._r1cs_handle:
            nat.lib.zkb_r1cs_free(self._r1cs_handle)
        if self._pk_handle:
            nat.lib.zkb_groth16_pk_free(self._pk_handle)
        self._r1cs_handle = self._pk_handle = None

    def __del__(self):
        try:
            self._release()
            if self._staging_ptr is not None:
                self._staging = None
                nat.lib.zkb_host_free(self._staging_ptr)
                self._staging_ptr = None
        except Exception:
            pass
