"""Multi-GPU split of the proving path: one process per GPU, contiguous slices of every MSM's point and
scalar range per rank, and ONE small collective per proof that exchanges the per-rank partial sums.

The reference has no multi-device path; its MSM (/root/reference/src/bn254/curve.rs:356-392) is a sum of independent terms, so
the slice sums add up to the same group element and the proof bytes do not depend on the world size.

`torch.distributed` is only the plumbing (NCCL on GPUs, gloo in the CPU tests): the payload is 5 affine points + 5 flags per
rank (< 1 KiB).
"""
import ctypes
import os

import numpy as np

from . import _native as nat

MSM_SLOTS = 5            # A, B1, B2 (G2), HZ, KW  -- protocol.py:133-155
SLOT_LIMBS = 24          # uint64 per slot in the C ABI (room for a BLS12-381 G2 affine point)
SLOT_GROUP = (1, 1, 2, 1, 1)


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` items owned by `rank` (sizes differ by at most one item)."""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi


def world():
    """(rank, world_size) of the current torch.distributed job, (0, 1) when not initialised."""
    try:
        import torch.distributed as td
    except ImportError:
        return 0, 1
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


# host<->device bytes moved by this module through torch (the library counts its own copies: zkb_transfer_count)
TRANSFER = {"h2d": 0, "d2h": 0}


def nccl_ready():
    try:
        import torch.distributed as td
        return td.is_available() and td.is_initialized() and td.get_backend() == "nccl"
    except Exception:
        return False


_lib_stream = None
_upload_bufs = {}


def library_stream():
    """torch view of the library's CUDA stream (zkb_stream): torch copies and NCCL collectives issued under
    `torch.cuda.stream(library_stream())` are ORDERED with the library's kernels -- no host synchronisation between a
    collective and the kernels that consume it (ProcessGroupNCCL makes the current stream wait for its own NCCL stream)."""
    global _lib_stream
    import torch
    if _lib_stream is None:
        nat.ensure_init()
        _lib_stream = torch.cuda.ExternalStream(nat.lib.zkb_stream())
    return _lib_stream


def upload_sharded(arr):
    """A host array that every rank holds (the witness) -> a device copy on every rank, moving only 1/world of it over each
    rank's PCIe link: every rank uploads its own slice and one NCCL all-gather over NVLink completes the vector.  With N
    ranks uploading the whole 32 MiB witness at once the host side is the bottleneck (measured: +1.0 ms at 4 GPUs against
    +0.4 ms at 1).  Both steps are enqueued on the library stream and the buffers persist across proofs: the call returns
    without synchronising.  Returns a torch uint8 CUDA tensor (valid until the next call with the same size)."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    flat = np.ascontiguousarray(arr).reshape(-1).view(np.uint8)
    nbytes = flat.size
    per = -(-nbytes // (ws * 256)) * 256                      # slice size, 256-byte aligned
    bufs = _upload_bufs.get((nbytes, ws))
    if bufs is None:     # one pair per distinct size, kept for the life of the process (the library stream may still be reading them)
        with torch.cuda.stream(library_stream()):
            bufs = (torch.empty(per * ws, dtype=torch.uint8, device="cuda"), torch.empty(per, dtype=torch.uint8, device="cuda"))
        _upload_bufs[(nbytes, ws)] = bufs
    full, mine = bufs
    lo = min(rank * per, nbytes)
    hi = min(lo + per, nbytes)
    with torch.cuda.stream(library_stream()):
        if hi > lo:
            # (the source is cudaHostAlloc'ed by the caller when it wants full PCIe rate; the driver recognises it by address)
            # counted by the library (zkb_transfer_count), not in TRANSFER
            nat.check(nat.lib.zkb_h2d_async(ctypes.c_void_p(mine.data_ptr()), ctypes.c_void_p(flat[lo:hi].ctypes.data), hi - lo))
        td.all_gather_into_tensor(full, mine)
    return full


# ---- the three interpolation -> coset-evaluation chains of the Groth16 quotient, one per rank ---------------------------------------
# QAP.evaluate_witness (/root/reference/python/zksnake/groth16/qap.py:57-69) transforms A.w, B.w and C.w independently and then forms
# H from the three results.  On one GPU the chains are one batched launch per pass; on several, rank chain_owner(c) runs chain c alone
# (zkb_groth16_spread_begin), U and V -- the MSM scalars of every rank's window shard -- are broadcast over NVLink on the library
# stream, the three coset evaluation vectors go point to point to ONE rank, quotient_owner(), which forms H (one fused kernel + one
# inverse transform, zkb_groth16_spread_quotient) and broadcasts it.  On every other rank that last exchange runs on a second stream
# beside the MSMs over U, V and the witness; only the [H Z] MSM, the last of the batch, waits for it (zkb_groth16_spread_finish).
# Per rank: 2 transforms (chain owners), 1 (the quotient rank) or none, instead of 7; a rank without a chain runs its [K w] MSM while
# the owners transform.  Two ranks: rank 0 takes the U and W chains, rank 1 the V chain and the quotient (4 + 3 transforms).
_spread_bufs = {}
_spread_state = {}
# ZKB_SPREAD_TRACE=1: device timestamps of the exchange steps (ms after the proof's first kernel: own chains done | U, V received |
# evaluation vectors received | quotient formed | H received | last kernel), kept in SPREAD_TRACE and reported by bench.py
SPREAD_TRACE = {} if os.environ.get("ZKB_SPREAD_TRACE", "0") != "0" else None


def spread_trace_mark(key):
    """Record trace event `key` ("start" / "end") on the library stream (no-op unless tracing)."""
    if SPREAD_TRACE is not None and _spread_state:
        _spread_state[key].record(library_stream())


def spread_trace_collect():
    """After a proof has completed: add its event times to SPREAD_TRACE (sums; "proofs" counts them)."""
    st = _spread_state
    if SPREAD_TRACE is None or "start" not in st:
        return
    rank, ws = world()
    keys = ["chains", "uv", "h", "end"] + (["recv", "quot"] if rank == quotient_owner(ws) else [])
    try:
        st["end"].synchronize()
        st["h"].synchronize()
        for k in keys:
            SPREAD_TRACE[k] = SPREAD_TRACE.get(k, 0.0) + st["start"].elapsed_time(st[k])
        SPREAD_TRACE["proofs"] = SPREAD_TRACE.get("proofs", 0) + 1
    except Exception:       # (the first proof creates the events after "start" would have been recorded)
        pass


def chain_owner(chain, world_size):
    """Rank that runs chain 0 (U), 1 (V) or 2 (W).  Three or more ranks: one chain each; two ranks: rank 1 runs all three (and the
    quotient) while rank 0 runs most of the [K w] MSM, which needs the witness only (kw_windows)."""
    return chain % world_size if world_size >= 3 else world_size - 1


def chain_mask(rank, world_size):
    return sum(1 << c for c in range(3) if chain_owner(c, world_size) == rank)


def quotient_owner(world_size):
    """Rank that forms H from the three coset evaluation vectors: the first rank without a chain (four ranks or more), else the
    last chain owner (nothing to send on two ranks)."""
    return 3 if world_size >= 4 else world_size - 1


# Rough device times of one proof's stages in units of ONE window of the [K w] MSM (2^20 BN254: ~0.2 ms; all of them scale with the
# domain size together, so only the ratios matter): one interpolation -> coset-evaluation chain, the wait for U and V behind it, the
# quotient's last step, one evaluation vector received for it.
# Fitted on the B200 boxes: 2 GPUs -- 14 / 2 windows left rank 1 0.4 ms early (profiles/R3b_n2_spread1.json); 4 GPUs -- the even
# 4 / 4 / 4 / 4 is balanced to 0.1 ms and a fifth window on the quotient rank makes it the slowest (R2z_n4_spread1, R3c_n4).
_CHAIN_COST, _UV_WAIT_COST, _QUOTIENT_COST, _RECV_COST = 2.6, 1.0, 1.6, 0.5


def kw_windows(rank, world_size, n_windows):
    """(first, count): the windows of the [K w] MSM that `rank` runs when the transform chains are spread over the ranks.  That MSM
    needs the witness only, so it is what a rank does while it has nothing else: the windows are dealt one by one to the rank
    whose modelled busy time (chains, the wait for U and V, the quotient) is lowest, and laid out as contiguous ranges in rank
    order.  Every rank computes the same table; the ranges partition [0, n_windows)."""
    busy = []
    for r in range(world_size):
        chains = bin(chain_mask(r, world_size)).count("1")
        b = chains * _CHAIN_COST + (_UV_WAIT_COST if chains else 0.0)
        if r == quotient_owner(world_size) and world_size > 1:
            b += _QUOTIENT_COST + _RECV_COST * (3 - chains)
        busy.append(b)
    count = [0] * world_size
    for _ in range(n_windows):
        r = min(range(world_size), key=lambda k: (busy[k] + count[k], k))
        count[r] += 1
    return sum(count[:rank]), count[rank]


def spread_buffers(n_elems):
    """Persistent device buffers as torch uint8 tensors: coefficients U|V|W (3 n Fr elements), coset evaluations (3 n), H (n)."""
    import torch
    bufs = _spread_bufs.get(n_elems)
    if bufs is None:
        with torch.cuda.stream(library_stream()):
            bufs = (torch.empty(3 * n_elems * 32, dtype=torch.uint8, device="cuda"),
                    torch.empty(3 * n_elems * 32, dtype=torch.uint8, device="cuda"),
                    torch.empty(n_elems * 32, dtype=torch.uint8, device="cuda"))
        _spread_bufs[n_elems] = bufs
    return bufs


def exchange_chains(coeffs, evals, hbuf, n_elems, run_quotient):
    """The exchange between zkb_groth16_spread_begin and _finish (see the comment above).  run_quotient() is called on the quotient
    rank once its evaluation vectors are ordered on the library stream.  Returns the raw cudaEvent_t (int) behind which `hbuf` holds
    H, or None where H is ordered on the library stream itself.  No host synchronisation anywhere."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    nb = n_elems * 32
    lib = library_stream()
    h = quotient_owner(ws)
    st = _spread_state
    if not st:
        trace = SPREAD_TRACE is not None
        st["comm"] = torch.cuda.Stream()
        st["chains"] = torch.cuda.Event(enable_timing=trace)
        st["h"] = torch.cuda.Event(enable_timing=trace)
        if trace:
            for k in ("start", "uv", "recv", "quot", "end"):
                st[k] = torch.cuda.Event(enable_timing=True)
    mark = (lambda k: st[k].record(lib)) if SPREAD_TRACE is not None else (lambda k: None)
    st["chains"].record(lib)                       # this rank's chains are complete here (before the broadcasts queue up)
    with torch.cuda.stream(lib):
        for c in (0, 1):
            td.broadcast(coeffs[c * nb:(c + 1) * nb], src=chain_owner(c, ws))
    mark("uv")
    senders = [(c, chain_owner(c, ws)) for c in range(3) if chain_owner(c, ws) != h]
    # the evaluation vectors travel as ONE batch of point-to-point operations per rank (unbatched send / recv calls are serialised
    # with everything else on the communicator: three receives took 0.86 ms on 8 x B200, profiles/R3d_n8_trace.json)
    if rank == h:
        with torch.cuda.stream(lib):
            if senders:
                for work in td.batch_isend_irecv([td.P2POp(td.irecv, evals[c * nb:(c + 1) * nb], o) for c, o in senders]):
                    work.wait()
            mark("recv")
            try:
                run_quotient()
            finally:
                mark("quot")
                td.broadcast(hbuf, src=h)          # (issued even if the quotient failed: the other ranks are already waiting in it)
        st["h"].record(lib)
        return None
    comm = st["comm"]
    with torch.cuda.stream(comm):
        comm.wait_event(st["chains"])
        mine = [td.P2POp(td.isend, evals[c * nb:(c + 1) * nb], h) for c, o in senders if o == rank]
        if mine:
            for work in td.batch_isend_irecv(mine):
                work.wait()
        td.broadcast(hbuf, src=h)
        st["h"].record(comm)
    return st["h"].cuda_event


# ---- host-side exchange of small payloads between the ranks of ONE node -----------------------------------------------------
class HostExchange:
    """All-gather of a few hundred bytes per rank through a POSIX shared-memory mailbox.

    The per-rank partial MSM results are BORN ON THE HOST: the per-window sums of every MSM come back from the GPU and are
    recombined by host threads (csrc/host_math.cpp: a serial chain of ~100 group operations is ~100x faster on a CPU core than in
    one GPU thread).  Sending those 5 points through NCCL means host -> device -> NVLink -> device -> host with two stream
    synchronisations (round 1: ~0.3 ms of the 8 ms proof at 8 GPUs); all ranks of the bench are processes of one node
    (`torchrun --nnodes=1`), so they are exchanged where they are: every rank owns one slot of a shared segment, publishes
    (payload, sequence number) and spins until every other slot shows the same sequence number -- a few microseconds.  Two
    alternating buffers per slot make the protocol safe without a second barrier (a rank can be at most one round ahead).
    NCCL carries the data-path traffic (the witness all-gather); multi-node jobs fall back to the torch collective."""
    SLOT = 2048

    def __init__(self):
        import atexit
        from multiprocessing import shared_memory
        import torch.distributed as td
        self.rank, self.ws = world()
        name = [None]
        if self.rank == 0:
            name[0] = f"zkb200_{os.getpid()}_{int.from_bytes(os.urandom(4), 'little')}"
            self.shm = shared_memory.SharedMemory(name=name[0], create=True, size=self.ws * 2 * (self.SLOT + 64))
            self.shm.buf[:] = bytes(len(self.shm.buf))
        td.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        td.barrier()
        self._owner = self.rank == 0
        stride = self.SLOT + 64
        raw = np.frombuffer(self.shm.buf, dtype=np.uint8, count=self.ws * 2 * stride)
        self._seq = [[raw[(r * 2 + b) * stride:(r * 2 + b) * stride + 8].view(np.uint64) for b in range(2)] for r in range(self.ws)]
        self._data = [[raw[(r * 2 + b) * stride + 64:(r * 2 + b + 1) * stride] for b in range(2)] for r in range(self.ws)]
        self._raw = raw
        self.round = 0
        atexit.register(self.close)

    def close(self):
        shm, self.shm = getattr(self, "shm", None), None
        if shm is None:
            return
        self._seq = self._data = self._raw = None
        try:
            shm.close()
            if self._owner:
                shm.unlink()
        except Exception:
            pass

    def all_gather(self, payload):
        """payload: uint8 array of at most SLOT bytes, same length on every rank -> (world, len) uint8 array"""
        import time
        n = payload.size
        assert n <= self.SLOT
        self.round += 1
        b = self.round & 1
        self._data[self.rank][b][:n] = payload
        self._seq[self.rank][b][0] = self.round            # published after the payload (x86 keeps the store order)
        out = np.empty((self.ws, n), dtype=np.uint8)
        deadline = time.monotonic() + 120.0
        for r in range(self.ws):
            seq = self._seq[r][b]
            spins = 0
            while seq[0] != self.round:
                spins += 1
                os.sched_yield()       # (the ranks share the host's cores: let a rank that still enqueues kernels have this one)
                if spins & 0xFFFF == 0 and time.monotonic() > deadline:
                    raise RuntimeError(f"rank {self.rank}: no partial sums from rank {r} after 120 s")
            out[r] = self._data[r][b][:n]
        return out


_host_exchange = None


def host_exchange():
    """The node-local mailbox, created on first use (None when the job spans several nodes or shared memory is unavailable)."""
    global _host_exchange
    if _host_exchange is None:
        local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0)
        rank, ws = world()
        if os.environ.get("ZKB_HOST_EXCHANGE", "1") == "0" or (local and local != ws):
            _host_exchange = False
        else:
            try:
                _host_exchange = HostExchange()
            except Exception:
                _host_exchange = False
    return _host_exchange or None


def all_gather_partials(msm_xy, msm_inf, device=None):
    """Exchange the per-rank partial MSM results.  msm_xy: (5, 24) uint64, msm_inf: (5,) int32.
    Returns (world, 5, 24) uint64 and (world, 5) int32 arrays, identical on every rank."""
    rank, ws = world()
    if ws == 1:
        return msm_xy[None].copy(), msm_inf[None].copy()
    payload = np.zeros(MSM_SLOTS * SLOT_LIMBS + MSM_SLOTS, dtype=np.int64)
    payload[:MSM_SLOTS * SLOT_LIMBS] = msm_xy.reshape(-1).view(np.int64)
    payload[MSM_SLOTS * SLOT_LIMBS:] = msm_inf
    hx = host_exchange()
    if hx is not None:
        arr = hx.all_gather(payload.view(np.uint8)).view(np.int64).reshape(ws, -1)
    else:
        import torch
        import torch.distributed as td
        t = torch.from_numpy(payload)
        if td.get_backend() == "nccl":
            t = t.cuda(device) if device is not None else t.cuda()
            TRANSFER["h2d"] += payload.nbytes
            TRANSFER["d2h"] += ws * payload.nbytes
        out = torch.empty(ws * t.numel(), dtype=torch.int64, device=t.device)
        td.all_gather_into_tensor(out, t)
        arr = out.cpu().numpy().reshape(ws, -1)
    xy = arr[:, :MSM_SLOTS * SLOT_LIMBS].copy().view(np.uint64).reshape(ws, MSM_SLOTS, SLOT_LIMBS)
    inf = arr[:, MSM_SLOTS * SLOT_LIMBS:].astype(np.int32)
    return xy, inf


def add_partials(curve, all_xy, all_inf):
    """Sum the ranks' partial points slot by slot on the host (zkb_point_lincomb: exact group law, no GPU needed).
    Returns (5, 24) uint64 and (5,) int32."""
    ws = all_xy.shape[0]
    out_xy = np.zeros((MSM_SLOTS, SLOT_LIMBS), dtype=np.uint64)
    out_inf = np.zeros(MSM_SLOTS, dtype=np.int32)
    for slot, grp in enumerate(SLOT_GROUP):
        limbs = nat.lib.zkb_affine_bytes(curve, grp) // 8
        pts = np.ascontiguousarray(all_xy[:, slot, :limbs])
        infs = np.ascontiguousarray((all_inf[:, slot] & 1).astype(np.int32))      # (bit 1 of the HZ slot: zkb200.h, "folded")
        scal = np.zeros((ws, 4), dtype=np.uint64)
        has = np.zeros(ws, dtype=np.int32)
        res = np.zeros(limbs, dtype=np.uint64)
        inf = ctypes.c_int()
        nat.check(nat.lib.zkb_point_lincomb(curve, grp, ws, nat.ptr(pts), nat.ptr(infs), nat.ptr(scal), nat.ptr(has),
                                            nat.ptr(res), ctypes.byref(inf)))
        out_xy[slot, :limbs] = res
        out_inf[slot] = inf.value
    folded = int(np.count_nonzero(all_inf[:, 3] & 2))
    if folded not in (0, ws):
        raise ValueError("partial sums of some ranks carry the folded s [U] + r [V] term and others do not")
    if folded:
        out_inf[3] |= 2
    return out_xy, out_inf


def all_gather_array(arr, device=None):
    """all-gather of one small numpy array per rank (same shape and dtype everywhere) -> array of shape (world,) + arr.shape.
    NCCL on GPUs (the payload is staged through a CUDA tensor), gloo on CPU; identity when not distributed."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    if ws == 1:
        return arr[None].copy()
    flat = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    if _host_exchange and flat.size <= HostExchange.SLOT:     # small host payloads go through the node-local mailbox once it exists
        return _host_exchange.all_gather(flat).view(arr.dtype).reshape((ws,) + arr.shape)
    t = torch.from_numpy(flat.copy())
    if td.get_backend() == "nccl":
        t = t.cuda(device) if device is not None else t.cuda()
    out = torch.empty(ws * t.numel(), dtype=torch.uint8, device=t.device)
    td.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(arr.dtype).reshape((ws,) + arr.shape)


def require_world(rank, world_size, emulate=False):
    """An explicit shard=(rank, world) must BE the torch.distributed world: the cooperative provers draw shared randomness and
    exchange partial sums with collectives, and without a process group of that size every rank would silently produce a proof
    from its own slice and its own randomness.  `emulate=True` (tests that drive zkb_groth16_partial / assemble rank by rank in one
    process) skips the check; the collectives below then refuse to run."""
    if world_size == 1 or emulate:
        return
    if world() != (rank, world_size):
        raise RuntimeError(f"shard=({rank}, {world_size}) but the torch.distributed world is {world()}: initialise a process "
                           "group of that size (torchrun) before constructing a sharded prover")


def shared_draws(draw, n_max, count, world_size=None):
    """`count` values of the randomness hook `draw(n_max)` (uniform in [1, n_max]) that every rank agrees on: rank 0 draws
    them, one small all-gather hands them to everybody.  Toxic waste, prover randomness and blinding scalars must be identical
    on every rank, otherwise the partial sums belong to different keys / proofs.  Because the values ARE rank 0's draws, a
    seeded hook gives the same key and the same proof bytes at every world size (bench.py's `proof_sha`)."""
    rank, ws = world()
    if world_size is not None and world_size > 1 and ws != world_size:
        raise RuntimeError(f"shared randomness for {world_size} ranks needs a torch.distributed world of that size (have {ws})")
    mine = [int(draw(n_max)) for _ in range(count)] if (rank == 0 or ws == 1) else [0] * count
    if ws == 1:
        return mine
    nbytes = (max(int(n_max).bit_length(), 1) + 7) // 8
    raw = np.frombuffer(b"".join(v.to_bytes(nbytes, "little") for v in mine), dtype=np.uint8).copy()
    got = bytes(all_gather_array(raw)[0])
    return [int.from_bytes(got[i * nbytes:(i + 1) * nbytes], "little") for i in range(count)]


def shared_random(draw):
    """Function form of shared_draws for call sites that draw one value at a time (one collective per draw)."""
    return lambda n_max: shared_draws(draw, n_max, 1)[0]
