"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol include/zkb200.h declares, refuses compute
without a device (no CPU fallback), and the multi-rank host logic (slicing, the gloo all-gather of partial MSM sums, their
exact addition) is correct at world_size 2."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    syms = header_symbols()
    assert len(syms) >= 50
    lib = ctypes.CDLL(native.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table binds exactly the header's functions
    assert sorted(native.EXPORTED) == syms


def test_no_cpu_fallback(native):
    if native.gpu_available():
        pytest.skip("a GPU is visible")
    out = np.zeros((4, 4), dtype=np.uint64)
    rc = native.lib.zkb_ntt(0, 0, 0, 2, native.ptr(out), 4, native.ptr(out))
    assert rc == native.ERR_NOINIT and "no CPU fallback" in native.last_error()
    with pytest.raises(native.ZkbError):
        native.ensure_init()


def test_oracle_is_not_imported_by_the_product():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zksnake_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libzkcpu" not in src, f


def test_shard_range_partitions():
    from zksnake_b200.dist import shard_range
    for total in (0, 1, 5, 1 << 20, (1 << 20) + 1):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys, random
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as td
from zksnake_b200 import dist
from zksnake_b200 import _native as nat
from oracle import cport
from oracle.fields import PARAMS
td.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.world()
assert world == 2
for curve in (0, 1):
    r = PARAMS[curve].r
    n = 257
    rnd = random.Random(5)
    scal = cport.pack([rnd.randrange(r) for _ in range(n)])
    lo, hi = dist.shard_range(n, rank, world)
    xy = np.zeros((dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
    inf = np.zeros(dist.MSM_SLOTS, dtype=np.int32)
    full = []
    for slot, grp in enumerate(dist.SLOT_GROUP):
        pts = cport.chain_points(curve, grp, 3 + slot, n)
        if slot == 4:                       # one rank contributes the identity
            part, pinf = (np.zeros(cport.affine_limbs(curve, grp), np.uint64), True) if rank == 1 else cport.msm(curve, grp, pts, scal)
            want = cport.msm(curve, grp, pts, scal)
        else:
            part, pinf = cport.msm(curve, grp, pts[lo:hi], scal[lo:hi])   # stands in for the GPU's partial MSM
            want = cport.msm(curve, grp, pts, scal)
        xy[slot, :len(part)] = part
        inf[slot] = int(pinf)
        full.append(want)
    all_xy, all_inf = dist.all_gather_partials(xy, inf)
    assert all_xy.shape == (2, 5, 24) and (all_xy[rank] == xy).all() and (all_inf[rank] == inf).all()
    sxy, sinf = dist.add_partials(curve, all_xy, all_inf)
    for slot, grp in enumerate(dist.SLOT_GROUP):
        limbs = cport.affine_limbs(curve, grp)
        assert (sxy[slot, :limbs] == full[slot][0]).all() and bool(sinf[slot]) == full[slot][1], (curve, slot)
td.barrier()
td.destroy_process_group()
print("rank", rank, "ok")
"""


def test_partial_msm_exchange_world_size_2_gloo(tmp_path):
    """Two gloo ranks: slice the MSM range, all-gather the partial sums, add them with the host group law (zkb_point_lincomb,
    no GPU) -- must equal the unsliced MSM.  The CPU oracle stands in for the GPU's partial MSM here (tests only)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def _gather_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import numpy as np
    import torch.distributed as td
    td.init_process_group("gloo", rank=rank, world_size=world)
    from zksnake_b200 import dist
    arr = np.arange(6, dtype=np.uint64).reshape(2, 3) + 100 * rank
    got = dist.all_gather_array(arr)
    import random
    own = random.Random(100 + rank)                          # each rank would draw something different on its own
    hook = lambda n_max: own.randint(1, n_max)               # noqa: E731
    vals = dist.shared_draws(hook, 10 ** 30, 3, 2) + [dist.shared_random(hook)(10 ** 30)]
    try:
        dist.require_world(0, 3)
        vals.append("no error")
    except RuntimeError:
        pass
    try:
        dist.shared_draws(hook, 10, 1, world_size=4)
        vals.append("no error")
    except RuntimeError:
        pass
    # the node-local shared-memory mailbox that carries the partial MSM sums (several rounds: the two buffers alternate)
    hx = dist.host_exchange()
    rounds = []
    for k in range(5):
        xy = (np.arange(dist.MSM_SLOTS * dist.SLOT_LIMBS, dtype=np.uint64).reshape(dist.MSM_SLOTS, dist.SLOT_LIMBS) + 1000 * rank + k)
        inf = np.array([rank, k, 0, 1, rank ^ 1], dtype=np.int32)
        all_xy, all_inf = dist.all_gather_partials(xy, inf)
        rounds.append((all_xy[:, 0, 0].tolist(), all_xy[:, 4, 23].tolist(), all_inf.tolist()))
    vals.append(("hx", hx is not None, rounds))
    if hx is not None:
        hx.close()
    q.put((rank, got.tolist(), vals))
    td.destroy_process_group()


def test_all_gather_array_and_shared_draws_gloo():
    """world-size-2 gloo run of the helpers the multi-GPU provers rely on: every rank sees every rank's array, and the shared
    random stream is identical everywhere (it is derived from rank 0's draw)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29631
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    want = [[[0, 1, 2], [3, 4, 5]], [[100, 101, 102], [103, 104, 105]]]
    assert res[0][1] == want and res[1][1] == want
    import random
    rank0 = random.Random(100)
    assert res[0][2][:4] == res[1][2][:4] == [rank0.randint(1, 10 ** 30) for _ in range(4)]    # the values ARE rank 0's draws
    for r in (0, 1):
        tag, have, rounds = res[r][2][4]
        assert tag == "hx" and have
        for k, (first, last, inf) in enumerate(rounds):
            assert first == [k, 1000 + k] and last == [4 * 24 + 23 + k, 1000 + 4 * 24 + 23 + k]
            assert inf == [[0, k, 0, 1, 1], [1, k, 0, 1, 0]]


def test_sharded_prover_needs_a_matching_process_group():
    """an explicit shard=(rank, world) without a torch.distributed world of that size must raise, not prove from one
    slice with private randomness."""
    from zksnake_b200 import dist
    dist.require_world(0, 1)
    dist.require_world(1, 4, emulate=True)
    with pytest.raises(RuntimeError, match="torch.distributed world"):
        dist.require_world(1, 4)
    with pytest.raises(RuntimeError):
        dist.shared_draws(lambda n: 1, 10, 2, world_size=2)


def test_chain_spreading_assignment():
    """Every transform chain of the Groth16 quotient has exactly one owner at every world size, the masks partition {U, V, W},
    and with three or more ranks no rank runs more than one chain (zksnake_b200/dist.py)."""
    from zksnake_b200 import dist
    for ws in range(1, 10):
        owners = [dist.chain_owner(c, ws) for c in range(3)]
        assert all(0 <= o < ws for o in owners)
        masks = [dist.chain_mask(r, ws) for r in range(ws)]
        assert sum(masks) == 7 and all(masks[a] & masks[b] == 0 for a in range(ws) for b in range(a))
        for c in range(3):
            assert masks[owners[c]] >> c & 1
        if ws >= 3:
            assert all(bin(m).count("1") <= 1 for m in masks)
        if ws >= 2:
            # the rank that forms H: a valid rank, without a chain where one exists, never the rank with the most chains; and the
            # point-to-point schedule of dist.exchange_chains pairs every send with one receive in the same order on both sides
            h = dist.quotient_owner(ws)
            assert 0 <= h < ws
            assert masks[h] == 0 if ws >= 4 else masks[h] != 0         # (two ranks: rank 1 transforms everything, rank 0 nothing)
            senders = [(c, owners[c]) for c in range(3) if owners[c] != h]
            assert len(senders) == 3 - bin(masks[h]).count("1")
            for r in range(ws):
                if r != h:
                    assert [c for c, o in senders if o == r] == [c for c in range(3) if masks[r] >> c & 1]
        # the [K w] windows: contiguous ranges in rank order that partition [0, W), and a rank that transforms never gets more
        # windows than one that does not
        for n_windows in (1, 7, 13, 14, 16, 17):
            ranges = [dist.kw_windows(r, ws, n_windows) for r in range(ws)]
            at = 0
            for first, count in ranges:
                assert first == at and count >= 0
                at += count
            assert at == n_windows
            if ws >= 2:
                idle = [ranges[r][1] for r in range(ws) if masks[r] == 0 and r != dist.quotient_owner(ws)]
                busy = [ranges[r][1] for r in range(ws) if masks[r] != 0]
                if idle and busy:
                    assert min(idle) >= max(busy)


def test_add_partials_masks_and_carries_the_folded_marker(native):
    """Bit 1 of the HZ slot's flag says that a rank folded s [U]_i + r [V]_i into that slot (include/zkb200.h).  The Python slot
    sums (host group law, no GPU) must treat it as a marker, not as "infinity", hand it on when EVERY rank set it and refuse a mix."""
    import random

    import numpy as np

    from oracle import cport
    from zksnake_b200 import dist
    curve, ws = 0, 3
    rnd = random.Random(9)
    n = 8
    xy = np.zeros((ws, dist.MSM_SLOTS, dist.SLOT_LIMBS), dtype=np.uint64)
    inf = np.zeros((ws, dist.MSM_SLOTS), dtype=np.int32)
    want = []
    for slot, grp in enumerate(dist.SLOT_GROUP):
        pts = cport.chain_points(curve, grp, 5 + slot, n * ws)
        scal = cport.pack([rnd.randrange(1, 1 << 64) for _ in range(n * ws)])
        for k in range(ws):
            part, pinf = cport.msm(curve, grp, pts[k * n:(k + 1) * n], scal[k * n:(k + 1) * n])
            xy[k, slot, :len(part)] = part
            inf[k, slot] = int(pinf)
        want.append(cport.msm(curve, grp, pts, scal))
    plain_xy, plain_inf = dist.add_partials(curve, xy, inf)
    for slot, grp in enumerate(dist.SLOT_GROUP):
        limbs = cport.affine_limbs(curve, grp)
        assert (plain_xy[slot, :limbs] == want[slot][0]).all() and int(plain_inf[slot]) == int(want[slot][1])
    folded = inf.copy()
    folded[:, 3] |= 2
    f_xy, f_inf = dist.add_partials(curve, xy, folded)
    assert (f_xy == plain_xy).all() and int(f_inf[3]) == (int(plain_inf[3]) | 2)
    assert [int(v) for k, v in enumerate(f_inf) if k != 3] == [int(v) for k, v in enumerate(plain_inf) if k != 3]
    mixed = folded.copy()
    mixed[1, 3] &= 1
    with pytest.raises(ValueError, match="folded"):
        dist.add_partials(curve, xy, mixed)
