"""One NTT size, one kernel variant, a few repetitions -- the target of `ncu` captures.
  python tools/ntt_one.py <log_n> <variant: smem|el2> [reps=3] [curve=0] [inverse=0] [coset=0]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200 import _native as nat  # noqa: E402

log_n = int(sys.argv[1])
variant = sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
curve = int(sys.argv[4]) if len(sys.argv) > 4 else 0
inverse = int(sys.argv[5]) if len(sys.argv) > 5 else 0
coset = int(sys.argv[6]) if len(sys.argv) > 6 else 0
os.environ["ZKB_NTT_WARP"] = "0" if variant == "smem" else "1"
os.environ["ZKB_NTT_WARP_SINGLE"] = "0" if variant == "smem" else "1"
os.environ["ZKB_NTT_EL"] = variant[2:] if variant.startswith("el") else "0"
nat.ensure_init()
n = 1 << log_n
rng = np.random.default_rng(1)
x = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
x[:, 3] &= np.uint64((1 << 59) - 1)
d_in = nat.DeviceBuffer(n * 32).upload(x)
d_out = nat.DeviceBuffer(n * 32)
ts = []
for _ in range(reps + 1):
    with nat.Timer() as t:
        nat.check(nat.lib.zkb_ntt_dev(curve, inverse, coset, log_n, d_in.ptr, n, d_out.ptr))
    ts.append(t.ms)
print(f"log_n={log_n} {variant}: first {ts[0]:.4f} ms, then {min(ts[1:]):.4f} ms")
