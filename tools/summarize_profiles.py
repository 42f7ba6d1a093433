"""Turns the ncu exports that tools/profile.sh left in gpurun_out/ into the small tracked files under profiles/:
  <tag>_launches.csv (per-kernel totals of the launch-list pass), <tag>_<kernel>.txt (key counters of the --set full capture),
  traffic.json (per-launch DRAM bytes of the dominant kernels, read by bench.py).    usage: summarize_profiles.py <tag>"""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def read_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.reader(io.StringIO("".join(lines))))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(tag):
    os.makedirs(PROF, exist_ok=True)
    # launch list -> per-kernel totals
    rows = read_csv(os.path.join(OUT, f"launches_{tag}.csv"))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v for _, v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
        f.write("# (setup + 1 warm-up pair + 1 timed device-resident prove + 1 timed e2e prove; cold-cache, serialised)\n")
        f.write("kernel,launches,total_us,share\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k[:110]}\",{c},{v / 1e3:.1f},{v / total:.4f}\n")
    try:        # entries of other kernels / earlier captures stay (bench.py reads all of them)
        with open(os.path.join(PROF, "traffic.json")) as f:
            traffic = json.load(f)
    except (OSError, ValueError):
        traffic = {}
    for fn in sorted(os.listdir(OUT)):
        if not (fn.startswith(f"full_{tag}_") and fn.endswith(".csv")):
            continue
        name = fn[len(f"full_{tag}_"):-4].strip("_")
        rows = read_csv(os.path.join(OUT, fn))
        if len(rows) < 3:
            continue
        h, u, v = rows[0], rows[1], rows[2]
        col = {x: i for i, x in enumerate(h)}
        with open(os.path.join(PROF, f"{tag}_{name}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{name} -s 6 -c 1  (bench.py --steps 1 --warmup 1)\n")
            f.write(f"kernel: {v[col['Kernel Name']]}\ngrid: {v[col['Grid Size']]}  block: {v[col['Block Size']]}\n")
            for k in KEYS:
                if k in col:
                    f.write(f"{k:80s} {v[col[k]]:>18s} {u[col[k]]}\n")
        if "dram__bytes_read.sum" in col:
            rd = to_bytes(v[col["dram__bytes_read.sum"]], u[col["dram__bytes_read.sum"]])
            wr = to_bytes(v[col["dram__bytes_write.sum"]], u[col["dram__bytes_write.sum"]])
            key = {"msm_accumulate_kernel": "msm_accumulate_g1", "ntt_pass_kernel": "ntt_pass"}.get(name, name)
            traffic[key] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "tag": tag, "kernel": v[col["Kernel Name"]][:80],
                            "duration_us": float(v[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u[col["gpu__time_duration.sum"]], 1)}
    with open(os.path.join(PROF, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    # top stall lines of the source page (SASS-level) for the first kernel, if exported
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main(sys.argv[1])
