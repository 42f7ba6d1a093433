#!/bin/bash
# A/B of library variants on one GPU box: tools/ab_variants.sh TAG "bench args" main v0 v2 ...
# Each variant is a prebuilt variants/libzkb200_<v>.so ("main" = the in-tree library); the bench line of every run lands in
# gpurun_out/<TAG>_<variant>_<curve>.json.  Development tool, not part of the product.
set -u
TAG=$1; shift
ARGS=$1; shift
mkdir -p gpurun_out
cp zksnake_b200/libzkb200.so /tmp/libzkb200_main.so
for spec in "$@"; do   # spec = lib[:ENV=VAL[,ENV=VAL...]]
  v=${spec%%:*}
  envs=""; if [ "$spec" != "$v" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  label=$(echo "$spec" | tr -c 'A-Za-z0-9_\n' '_')
  if [ "$v" = main ]; then cp /tmp/libzkb200_main.so zksnake_b200/libzkb200.so; else cp variants/libzkb200_$v.so zksnake_b200/libzkb200.so; fi
  for curve in ${CURVES:-BN254 BLS12_381}; do
    env $envs timeout 600 python bench.py --curve $curve $ARGS > gpurun_out/${TAG}_${label}_${curve}.json 2> gpurun_out/${TAG}_${label}_${curve}.err
    echo "$spec $curve rc=$?"
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${label}_${curve}.json").read().strip().splitlines()[-1])
    g2 = d.get("roofline_g2") or {}
    print("  value", round(d["value"], 3), "parity", d.get("parity"), "g2_launch_ms", g2.get("launch_ms"), "g2_frac", g2.get("frac"),
          "g1_launch_ms", (d.get("roofline") or {}).get("launch_ms"), "breakdown", {k: round(v, 2) for k, v in (d.get("breakdown_ms") or {}).items()})
except Exception as e:
    print("  no line:", e)
PY
  done
done
cp /tmp/libzkb200_main.so zksnake_b200/libzkb200.so
