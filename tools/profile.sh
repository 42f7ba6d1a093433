#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench, ncu launch list of the same command, and `--set full` captures of the top kernels.
# usage: tools/profile.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r1}; shift || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
for K in "$@"; do
  NAME=$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o $OUT/prof_${TAG}_$NAME $CMD > $OUT/ncu_full_${TAG}_$NAME.log 2>&1
  echo "full $K rc=$?"
done
ls -la $OUT
