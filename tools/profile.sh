#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench, ncu launch list of the same command, and `--set full` captures of the top kernels.
# usage: tools/profile.sh <tag> [kernel-regex ...]
# Only the CSV / text exports travel back (gpurun_out/ is capped at 64 MiB); the .ncu-rep of the FIRST kernel is kept.
set -u
TAG=${1:-r1}; shift || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
FIRST=1
for K in "$@"; do
  NAME=$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
  REP=$OUT/prof_${TAG}_$NAME
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o $REP $CMD > $OUT/ncu_full_${TAG}_$NAME.log 2>&1
  echo "full $K rc=$?"
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/full_${TAG}_$NAME.csv 2>/dev/null
  ncu -i $REP.ncu-rep --page details > $OUT/details_${TAG}_$NAME.txt 2>/dev/null
  ncu -i $REP.ncu-rep --page source --csv > $OUT/source_${TAG}_$NAME.csv 2>/dev/null
  if [ $FIRST -eq 0 ]; then rm -f $REP.ncu-rep; fi
  FIRST=0
done
ls -la $OUT
