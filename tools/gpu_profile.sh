#!/bin/bash
# Run on the GPU box (through gpurun): bench line, ncu launch list, ncu full capture of our kernels.
# Usage: tools/gpu_profile.sh <tag> [workload]
set -u
TAG=${1:-r01}
WL=${2:-nusc_pillar_mean_canvas_b8}
OUT=gpurun_out
KPS=5    # kernels per step
mkdir -p $OUT
python bench.py --workload $WL > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
tail -c 3000 $OUT/bench_${TAG}.json
# launch list: eager single-stream launches so every kernel is its own record; skip the 4 set-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^kf?_ -s $((4*KPS)) -c $((8*KPS)) --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --workload $WL --steps 4 --warmup 3 --no-graph --streams 1 \
    --no-cpu-baseline > $OUT/ncu_launches_${TAG}.log 2>&1
# full capture of one steady-state step
ncu --set full --clock-control none --import-source on -k regex:^kf?_ -s $((4*KPS)) -c $KPS -f -o $OUT/prof_${TAG} \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-graph --streams 1 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
ls $OUT | grep ${TAG}
