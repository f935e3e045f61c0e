#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list and one full capture of the top kernels.
# Usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
echo "bench rc=$?"; tail -c 3000 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
# launch list (per-launch device time; cold-cache, serialised: compare shares)
$SMALL > $OUT/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $SMALL > $OUT/ncu_launch_${TAG}.log 2>&1
echo "ncu launch list rc=$?"
# full capture of the streaming kernels (one launch each of the last step)
$SMALL > $OUT/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'pool_fwd_ldg|pool_fwd_tma|pool_bwd_kernel|bwd_finish|disc_fused|mc_stats|pool_finish_cons|retrify' -s 18 -c 6 -o $OUT/prof_${TAG} -f $SMALL > $OUT/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -20
