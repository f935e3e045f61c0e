#!/bin/bash
# GPU run 8 (N GPUs): schedule A/B at world N, two repetitions each, interleaved
set -u
N=${1:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
Q="--gpus $N --steps 1500 --warmup 30 --no-e2e --no-parity"
i=0
for cfg in "" "--tunable sched=2 --tunable dfin_split=1" "" "--tunable sched=2 --tunable dfin_split=1"; do
  i=$((i+1))
  timeout 600 $TR --master-port 2977$i bench.py $Q $cfg 2> $O/r02h_bench_n${N}_$i.err | grep '^{' > $O/r02h_bench_n${N}_$i.json
  python - "$O/r02h_bench_n${N}_$i.json" "$cfg" $N <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("N=%s [%s] ms/step %.4f value %.1f" % (sys.argv[3], sys.argv[2], d["ms_per_step"], d["value"]))
except Exception as e: print("  ERR", e)
PY
done
