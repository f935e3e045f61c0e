#!/bin/bash
# GPU run 6 (8 GPUs, lean): bench variants after the parallel-poll fix + per-rank timeline of the default schedule
set -u
N=${1:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
Q="--gpus $N --steps 1000 --warmup 20 --no-e2e --no-parity"
i=0
for cfg in "" "--tunable dfin_split=1" "--tunable sched=2 --tunable dfin_split=1" "--exchange nccl"; do
  i=$((i+1))
  timeout 600 $TR --master-port 2967$i bench.py $Q $cfg 2> $O/r02f_bench_n${N}_$i.err | grep '^{' > $O/r02f_bench_n${N}_$i.json; echo "bench [$cfg] rc=$?"
  python - "$O/r02f_bench_n${N}_$i.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("  [%s] ms/step %.4f value %.1f timed %s" % (sys.argv[2], d["ms_per_step"], d["value"], d["timed_region_check"]))
except Exception as e: print("  ERR", e)
PY
done
timeout 600 $TR --master-port 29681 tools/timeline.py --steps 3 --all-ranks --json $O/r02f_timeline_n${N}.json > $O/r02f_timeline_n${N}.txt 2>&1; echo "timeline rc=$?"
tail -12 $O/r02f_timeline_n${N}.txt
python - $N <<'PY'
import json,sys,glob
N=int(sys.argv[1])
for f in sorted(glob.glob("gpurun_out/r02f_timeline_n%d.json.rank*"%N)):
    d=json.load(open(f)); st=d["steps"][-1]; k={x["name"]:x for x in st["kernels"]}
    def busy(n): return (k[n]["exit_us"]-k[n]["ready_us"]) if n in k else float("nan")
    print("rank %d span %.1f | align %.1f cons %.1f disc %.1f dfin %.1f bwd_t %.1f bwd_s %.1f pool %.1f mc %.1f" % (d["rank"], st["span_us"], busy("align_finalize"), busy("cons_fwd"), busy("disc_fused"), busy("disc_finalize"), busy("pool_bwd_target"), busy("pool_bwd_source"), busy("pool_fwd"), busy("mc_stats")))
PY
