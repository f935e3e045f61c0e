#!/bin/bash
# GPU run 5 (N GPUs, default 8): sharded parity at world N, bench for the schedule / split variants, per-rank timelines
set -u
N=${1:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 tools/dist_check.py > $O/r02e_dist_check_world$N.txt 2>&1; echo "dist_check rc=$?"; grep -c " ok" $O/r02e_dist_check_world$N.txt; grep -c FAIL $O/r02e_dist_check_world$N.txt; tail -1 $O/r02e_dist_check_world$N.txt
timeout 600 $TR --master-port 29562 tests/tools/ddp_check.py > $O/r02e_ddp_check_world$N.txt 2>&1; echo "ddp_check rc=$?"; grep "rank 0" $O/r02e_ddp_check_world$N.txt | tail -4; tail -1 $O/r02e_ddp_check_world$N.txt
Q="--gpus $N --steps 1000 --warmup 20 --no-e2e --no-parity"
i=0
for cfg in "--tunable sched=1 --tunable dfin_split=2" "--tunable sched=1 --tunable dfin_split=1" "--tunable sched=2 --tunable dfin_split=2" "--tunable sched=2 --tunable dfin_split=1" ""; do
  i=$((i+1))
  timeout 600 $TR --master-port 2957$i bench.py $Q $cfg 2> $O/r02e_bench_n${N}_$i.err | grep '^{' > $O/r02e_bench_n${N}_$i.json; echo "bench [$cfg] rc=$?"
  python - "$O/r02e_bench_n${N}_$i.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("  [%s] ms/step %.4f value %.1f timed %s" % (sys.argv[2], d["ms_per_step"], d["value"], d["timed_region_check"]))
except Exception as e: print("  ERR", e)
PY
done
# the default configuration once more WITH the parity block (what the driver runs)
timeout 900 $TR --master-port 29579 bench.py --gpus $N --steps 200 --warmup 10 2> $O/r02e_bench_n${N}_default.err | grep '^{' > $O/r02e_bench_n${N}_default.json; echo "bench default rc=$?"
python - "$O/r02e_bench_n${N}_default.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); p=d["parity"]; print("  default ms/step %.4f value %.1f e2e %s parity ok %s cross %s" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value"), p["ok"], p["cross_rank"]))
    print("  ", json.dumps(p["steps"])[:700])
except Exception as e: print("  ERR", e)
PY
for s in 1 2; do
  timeout 600 $TR --master-port 2958$s tools/timeline.py --steps 3 --all-ranks --tunable sched=$s --json $O/r02e_timeline_n${N}_sched$s.json > $O/r02e_timeline_n${N}_sched$s.txt 2>&1; echo "timeline sched=$s rc=$?"
done
tail -14 $O/r02e_timeline_n${N}_sched1.txt
ls $O | grep r02e | wc -l
