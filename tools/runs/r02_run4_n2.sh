#!/bin/bash
# GPU run 4 (2 GPUs): sharded-step parity (repo single-GPU step and eager port under real DDP), bench N=2 for both schedules
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/r02d_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $O/r02d_pytest_multi.log
$TR --master-port 29541 tools/dist_check.py > $O/r02d_dist_check_world2.txt 2>&1; echo "dist_check rc=$?"; grep -c " ok" $O/r02d_dist_check_world2.txt; tail -2 $O/r02d_dist_check_world2.txt
$TR --master-port 29542 tests/tools/ddp_check.py > $O/r02d_ddp_check_world2.txt 2>&1; echo "ddp_check rc=$?"; tail -6 $O/r02d_ddp_check_world2.txt
Q="--gpus 2 --steps 1000 --warmup 20 --no-e2e"
for s in 0 1 2; do
  $TR --master-port 2955$s bench.py $Q --tunable sched=$s > $O/r02d_bench_n2_sched$s.json 2> $O/r02d_bench_n2_sched$s.err; echo "bench n2 sched=$s rc=$?"
done
python - <<'PY'
import json
for s in (0,1,2):
    try:
        d=json.load(open("gpurun_out/r02d_bench_n2_sched%d.json"%s)); r=d["roofline"]
        print("sched", s, "ms/step %.4f value %.1f | sched %s | parity %s %s | timed check %s" % (d["ms_per_step"], d["value"], r.get("schedule"), (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("cross_rank"), d["timed_region_check"]))
    except Exception as e: print(s, "ERR", e)
PY
tail -c 400 $O/r02d_bench_n2_sched0.err
