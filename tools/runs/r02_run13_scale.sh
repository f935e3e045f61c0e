#!/bin/bash
# GPU run 13 (N GPUs): the driver's scaling invocation at world N (default settings, parity block included) + multi-GPU tests at N = 2
set -u
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_multi.py -q -m gpu > $O/r02m_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/r02m_pytest_multi.log; fi
timeout 900 $TR --master-port 29591 bench.py --gpus $N --steps 20 --warmup 5 2> $O/r02m_scale_n$N.err | grep '^{' > $O/r02m_scale_n$N.json; echo "bench rc=$?"
timeout 900 $TR --master-port 29592 bench.py --gpus $N --steps 2000 --warmup 20 --no-e2e 2> /dev/null | grep '^{' > $O/r02m_scale_n${N}_long.json
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("r02m_scale_n%s"%N, "r02m_scale_n%s_long"%N):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); p=d["parity"]
        print(f, "ms/step %.4f value %.1f e2e %s sched %s parity ok %s cross %s timed %s" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value"), d["roofline"].get("schedule"), p["ok"], p["cross_rank"], d["timed_region_check"]))
    except Exception as e: print(f, "ERR", e)
PY
