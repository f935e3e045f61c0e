#!/bin/bash
# GPU run 11 (1 GPU): schedule 3 ([pool(xs) | mc_stats] in one launch) -- equality test, A/B vs schedules 1 / 2, pool share sweep
set -u
O=gpurun_out
python -m pytest tests/test_gpu_step.py -q -m gpu -x -k "schedule or config1 or graph" > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02k_pytest.log
Q="--steps 800 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager"
for cfg in "--tunable sched=1" "--tunable sched=2 --tunable dfin_split=2" "--tunable sched=3 --tunable dfin_split=2" "--tunable sched=3 --tunable dfin_split=2 --tunable pool_pct=35" "--tunable sched=3 --tunable dfin_split=2 --tunable pool_pct=40" "--tunable sched=3 --tunable dfin_split=2 --tunable pool_pct=50" "--tunable sched=3 --tunable dfin_split=2 --tunable pool_pct=55" "--tunable sched=3 --tunable dfin_split=1" "--tunable sched=1"; do
  python bench.py $Q $cfg > $O/r02k_tmp.json 2>$O/r02k_tmp.err
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/r02k_tmp.json")); t=d["roofline"]["device_trace_us"]
    print("AB [%s] ms/step %.4f parity %s | pool %.1f pool_t %s fin_s %s mc %.1f retr %.1f cons %.1f disc %.1f span %.1f" % (sys.argv[1], d["ms_per_step"], (d.get("parity") or {}).get("ok"), t.get("pool_fwd",0), t.get("pool_fwd_target"), t.get("finish_source"), t.get("mc_stats",0), t.get("retrify_weights",0), t.get("cons_fwd",0), t.get("disc_fused",0), t.get("step_span",0)))
except Exception as e:
    print("AB [%s] ERR %s" % (sys.argv[1], e)); print(open("gpurun_out/r02k_tmp.err").read()[-600:])
PY
done
python tools/timeline.py --steps 2 --tunable sched=3 --tunable dfin_split=2 > $O/r02k_timeline_s3.txt 2>&1; tail -14 $O/r02k_timeline_s3.txt
