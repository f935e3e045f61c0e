#!/bin/bash
# GPU run 3 (1 GPU): schedule 2 vs schedule 1 -- tests, bench A/B, timelines
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02c_pytest.log
Q="--steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager"
python bench.py $Q > $O/r02c_bench_s2.json 2> $O/r02c_bench_s2.err; echo "s2 rc=$?"; tail -c 300 $O/r02c_bench_s2.err
python bench.py $Q --tunable sched_v1=1 > $O/r02c_bench_s1.json 2> $O/r02c_bench_s1.err; echo "s1 rc=$?"
python bench.py $Q > $O/r02c_bench_s2b.json 2>/dev/null
python bench.py $Q --tunable sched_v1=1 > $O/r02c_bench_s1b.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02c_bench20.json 2>/dev/null
python tools/timeline.py --steps 2 --pipelined 200 --json $O/r02c_timeline_s2.json > $O/r02c_timeline_s2.txt 2>&1; echo "tl2 rc=$?"
python tools/timeline.py --steps 2 --pipelined 200 --tunable sched_v1=1 --json $O/r02c_timeline_s1.json > $O/r02c_timeline_s1.txt 2>&1; echo "tl1 rc=$?"
python bench.py --workload dropin --steps 200 --warmup 10 > $O/r02c_dropin.json 2> $O/r02c_dropin.err; echo "dropin rc=$?"; tail -c 300 $O/r02c_dropin.err
python - <<'PY'
import json
for f in ("r02c_bench_s2","r02c_bench_s1","r02c_bench_s2b","r02c_bench_s1b","r02c_bench20"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); r=d["roofline"]
        print(f, "ms/step %.4f value %.1f | sched %s pool %.1f us frac %.3f | step frac %.3f | parity %s | launches %d | trace %s" % (d["ms_per_step"], d["value"], r.get("schedule"), r["kernel_us"], r["frac"], r["step"]["frac"], (d.get("parity") or {}).get("ok"), d["gpu_launches"], r["device_trace_us"]))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("gpurun_out/r02c_dropin.json")); print("dropin ms %.4f eager ms %.4f floor %.4f parity %s launches %d" % (d["ms_per_step"], d["gpu_eager_baseline"]["ms_per_step"], d["inline_torch_floor"]["ms_per_step"], d["parity"]["ok"], d["gpu_launches"]))
except Exception as e: print("dropin ERR", e)
PY
cat $O/r02c_timeline_s2.txt | tail -30
cat $O/r02c_timeline_s1.txt | tail -16
