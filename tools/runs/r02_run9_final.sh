#!/bin/bash
# GPU run 9 (1 GPU): validation of the final build -- GPU tests, smoke, bench (driver settings + default), sweep, ncu launch list + full capture
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench20.json 2> $O/${TAG}_bench20.err; echo "bench20 rc=$?"
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> /dev/null; echo "reference arm rc=$?"
python bench.py --workload align --steps 1000 --no-cpu-baseline > $O/${TAG}_bench_align.json 2>/dev/null; echo "align rc=$?"
python bench.py --workload dropin --steps 200 --warmup 10 > $O/${TAG}_dropin.json 2>/dev/null; echo "dropin rc=$?"
python tests/perf/sweep.py --json $O/${TAG}_sweep.json > $O/${TAG}_sweep.log 2>&1; echo "sweep rc=$?"
python tools/timeline.py --steps 2 --pipelined 200 --json $O/${TAG}_timeline.json > $O/${TAG}_timeline.txt 2>&1; echo "timeline rc=$?"
python - $TAG <<'PY'
import json,sys
T=sys.argv[1]
for f in ("bench20","bench","bench_align"):
    try:
        d=json.load(open("gpurun_out/%s_%s.json"%(T,f))); r=d["roofline"]
        print(f, "ms/step %.4f value %.1f e2e %s | pool %.1f us frac %.3f bwd frac %.3f | step frac %.3f | parity %s | eager %s | cpu %s" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value"), r["kernel_us"], r["frac"], r["bwd_kernel"]["frac"], r["step"]["frac"], (d.get("parity") or {}).get("ok"), (d.get("gpu_eager_baseline") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("gpurun_out/%s_dropin.json"%T)); print("dropin ms %.4f eager %.4f floor %.4f ops share %.4f" % (d["ms_per_step"], d["gpu_eager_baseline"]["ms_per_step"], d["inline_torch_floor"]["ms_per_step"], d["ops_share_ms"]))
except Exception as e: print("dropin ERR", e)
for r in json.load(open("gpurun_out/%s_sweep.json"%T)):
    print("sweep", {k:r.get(k) for k in ("B","C","H","K","ms_per_step","mpixel_s","gbs","parity_ok","error")})
PY
SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-gpu-eager"
$SMALL > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $SMALL > $O/${TAG}_ncu_launch.log 2>&1
echo "ncu launch list rc=$?"
$SMALL > $O/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'pool_fwd_ldg|pool_bwd_kernel|bwd_finish|disc_fused|mc_stats|pool_finish_cons|retrify' -s 18 -c 6 -o $O/prof_${TAG} -f $SMALL > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $O | grep ${TAG} | awk '{print $5, $9}' | tail -25
