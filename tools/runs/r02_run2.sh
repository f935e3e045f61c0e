#!/bin/bash
# GPU run 2 of round 2: regressions fixed (warp-cooperative guard band, pooling kernel restored): tests, bench, drop-in workload, probe
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02b_pytest.log
python tools/aten_order_probe.py --seeds 50 > $O/r02b_mask_flips.json 2> $O/r02b_probe.err; echo "probe rc=$?"; tail -c 300 $O/r02b_probe.err
python bench.py --steps 20 --warmup 5 > $O/r02b_bench20.json 2> $O/r02b_bench20.err; echo "bench20 rc=$?"; tail -c 300 $O/r02b_bench20.err
python bench.py > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench rc=$?"
python bench.py --workload dropin --steps 200 --warmup 10 > $O/r02b_dropin.json 2> $O/r02b_dropin.err; echo "dropin rc=$?"; tail -c 300 $O/r02b_dropin.err
python bench.py --workload align --steps 500 --no-cpu-baseline > $O/r02b_align.json 2> $O/r02b_align.err; echo "align rc=$?"
python bench.py --C 305 --steps 500 --no-cpu-baseline --no-e2e > $O/r02b_c305.json 2> /dev/null; echo "c305 rc=$?"
python - <<'PY'
import json
for f in ("r02b_bench20","r02b_bench","r02b_align","r02b_c305"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); r=d["roofline"]
        print(f, "ms/step %.4f value %.1f | pool %.1f us frac %.3f | step frac %.3f | parity %s | trace %s" % (d["ms_per_step"], d["value"], r["kernel_us"], r["frac"], r["step"]["frac"], (d.get("parity") or {}).get("ok"), r["device_trace_us"]))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("gpurun_out/r02b_dropin.json")); print("dropin ms %.4f eager ms %.4f parity %s launches %d" % (d["ms_per_step"], d["gpu_eager_baseline"]["ms_per_step"], d["parity"], d["gpu_launches"]))
except Exception as e: print("dropin ERR", e)
PY
