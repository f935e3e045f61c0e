#!/bin/bash
# GPU run 15 (1 GPU): disc_lag A/B + its equality test
set -u
O=gpurun_out
python -m pytest tests/test_gpu_step.py -q -m gpu -x -k "variants or config1 or disc_fused" > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02p_pytest.log
Q="--steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager --no-parity"
for cfg in "" "--tunable disc_lag=1" "" "--tunable disc_lag=1"; do
  python bench.py $Q $cfg > $O/r02p_tmp.json 2>/dev/null
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02p_tmp.json")); t=d["roofline"]["device_trace_us"]
print("AB [%s] ms/step %.4f | disc %.1f cons %.1f pool %.1f span %.1f" % (sys.argv[1], d["ms_per_step"], t.get("disc_fused",0), t.get("cons_fwd",0), t.get("pool_fwd",0), t.get("step_span",0)))
PY
done
python bench.py --C 305 $Q > $O/r02p_tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02p_tmp.json')); print('C305 base', d['ms_per_step'], d['roofline']['device_trace_us'].get('disc_fused'))"
python bench.py --C 305 $Q --tunable disc_lag=1 > $O/r02p_tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02p_tmp.json')); print('C305 lag', d['ms_per_step'], d['roofline']['device_trace_us'].get('disc_fused'))"
