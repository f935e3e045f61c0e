#!/bin/bash
# GPU run 12 (1 GPU): tap-row mean map A/B, pipelined e2e, full GPU tests
set -u
O=gpurun_out
python -m pytest tests -q -m gpu -x > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02l_pytest.log
Q="--steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager --no-parity"
for cfg in "" "--tunable mc_all_rows=1" "" "--tunable mc_all_rows=1"; do
  python bench.py $Q $cfg > $O/r02l_tmp.json 2>/dev/null
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02l_tmp.json")); t=d["roofline"]["device_trace_us"]
print("AB [%s] ms/step %.4f | mc %.1f retr %.1f pool %.1f span %.1f" % (sys.argv[1], d["ms_per_step"], t.get("mc_stats",0), t.get("retrify_weights",0), t.get("pool_fwd",0), t.get("step_span",0)))
PY
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02l_bench20.json 2>$O/r02l_bench20.err; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench20.json')); print('bench20', d['ms_per_step'], d['value'], 'e2e', d['e2e'], 'parity', d['parity']['ok'])"
