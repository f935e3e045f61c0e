#!/bin/bash
# GPU run 10 (1 GPU): packed-FFMA2 pooling A/B for K = 4 / 8 (and K = 3), plus the K > 2 tests
set -u
O=gpurun_out
python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -q -m gpu -x > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02j_pytest.log
Q="--steps 400 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager"
for K in 4 8 3; do
 for cfg in "" "--tunable pool_pair=2" "" "--tunable pool_pair=2"; do
  python bench.py $Q --K $K $cfg > $O/r02j_tmp.json 2>/dev/null
  python - "$K" "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02j_tmp.json")); t=d["roofline"]["device_trace_us"]
print("K=%s [%s] ms/step %.4f  pool %.1f disc %.1f bwd_t %s bwd_s %s mc %.1f parity %s" % (sys.argv[1], sys.argv[2], d["ms_per_step"], t.get("pool_fwd",0), t.get("disc_fused",0), t.get("pool_bwd_target"), t.get("pool_bwd_source"), t.get("mc_stats",0), (d.get("parity") or {}).get("ok")))
PY
 done
done
