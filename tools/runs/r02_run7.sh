#!/bin/bash
# GPU run 7 (1 GPU): disc_ctas=3 A/B, guard-band retrify timing, mc_accumulate A/B, tests
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02g_pytest.log
Q="--steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-eager --no-parity"
for cfg in "" "--tunable disc_ctas=3" "" "--tunable disc_ctas=3" "--tunable disc_ctas=3 --tunable disc_reverse=1"; do
  python bench.py $Q $cfg > $O/r02g_tmp.json 2>/dev/null
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02g_tmp.json")); t=d["roofline"]["device_trace_us"]
print("AB [%s] ms/step %.4f  pool %.1f cons %.1f disc %.1f retr %.1f mc %.1f span %.1f" % (sys.argv[1], d["ms_per_step"], t.get("pool_fwd",0), t.get("cons_fwd",0), t.get("disc_fused",0), t.get("retrify_weights",0), t.get("mc_stats",0), t.get("step_span",0)))
PY
done
python bench.py --C 305 $Q > $O/r02g_tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02g_tmp.json')); print('C305 base', d['ms_per_step'], d['roofline']['device_trace_us'].get('disc_fused'))"
python bench.py --C 305 $Q --tunable disc_ctas=3 > $O/r02g_tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02g_tmp.json')); print('C305 ctas3', d['ms_per_step'], d['roofline']['device_trace_us'].get('disc_fused'))"
python tools/mc_accumulate_ab.py > $O/r02g_mc_accumulate.json 2> $O/r02g_mc_accumulate.err; echo "mc_acc rc=$?"; cat $O/r02g_mc_accumulate.json | head -30
