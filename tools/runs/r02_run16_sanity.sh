#!/bin/bash
# GPU run 16 (1 GPU): last sanity check of the committed tree -- full GPU tests, smoke, the driver's bench invocation
set -u
O=gpurun_out
python -m pytest tests -q -m gpu > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02q_pytest.log
python __graft_entry__.py smoke > $O/r02q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02q_smoke.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > $O/r02q_bench20.json 2> $O/r02q_bench20.err; echo "bench20 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02q_bench20.json')); print('bench20', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'], d['roofline']['step']['frac'])"
