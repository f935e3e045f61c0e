#!/bin/bash
# GPU run 1 of round 2: ATen-order probe, GPU tests, smoke, bench (driver settings + default), L2-harvest A/B
set -u
O=gpurun_out
mkdir -p $O
python tools/aten_order_probe.py --seeds 10 > $O/r02a_aten_probe.json 2> $O/r02a_aten_probe.err; echo "probe rc=$?"; tail -c 800 $O/r02a_aten_probe.err
python -m pytest tests -q -m gpu > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02a_pytest.log
python __graft_entry__.py smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r02a_smoke.log
python bench.py --steps 20 --warmup 5 > $O/r02a_bench20.json 2> $O/r02a_bench20.err; echo "bench20 rc=$?"; tail -c 600 $O/r02a_bench20.err
python bench.py > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench rc=$?"; tail -c 600 $O/r02a_bench.err
# L2 harvest A/B (live step time) ...
Q="--steps 600 --warmup 20 --no-cpu-baseline --no-e2e --no-parity --no-gpu-eager"
for cfg in "" "--tunable disc_reverse=1" "--tunable pool_order=1" "--tunable pool_order=1 --tunable disc_reverse=1" "--tunable pool_order=1 --tunable disc_reverse=1 --tunable cons_ef=1" "--tunable cons_ef=1"; do
  tag=$(echo "base $cfg" | tr -d ' =-' | sed 's/tunable/_/g')
  python bench.py $Q $cfg > $O/r02a_ab_$tag.json 2>/dev/null
  python - "$O/r02a_ab_$tag.json" "$cfg" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); t=d["roofline"]["device_trace_us"]
print("AB [%s] ms/step %.4f  pool %.1f cons %.1f disc %.1f bwd %s span %.1f" % (sys.argv[2], d["ms_per_step"], t.get("pool_fwd",0), t.get("cons_fwd",0), t.get("disc_fused",0), t.get("pool_bwd_target"), t.get("step_span",0)))
PY
done
# ... and DRAM bytes of the discriminative / pooling kernels with the caches left as the previous kernel left them
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-gpu-eager"
i=0
for cfg in "" "--tunable disc_reverse=1" "--tunable pool_order=1" "--tunable pool_order=1 --tunable disc_reverse=1" "--tunable pool_order=1 --tunable disc_reverse=1 --tunable cons_ef=1"; do
  i=$((i+1))
  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      -k regex:'disc_fused|pool_fwd_ldg|pool_finish_cons' -c 18 --csv --log-file $O/r02a_l2_$i.csv python bench.py $S $cfg > $O/r02a_l2_$i.log 2>&1
  echo "ncu [$cfg] rc=$?"
  python - "$O/r02a_l2_$i.csv" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
acc=collections.OrderedDict()
for r in rows[1:]:
    acc.setdefault((r[ii], r[ki][:28]), {})[r[mi]] = r[vi]
for (i,k),m in list(acc.items())[-6:]:
    print("   ", i, k, {a.split("__")[-1][:24]: b for a,b in m.items()})
PY
done
ls $O | grep r02a | head -40
