#!/bin/bash
# SASS evidence of the built objects (uda_clr_b200/build/*.o): which memory / sync / packed-math instructions the kernels use.
# Usage: bash tools/sass_evidence.sh > profiles/<tag>_sass_evidence.txt     (run after python -m uda_clr_b200.build)
cd "$(dirname "$0")/.."
echo "# SASS evidence (cuobjdump -sass of the built objects, sm_100a; $(nvcc --version | grep release | sed 's/.*release //'))"
for f in disc_fused pool_fwd pool_bwd mc_stats cons transnorm; do
  echo "$f.o:"
  cuobjdump -sass uda_clr_b200/build/$f.o | grep -o "UTMALDG[.0-9A-Z]*\|UBLKCP[.A-Z]*\|LDGSTS[.A-Z0-9]*\|LDG\.E\.[A-Z0-9.]*128[.A-Z]*\|LDG\.E\.[A-Z0-9.]*256[.A-Z]*\|STG\.E\.[A-Z0-9.]*128\|SYNCS\.[A-Z0-9.]*\|FFMA2\|MUFU\.[A-Z0-9]*\|SHFL\.[A-Z]*" | sort | uniq -c | awk '{printf "  %6d %s\n", $1, $2}'
done
echo "any kernel: griddepcontrol (PDL; ACQBULK / launch_dependents):"
for f in uda_clr_b200/build/*.o; do
  n=$(cuobjdump -sass $f | grep -c "ACQBULK\|PREEXIT")
  echo "$(basename $f): $n"
done
