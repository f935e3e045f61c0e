#!/bin/bash
# Run on the GPU box (via gpurun): full validation of the tree + bench lines + sweep + ncu launch list and one full capture of
# the streaming kernels, everything under gpurun_out/<tag>_*.  (The captures are taken with --no-parity --no-gpu-eager so that
# the launch list holds this library's kernels, not the eager port's.)  Afterwards, in the dev container:
#   python tools/ncu_extract.py gpurun_out/prof_<tag>.ncu-rep profiles/<tag>; python tools/launch_summary.py gpurun_out/<tag>_launches.csv
# Usage: bash tools/gpu_profile.sh <tag>
exec bash "$(dirname "$0")/runs/r02_run9_final.sh" "${1:-r02}"
