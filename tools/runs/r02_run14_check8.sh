#!/bin/bash
# GPU run 14 (N GPUs): sharded parity of the FINAL build -- dist_check (K = 2 / 8, NCCL + peer, both schedules) and the real-DDP check
set -u
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29601 tools/dist_check.py > $O/r02o_dist_check_world$N.txt 2>&1; echo "dist_check rc=$?"; grep -c " ok" $O/r02o_dist_check_world$N.txt; grep -c FAIL $O/r02o_dist_check_world$N.txt; tail -1 $O/r02o_dist_check_world$N.txt
timeout 300 $TR --master-port 29602 tests/tools/ddp_check.py > $O/r02o_ddp_check_world$N.txt 2>&1; echo "ddp_check rc=$?"; tail -1 $O/r02o_ddp_check_world$N.txt
