#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 tools/symm_probe.py > gpurun_out/n2b_symm.log 2>&1; echo "symm rc=$?"; grep -v OMP gpurun_out/n2b_symm.log | tail -12
timeout 300 python tools/mc_probe.py 2>&1 | tail -3
B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2b_phases.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n2b_phases.log | tail -50
