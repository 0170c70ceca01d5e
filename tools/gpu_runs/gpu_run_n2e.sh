#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_sequences.py -q --tb=short > gpurun_out/n2e_seq.log 2>&1; echo "seq tests rc=$?"; tail -40 gpurun_out/n2e_seq.log
B200RL_DP_CE=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2e_phases_ce.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n2e_phases_ce.log | tail -48
