#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/h_timeline.py > gpurun_out/run10_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/run10_timeline.log | tail -24
timeout 600 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "ddpg or d4pg" > gpurun_out/run10_ddpg.log 2>&1; echo "ddpg rc=$?"; tail -8 gpurun_out/run10_ddpg.log
