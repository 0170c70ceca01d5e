#!/bin/bash
# round 2, GPU call 1: full GPU test-suite, smoke, baseline bench, HBM-kernel timings + ncu --set full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/run1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run1_pytest.log
tail -5 gpurun_out/run1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/run1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/run1_smoke.log
timeout 600 python bench.py --steps 500 --warmup 20 > gpurun_out/run1_bench.json 2> gpurun_out/run1_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/run1_bench.json
timeout 300 python tools/ncu_hbm_kernels.py > gpurun_out/run1_hbm.log 2>&1; echo "hbm rc=$?"; tail -30 gpurun_out/run1_hbm.log
timeout 300 python tools/ncu_hbm_kernels.py --once > gpurun_out/run1_hbm_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/hbm_r02 python tools/ncu_hbm_kernels.py --once > gpurun_out/run1_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/run1_ncu.log
ls -la gpurun_out | head -30
