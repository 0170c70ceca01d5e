#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay.py -q --tb=short -k "save_restore or lockstep" > gpurun_out/run3_replay.log 2>&1; echo "replay rc=$?"; tail -15 gpurun_out/run3_replay.log
timeout 300 python bench.py --profile --steps 3 --warmup 5 --precision bf16 > gpurun_out/run3_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_r02_bf16_a.csv python bench.py --profile --steps 3 --warmup 5 --precision bf16 > gpurun_out/run3_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/run3_ncu.log
B200RL_STAMPS=1 timeout 300 python tools/step_phases.py bf16 > gpurun_out/run3_phases.log 2>&1; echo "phases rc=$?"; tail -12 gpurun_out/run3_phases.log
