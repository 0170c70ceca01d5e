#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py tests/test_gpu_replay.py -q --tb=short > gpurun_out/run4_a.log 2>&1; echo "layers+replay rc=$?"; tail -15 gpurun_out/run4_a.log
timeout 900 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "c2_shape or atari_network or learner_steps" > gpurun_out/run4_learner.log 2>&1; echo "learner rc=$?"; tail -15 gpurun_out/run4_learner.log
timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run4_bench_bf16.json 2> gpurun_out/run4_bench_bf16.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/run4_bench_bf16.json; tail -3 gpurun_out/run4_bench_bf16.err
timeout 300 python tools/step_phases.py bf16 > gpurun_out/run4_phases.log 2>&1; echo "phases rc=$?"; tail -9 gpurun_out/run4_phases.log
