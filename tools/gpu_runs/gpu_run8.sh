#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py -q --tb=short > gpurun_out/run9_layers.log 2>&1; echo "layers rc=$?"; tail -12 gpurun_out/run9_layers.log
timeout 900 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "c2_shape or atari_network or learner_steps" > gpurun_out/run9_learner.log 2>&1; echo "learner rc=$?"; tail -8 gpurun_out/run9_learner.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run9_bench_$name.json 2> gpurun_out/run9_bench_$name.err; echo "bench $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/run9_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,1) for k,v in d['stages_us'].items()})" 2>&1 | tail -1)"; tail -2 gpurun_out/run9_bench_$name.err; }
run pers X=1
run nopers B200RL_PERSISTENT=0
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > gpurun_out/run9_phases.log 2>&1; echo "phases rc=$?"; tail -32 gpurun_out/run9_phases.log
