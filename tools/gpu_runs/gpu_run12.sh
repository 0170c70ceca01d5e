#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/h_timeline.py > gpurun_out/run12_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/run12_timeline.log | tail -20
echo "--- adam IEEE"; timeout 120 python tools/adam_micro.py 2>&1 | tail -7
echo "--- adam fast"; B200RL_ADAM_FAST=1 timeout 120 python tools/adam_micro.py 2>&1 | tail -7
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py tests/test_gpu_learner.py -q --tb=short -k "bf16 or c2_shape or atari_network or learner_steps or adam" > gpurun_out/run12_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/run12_tests.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run12_bench_$name.json 2> gpurun_out/run12_bench_$name.err; echo "bench $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/run12_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,1) for k,v in d['stages_us'].items()})" 2>&1 | tail -1)"; tail -2 gpurun_out/run12_bench_$name.err; }
run default X=1
run adamfast B200RL_ADAM_FAST=1
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > gpurun_out/run12_phases.log 2>&1; tail -32 gpurun_out/run12_phases.log
