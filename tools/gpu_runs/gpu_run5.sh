#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py tests/test_gpu_replay.py -q --tb=short -k "fused_head or save_restore or gather_rows" > gpurun_out/run5_a.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/run5_a.log
timeout 900 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "c2_shape or learner_steps" > gpurun_out/run5_learner.log 2>&1; echo "learner rc=$?"; tail -6 gpurun_out/run5_learner.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run5_bench_$name.json 2> gpurun_out/run5_bench_$name.err; echo "bench $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/run5_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))")"; tail -2 gpurun_out/run5_bench_$name.err; }
run default X=1
run split B200RL_SPLIT_ADAM=1
run split_c2 B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=2
run split_c4 B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=4
run split_c1 B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=1
timeout 300 python tools/step_phases.py bf16 > gpurun_out/run5_phases.log 2>&1; echo "phases rc=$?"; tail -9 gpurun_out/run5_phases.log
