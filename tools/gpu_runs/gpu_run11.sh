#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/h_timeline.py > gpurun_out/run11_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/run11_timeline.log | tail -24
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_replay.py tests/test_gpu_agents.py -q --tb=short > gpurun_out/run11_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/run11_tests.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run11_bench_$name.json 2> gpurun_out/run11_bench_$name.err; echo "bench $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/run11_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches']//1000)" 2>&1 | tail -1)"; tail -2 gpurun_out/run11_bench_$name.err; }
run default X=1
run nosplit B200RL_SPLIT_ADAM=0
run tail3 B200RL_TAIL_ADAM_CTAS=3
run tail1 B200RL_TAIL_ADAM_CTAS=1
env timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline --frame-dedup > gpurun_out/run11_bench_dedup.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/run11_bench_dedup.json'));print('dedup', round(d['value'],1), round(d['e2e']['value'],1))"
