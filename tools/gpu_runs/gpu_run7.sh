#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_replay.py -q --tb=short > gpurun_out/run7_replay.log 2>&1; echo "replay rc=$?"; tail -12 gpurun_out/run7_replay.log
timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline --frame-dedup > gpurun_out/run7_bench_dedup.json 2> gpurun_out/run7_bench_dedup.err; echo "bench dedup rc=$?"; cut -c1-300 gpurun_out/run7_bench_dedup.json; tail -3 gpurun_out/run7_bench_dedup.err
python -c "
import json;d=json.load(open('gpurun_out/run7_bench_dedup.json'));print(d['value'], d['e2e'], d['stages_us'])"
