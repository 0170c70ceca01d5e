#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 1000 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/run6_bench_$name.json 2> gpurun_out/run6_bench_$name.err; echo "bench $name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/run6_bench_$name.json'));print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; }
run default X=1
run split_c1w B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=1
run split_c2w B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=2
run split_c3w B200RL_SPLIT_ADAM=1 B200RL_ADAM_CTAS_PER_SM=3
run c8wide B200RL_ADAM_WIDE=1
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > gpurun_out/run6_phases.log 2>&1; echo "phases rc=$?"; tail -45 gpurun_out/run6_phases.log
timeout 600 python bench.py --workload d4pg --steps 300 --warmup 10 > gpurun_out/run6_d4pg.json 2> gpurun_out/run6_d4pg.err; echo "d4pg rc=$?"; cut -c1-600 gpurun_out/run6_d4pg.json; tail -3 gpurun_out/run6_d4pg.err
timeout 600 python bench.py --workload sumtree --steps 20 --warmup 3 > gpurun_out/run6_tree.json 2> gpurun_out/run6_tree.err; echo "tree rc=$?"; cut -c1-600 gpurun_out/run6_tree.json; tail -3 gpurun_out/run6_tree.err
