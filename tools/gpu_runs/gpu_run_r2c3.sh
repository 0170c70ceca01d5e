#!/bin/bash
# round 2, session 3, call 3: programmatic dependent launch along the step's kernel chain; K2 beside the backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c3
B200RL_PDL=1 timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_bf16_layers.py tests/test_gpu_replay.py tests/test_gpu_agents.py -q --tb=short -x > ${O}_tests_pdl.log 2>&1; echo "tests with PDL rc=$?"; tail -25 ${O}_tests_pdl.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b serial B200RL_PDL=0 $A
b k2early B200RL_PDL=0 B200RL_K2_EARLY=1 $A
b pdl B200RL_PDL=1 $A
b pdl_k2early B200RL_PDL=1 B200RL_K2_EARLY=1 $A
b pdl_pipe B200RL_PDL=1 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ORDER=0 $A
b pdl_fp32 B200RL_PDL=1 python bench.py --steps 100 --warmup 10 --items 262144 --no-cpu-baseline --precision fp32
b pdl_d4pg B200RL_PDL=1 python bench.py --workload d4pg --steps 500 --warmup 20 --items 262144 --no-cpu-baseline
B200RL_FINE=1 B200RL_PDL=1 B200RL_K2_EARLY=1 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_pdl.log 2>&1; echo "phases rc=$?"; tail -34 ${O}_phases_pdl.log
