#!/bin/bash
# round 2, session 3, call 6: PDL domains (tensor-core chain on, small kernels off): A/B per workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c6
timeout 600 python -m pytest tests/test_gpu_learner.py -q --tb=short -x > ${O}_tests.log 2>&1; echo "learner tests rc=$?"; tail -3 ${O}_tests.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="--steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b dqn X=1 python bench.py $A
b dqn_small B200RL_PDL_SMALL=1 python bench.py $A
b d4pg X=1 python bench.py --workload d4pg $A
b d4pg_nopdl B200RL_PDL=0 python bench.py --workload d4pg $A
b fp32 X=1 python bench.py --precision fp32 --steps 100 --warmup 10 --items 262144 --no-cpu-baseline
b fp32_nopdl B200RL_PDL=0 python bench.py --precision fp32 --steps 100 --warmup 10 --items 262144 --no-cpu-baseline
b tf32 X=1 python bench.py --precision tf32 --steps 500 --warmup 20 --items 262144 --no-cpu-baseline
b dqn_pipe B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ORDER=0 python bench.py $A
