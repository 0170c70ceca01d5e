#!/bin/bash
# round 2, session 3, call 5: PDL as the default everywhere (DQN bf16/tf32, D4PG, helpers): whole GPU suite + A/B benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c5
timeout 900 python -m pytest tests -m gpu -q --tb=short > ${O}_all_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -12 ${O}_all_gpu_tests.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="--steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b dqn X=1 python bench.py $A
b dqn_nopdl B200RL_PDL=0 python bench.py $A
b d4pg X=1 python bench.py --workload d4pg $A
b d4pg_nopdl B200RL_PDL=0 python bench.py --workload d4pg $A
b d4pg_early B200RL_PDL_LATE=0 python bench.py --workload d4pg $A
b tf32 X=1 python bench.py --precision tf32 --steps 500 --warmup 20 --items 262144 --no-cpu-baseline
b tf32_nopdl B200RL_PDL=0 python bench.py --precision tf32 --steps 500 --warmup 20 --items 262144 --no-cpu-baseline
b fp32 X=1 python bench.py --precision fp32 --steps 100 --warmup 10 --items 262144 --no-cpu-baseline
b dedup X=1 python bench.py $A --frame-dedup
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 ${O}_smoke.log
