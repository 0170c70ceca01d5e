#!/bin/bash
# round 2, session 3, call 4: PDL trigger point (CTA start / accumulator ready) and scope (all / no target forward / backward only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c4
B200RL_PDL=1 B200RL_PDL_LATE=1 B200RL_K2_EARLY=1 timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_bf16_layers.py -q --tb=short -x > ${O}_tests_pdl_late.log 2>&1; echo "tests with PDL late rc=$?"; tail -4 ${O}_tests_pdl_late.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
P="B200RL_PDL=1 B200RL_K2_EARLY=1"
b all_early $P B200RL_PDL_LATE=0 $A
b all_late $P B200RL_PDL_LATE=1 $A
b noft_early $P B200RL_PDL_LATE=0 B200RL_PDL_SCOPE=noft $A
b noft_late $P B200RL_PDL_LATE=1 B200RL_PDL_SCOPE=noft $A
b bwd_early $P B200RL_PDL_LATE=0 B200RL_PDL_SCOPE=bwd $A
b bwd_late $P B200RL_PDL_LATE=1 B200RL_PDL_SCOPE=bwd $A
b all_late_pipe $P B200RL_PDL_LATE=1 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ORDER=0 $A
env $P B200RL_FINE=1 B200RL_PDL_LATE=1 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_all_late.log 2>&1; echo "phases rc=$?"; tail -30 ${O}_phases_all_late.log
env $P B200RL_FINE=1 B200RL_PDL_LATE=1 B200RL_PDL_SCOPE=noft timeout 200 python tools/step_phases.py bf16 > ${O}_phases_noft_late.log 2>&1; echo "phases rc=$?"; tail -30 ${O}_phases_noft_late.log
