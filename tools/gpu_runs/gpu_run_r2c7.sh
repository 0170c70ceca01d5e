#!/bin/bash
# round 2, session 3, call 7: split-K depth sweep for the few-tile layers (conv weight gradients, fc1 forward)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c7
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b base X=1 $A
b o2_m8_x128 B200RL_SPLITK_OCC=2 B200RL_SPLITK_MIN_KB=8 B200RL_SPLITK_MAX=128 $A
b o2_m4_x128 B200RL_SPLITK_OCC=2 B200RL_SPLITK_MIN_KB=4 B200RL_SPLITK_MAX=128 $A
b o2_m4_x256 B200RL_SPLITK_OCC=2 B200RL_SPLITK_MIN_KB=4 B200RL_SPLITK_MAX=256 $A
b o1_m4_x128 B200RL_SPLITK_OCC=1 B200RL_SPLITK_MIN_KB=4 B200RL_SPLITK_MAX=128 $A
b o3_m4_x256 B200RL_SPLITK_OCC=3 B200RL_SPLITK_MIN_KB=4 B200RL_SPLITK_MAX=256 $A
b o2_m16_x128 B200RL_SPLITK_OCC=2 B200RL_SPLITK_MIN_KB=16 B200RL_SPLITK_MAX=128 $A
b fp32 X=1 python bench.py --precision fp32 --steps 100 --warmup 10 --items 262144 --no-cpu-baseline
B200RL_FINE=1 B200RL_SPLITK_OCC=2 B200RL_SPLITK_MIN_KB=4 B200RL_SPLITK_MAX=128 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_o2_m4.log 2>&1; echo "phases rc=$?"; tail -30 ${O}_phases_o2_m4.log
