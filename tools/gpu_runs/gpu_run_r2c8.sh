#!/bin/bash
# round 2, session 3, call 8: high-priority capture stream for the critical chain
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c8
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'group_us', round(d['roofline']['group_seconds']*1e6,1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b base X=1 $A
b prio B200RL_STREAM_PRIO=1 $A
b base2 X=1 $A
b prio2 B200RL_STREAM_PRIO=1 $A
B200RL_FINE=1 B200RL_STREAM_PRIO=1 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_prio.log 2>&1; echo "phases rc=$?"; tail -30 ${O}_phases_prio.log
