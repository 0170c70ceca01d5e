#!/bin/bash
# round 2, session 3, call 2: pipelined single-GPU update -- stream priorities, fork order, cache-streaming Adam
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c2
timeout 600 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "pipelined or learner_steps_match or adam" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -25 ${O}_tests.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'adam_us', round(d['stages_us'].get('k7_adam',0),1))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b serial B200RL_PIPELINE_1GPU=0 $A
b serial_cs B200RL_PIPELINE_1GPU=0 B200RL_ADAM_CS=1 $A
b p1_o0 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=0 $A
b p1_o1 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=1 $A
b p1_o1_cs B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=1 B200RL_ADAM_CS=1 $A
b p1_o0_cs B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=0 B200RL_ADAM_CS=1 $A
b p1_o1_cs_c2 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=1 B200RL_ADAM_CS=1 B200RL_PIPE_ADAM_CTAS=2 $A
b p0_o1_c2 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=0 B200RL_PIPE_ORDER=1 B200RL_PIPE_ADAM_CTAS=2 $A
b p0_o1_cs_c3 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=0 B200RL_PIPE_ORDER=1 B200RL_ADAM_CS=1 B200RL_PIPE_ADAM_CTAS=3 $A
B200RL_FINE=1 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=0 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_p1_o0.log 2>&1; echo "phases rc=$?"; tail -27 ${O}_phases_p1_o0.log
B200RL_FINE=1 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_PRIO=1 B200RL_PIPE_ORDER=1 B200RL_ADAM_CS=1 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_p1_o1_cs.log 2>&1; echo "phases rc=$?"; tail -27 ${O}_phases_p1_o1_cs.log
