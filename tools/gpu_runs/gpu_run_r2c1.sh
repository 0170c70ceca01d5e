#!/bin/bash
# round 2, session 3, call 1: concurrency fix check, DQfD tests, single-GPU pipelined update A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/c1
for i in 1 2 3 4; do
  timeout 200 python -m pytest tests/test_gpu_replay.py -q --tb=short -k concurrent_actor > ${O}_conc_$i.log 2>&1; echo "concurrent try $i rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_dqfd.py tests/test_gpu_learner.py -q --tb=short -k "dqfd or demo or learner_steps_match" > ${O}_new_tests.log 2>&1; echo "new tests rc=$?"; tail -30 ${O}_new_tests.log
b() { name=$1; shift; timeout 600 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'launches', d.get('gpu_launches'))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
A="python bench.py --steps 1000 --warmup 20 --items 262144 --no-cpu-baseline"
b serial B200RL_PIPELINE_1GPU=0 $A
b pipe_c0 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ADAM_CTAS=0 $A
b pipe_c2 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ADAM_CTAS=2 $A
b pipe_c4 B200RL_PIPELINE_1GPU=1 B200RL_PIPE_ADAM_CTAS=4 $A
b serial2 B200RL_PIPELINE_1GPU=0 $A
B200RL_FINE=1 B200RL_PIPELINE_1GPU=1 timeout 200 python tools/step_phases.py bf16 > ${O}_phases_pipe.log 2>&1; echo "phases rc=$?"; tail -40 ${O}_phases_pipe.log
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > ${O}_all_gpu_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -5 ${O}_all_gpu_tests.log
