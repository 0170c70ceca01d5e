#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_sequences.py -q --tb=short > gpurun_out/n2f_seq.log 2>&1; echo "seq tests rc=$?"; tail -5 gpurun_out/n2f_seq.log
B200RL_DP_CE=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2f_phases_ce.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n2f_phases_ce.log | grep "copy-engine\|last exchange\|step\.\|sum"
CUDA_DEVICE_MAX_CONNECTIONS=32 B200RL_DP_CE=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2f_phases_ce32.log 2>&1; echo "phases conn32 rc=$?"; grep -v OMP gpurun_out/n2f_phases_ce32.log | grep "copy-engine\|last exchange\|step\.\|sum"
run() { name=$1; shift; env "$@" timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2f_$name.json 2> gpurun_out/n2f_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2f_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2f_$name.err | tail -3; }
run ce B200RL_DP_CE=1
run ce_conn32 B200RL_DP_CE=1 CUDA_DEVICE_MAX_CONNECTIONS=32
run base_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2f_n1_conn32.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/n2f_n1_conn32.json') if l.startswith('{')][0]); print('n1 conn32', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
