#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { n=$1; name=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $n --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n8_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n8_$name.err | tail -3; }
run 8 ce8 X=1
run 8 ce8_lanes1 B200RL_DP_CE_LANES=1
run 8 ce8_lanes4 B200RL_DP_CE_LANES=4
run 8 sm8 B200RL_DP_CE=0
run 4 ce4 X=1
B200RL_FINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n8_phases_ce.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n8_phases_ce.log | grep "copy-engine\|step\.\|sum\|on\.\|tgt\."
timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n8_n1.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/n8_n1.json') if l.startswith('{')][0]); print('n1', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
