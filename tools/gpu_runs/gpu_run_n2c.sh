#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2c_$name.json 2> gpurun_out/n2c_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2c_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2c_$name.err | tail -2; }
run cta3 X=1
run cta1 B200RL_DP_CTAS_PER_SM=1
run cta2 B200RL_DP_CTAS_PER_SM=2
run early1 B200RL_DP_EARLY_TAIL=1 B200RL_DP_CTAS_PER_SM=1
run early3 B200RL_DP_EARLY_TAIL=1
B200RL_DP_CTAS_PER_SM=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2c_phases_cta1.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n2c_phases_cta1.log | tail -34
