#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_peer_exchange.py -q --tb=short -x > gpurun_out/n2d_peer.log 2>&1; echo "peer test rc=$?"; tail -30 gpurun_out/n2d_peer.log
run() { name=$1; shift; env "$@" timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2d_$name.json 2> gpurun_out/n2d_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2d_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2d_$name.err | tail -3; }
run base X=1
run ce B200RL_DP_CE=1
run ce64 B200RL_DP_CE=1 B200RL_DP_CE_CTAS=74
run ce148 B200RL_DP_CE=1 B200RL_DP_CE_CTAS=148
B200RL_DP_CE=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2d_phases_ce.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n2d_phases_ce.log | tail -34
