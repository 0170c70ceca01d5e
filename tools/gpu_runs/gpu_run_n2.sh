#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_peer_exchange.py -q --tb=short > gpurun_out/n2_peer.log 2>&1; echo "peer test rc=$?"; tail -30 gpurun_out/n2_peer.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"; python -c "
import json
for l in open('gpurun_out/n2_bench.json'):
  if l.startswith('{'):
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parallelism'])"; tail -5 gpurun_out/n2_bench.err
