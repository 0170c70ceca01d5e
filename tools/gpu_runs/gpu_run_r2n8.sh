#!/bin/bash
# round 2, last session: eight-GPU check of the data-parallel path with programmatic dependent launch on (default reduce = mc at 8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29518 bench.py --gpus 8 --steps 400 --warmup 20 --items 262144 --no-cpu-baseline > gpurun_out/n8j_pdl.json 2> gpurun_out/n8j_pdl.err; echo "bench n8 rc=$?"
python -c "
import json
for l in open('gpurun_out/n8j_pdl.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"; grep -v OMP gpurun_out/n8j_pdl.err | tail -3
