#!/bin/bash
# round 2, last session: four-GPU check of the data-parallel path with programmatic dependent launch on
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29517 bench.py --gpus 4 --steps 500 --warmup 20 --items 262144 --no-cpu-baseline > gpurun_out/n4j_pdl.json 2> gpurun_out/n4j_pdl.err; echo "bench n4 rc=$?"
python -c "
import json
for l in open('gpurun_out/n4j_pdl.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"; grep -v OMP gpurun_out/n4j_pdl.err | tail -3
