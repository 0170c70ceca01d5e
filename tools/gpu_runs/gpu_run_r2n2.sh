#!/bin/bash
# round 2, session 3: two-GPU check of the data-parallel path with programmatic dependent launch on (default) and off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -q --tb=short -x > gpurun_out/n2j_peer.log 2>&1; echo "peer test rc=$?"; tail -6 gpurun_out/n2j_peer.log
run() { name=$1; shift; env "$@" timeout 400 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --items 262144 --no-cpu-baseline > gpurun_out/n2j_$name.json 2> gpurun_out/n2j_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2j_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2j_$name.err | tail -3; }
run pdl X=1
run nopdl B200RL_PDL=0
timeout 300 python bench.py --steps 500 --warmup 20 --items 262144 --no-cpu-baseline > gpurun_out/n2j_n1.json 2> gpurun_out/n2j_n1.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/n2j_n1.json') if l.startswith('{')][-1]); print('n1', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
