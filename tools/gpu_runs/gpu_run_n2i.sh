#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests/test_gpu_peer_exchange.py -q --tb=long -x > gpurun_out/n2i_peer.log 2>&1; echo "peer test rc=$?"; tail -25 gpurun_out/n2i_peer.log
run() { name=$1; shift; env "$@" timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2i_$name.json 2> gpurun_out/n2i_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2i_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['config']['parallelism'][100:190])" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2i_$name.err | tail -3; }
run ce_ce X=1
run mc_ce B200RL_DP_REDUCE=mc
run ce_mc B200RL_DP_BCAST=mc
timeout 300 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "c51 or d4pg" > gpurun_out/n2i_c51.log 2>&1; echo "c51 tests rc=$?"; tail -4 gpurun_out/n2i_c51.log
timeout 300 python tools/ncu_hbm_kernels.py > gpurun_out/n2i_hbm.log 2>&1; echo "hbm rc=$?"; tail -30 gpurun_out/n2i_hbm.log
