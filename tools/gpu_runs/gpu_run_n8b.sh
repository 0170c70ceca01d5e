#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { n=$1; name=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $n --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n8b_$name.json 2> gpurun_out/n8b_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n8b_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['config']['parallelism'][140:200])" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n8b_$name.err | tail -3; }
run 8 mc_mc B200RL_DP_REDUCE=mc B200RL_DP_BCAST=mc
run 8 mcfused B200RL_DP_REDUCE=mcfused
run 8 mc_ce B200RL_DP_REDUCE=mc B200RL_DP_BCAST=ce
run 8 ce_ce B200RL_DP_REDUCE=ce B200RL_DP_BCAST=ce
run 8 ce_mc B200RL_DP_REDUCE=ce B200RL_DP_BCAST=mc
run 8 mc_mc_c296 B200RL_DP_REDUCE=mc B200RL_DP_BCAST=mc B200RL_DP_MC_CTAS=296
run 4 n4_ce_ce B200RL_DP_REDUCE=ce B200RL_DP_BCAST=ce
run 4 n4_mc_mc B200RL_DP_REDUCE=mc B200RL_DP_BCAST=mc
B200RL_DP_REDUCE=mc B200RL_DP_BCAST=mc B200RL_FINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n8b_phases_mc.log 2>&1; echo "phases rc=$?"; grep -v OMP gpurun_out/n8b_phases_mc.log | grep "copy-engine\|last exchange\|step\.\|sum\|on\.\|tgt\."
