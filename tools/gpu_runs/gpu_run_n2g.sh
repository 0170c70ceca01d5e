#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_peer_exchange.py -q --tb=short -x > gpurun_out/n2g_peer.log 2>&1; echo "peer test rc=$?"; tail -12 gpurun_out/n2g_peer.log
for pc in 1 2 4 8; do
B200RL_DP_CE_PIECES=$pc B200RL_DP_CE=1 B200RL_FINE=1 timeout 300 $TR --master-port 29514 tools/step_phases.py bf16 > gpurun_out/n2g_phases_p$pc.log 2>&1; echo "phases pieces=$pc rc=$?"; grep -v OMP gpurun_out/n2g_phases_p$pc.log | grep "copy-engine\|step\.\|sum"
done
run() { name=$1; shift; env "$@" timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/n2g_$name.json 2> gpurun_out/n2g_$name.err; echo "bench $name rc=$? $(python -c "
import json
for l in open('gpurun_out/n2g_$name.json'):
  if l.startswith('{'):
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; grep -v OMP gpurun_out/n2g_$name.err | tail -3; }
run ce_p1 B200RL_DP_CE=1 B200RL_DP_CE_PIECES=1
run ce_p4 B200RL_DP_CE=1 B200RL_DP_CE_PIECES=4
run ce_p8 B200RL_DP_CE=1 B200RL_DP_CE_PIECES=8
timeout 600 python bench.py --workload d4pg --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/n2g_d4pg.json 2> gpurun_out/n2g_d4pg.err; echo "d4pg rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/n2g_d4pg.json') if l.startswith('{')][0]); print('d4pg', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"; tail -3 gpurun_out/n2g_d4pg.err
B200RL_D4PG_STREAMS=0 timeout 600 python bench.py --workload d4pg --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/n2g_d4pg_1s.json 2> gpurun_out/n2g_d4pg_1s.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/n2g_d4pg_1s.json') if l.startswith('{')][0]); print('d4pg one stream', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
timeout 600 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "d4pg or ddpg" > gpurun_out/n2g_d4pg_tests.log 2>&1; echo "d4pg tests rc=$?"; tail -5 gpurun_out/n2g_d4pg_tests.log
