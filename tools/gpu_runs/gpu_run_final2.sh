#!/bin/bash
# round 2, evidence run on one GPU (ncu reports are summarised on the box and kept out of gpurun_out: 64 MiB cap)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/f2
for mode in 1 0; do for i in 1 2 3; do
  B200RL_ASYNC_INSERT=$mode timeout 300 python -m pytest tests/test_gpu_replay.py -q --tb=short -k concurrent_actor > ${O}_conc_${mode}_$i.log 2>&1; echo "concurrent test async=$mode try $i rc=$?"
done; done
grep -h "Error\|assert\|^E " ${O}_conc_*.log | sort | uniq -c | head -20
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_bf16_layers.py -q --tb=short -k "conv_fwd or linear_fwd or atari_network or learner_steps or c2_shape" > ${O}_fp32_tests.log 2>&1; echo "fp32-path tests rc=$?"; tail -5 ${O}_fp32_tests.log
b() { name=$1; shift; timeout 900 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), d.get('unit','')[:12], round(d.get('ms_per_step',0),4), round((d.get('e2e') or {}).get('value',0),1), 'cpu', (d.get('cpu_baseline') or {}).get('value'))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
b fp32 X=1 python bench.py --steps 200 --warmup 10 --precision fp32 --no-cpu-baseline
b fp32_old B200RL_SIMT_BIG=0 B200RL_SIMT_DGRAD_PHASES=0 python bench.py --steps 200 --warmup 10 --precision fp32 --no-cpu-baseline
b bf16 X=1 python bench.py --steps 1000 --warmup 20
b bf16_syncinsert B200RL_ASYNC_INSERT=0 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline
b dedup X=1 python bench.py --steps 1000 --warmup 20 --frame-dedup --no-cpu-baseline
b tf32 X=1 python bench.py --steps 500 --warmup 20 --precision tf32 --no-cpu-baseline
b d4pg X=1 python bench.py --workload d4pg --steps 1000 --warmup 20
b sumtree X=1 python bench.py --workload sumtree --steps 20 --warmup 3
b reference X=1 python bench.py --impl reference --steps 3 --warmup 1
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > ${O}_phases.log 2>&1; echo "phases rc=$?"
timeout 300 python tools/ncu_hbm_kernels.py > ${O}_hbm.log 2>&1; echo "hbm rc=$?"; cp gpurun_out/hbm_kernels.json ${O}_hbm_kernels.json
timeout 300 python bench.py --profile --steps 2 --warmup 5 --items 131072 > ${O}_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --profile --steps 2 --warmup 5 --items 131072 > ${O}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/ncu/step python bench.py --profile --steps 1 --warmup 5 --items 131072 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 ${O}_ncu_full.log
python tools/ncu_summarize.py /tmp/ncu/step.ncu-rep ${O}_ncu_full_step 1; echo "summarize rc=$?"
timeout 300 python tools/ncu_hbm_kernels.py --once > ${O}_hbm_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/ncu/hbm python tools/ncu_hbm_kernels.py --once > ${O}_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
python tools/ncu_summarize.py /tmp/ncu/hbm.ncu-rep ${O}_ncu_full_hbm 1; echo "summarize hbm rc=$?"
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
