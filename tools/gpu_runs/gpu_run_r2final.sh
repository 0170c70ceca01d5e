#!/bin/bash
# round 2, last session: evidence run on one GPU with the final defaults (PDL on, K2 beside the backward)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/f3
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > ${O}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
b() { name=$1; shift; timeout 900 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), d.get('unit','')[:10], round(d.get('ms_per_step',0),4), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
b bf16 X=1 python bench.py --steps 1000 --warmup 20
b driver_like X=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
b bf16_nopdl B200RL_PDL=0 B200RL_K2_EARLY=0 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline
b bf16_pipelined B200RL_PIPELINE_1GPU=1 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline
b dedup X=1 python bench.py --steps 1000 --warmup 20 --frame-dedup --no-cpu-baseline
b tf32 X=1 python bench.py --steps 500 --warmup 20 --precision tf32 --no-cpu-baseline
b fp32 X=1 python bench.py --steps 200 --warmup 10 --precision fp32 --no-cpu-baseline
b d4pg X=1 python bench.py --workload d4pg --steps 1000 --warmup 20
b reference X=1 python bench.py --impl reference --steps 3 --warmup 1
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > ${O}_phases.log 2>&1; echo "phases rc=$?"
timeout 300 python bench.py --profile --steps 2 --warmup 5 --items 131072 > ${O}_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --profile --steps 2 --warmup 5 --items 131072 > ${O}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/ncu/step python bench.py --profile --steps 1 --warmup 5 --items 131072 > ${O}_ncu_full.log 2>&1; rc=$?; echo "ncu full rc=$rc"; tail -2 ${O}_ncu_full.log
if [ $rc -ne 0 ]; then
  B200RL_PDL=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/ncu/step python bench.py --profile --steps 1 --warmup 5 --items 131072 > ${O}_ncu_full_nopdl.log 2>&1; echo "ncu full (PDL off) rc=$?"
fi
python tools/ncu_summarize.py /tmp/ncu/step.ncu-rep ${O}_ncu_full_step 1; echo "summarize rc=$?"
du -sh gpurun_out
