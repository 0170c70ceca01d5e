#!/bin/bash
# round 2, evidence run on one GPU: full GPU test-suite, smoke, every bench line, timelines, ncu launch list + --set full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/f1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 ${O}_pytest.log
timeout 300 python __graft_entry__.py smoke > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
b() { name=$1; shift; timeout 900 env "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$? $(python -c "
import json
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print(round(d.get('value',0),1), d.get('unit','')[:12], round(d.get('ms_per_step',0),4), round((d.get('e2e') or {}).get('value',0),1), 'cpu', (d.get('cpu_baseline') or {}).get('value'))" 2>&1 | tail -1)"; tail -2 ${O}_bench_$name.err; }
b bf16 X=1 python bench.py --steps 1000 --warmup 20
b bf16_syncinsert B200RL_ASYNC_INSERT=0 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline
b dedup X=1 python bench.py --steps 1000 --warmup 20 --frame-dedup --no-cpu-baseline
b tf32 X=1 python bench.py --steps 500 --warmup 20 --precision tf32 --no-cpu-baseline
b fp32 X=1 python bench.py --steps 200 --warmup 10 --precision fp32 --no-cpu-baseline
b d4pg X=1 python bench.py --workload d4pg --steps 1000 --warmup 20
b sumtree X=1 python bench.py --workload sumtree --steps 20 --warmup 3
b reference X=1 python bench.py --impl reference --steps 3 --warmup 1
B200RL_FINE=1 timeout 300 python tools/step_phases.py bf16 > ${O}_phases.log 2>&1; echo "phases rc=$?"; tail -34 ${O}_phases.log
timeout 300 python tools/ncu_hbm_kernels.py > ${O}_hbm.log 2>&1; echo "hbm rc=$?"
timeout 300 python bench.py --profile --steps 2 --warmup 5 > ${O}_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --profile --steps 2 --warmup 5 > ${O}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"; tail -2 ${O}_ncu_launch.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o ${O}_step python bench.py --profile --steps 1 --warmup 5 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 ${O}_ncu_full.log
timeout 300 python tools/ncu_hbm_kernels.py --once > ${O}_hbm_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o ${O}_hbm python tools/ncu_hbm_kernels.py --once > ${O}_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"; tail -2 ${O}_ncu_hbm.log
ls -la gpurun_out | grep f1_ | head -40
