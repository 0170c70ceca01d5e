#!/bin/bash
# round 2, GPU call 2: bf16 dataflow kernels + fused head/TD, then the whole suite and benches per precision
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_layers.py -q --tb=short > gpurun_out/run2_layers.log 2>&1; echo "layers rc=$?"; tail -25 gpurun_out/run2_layers.log
timeout 900 python -m pytest tests/test_gpu_learner.py -q --tb=short -k "c2_shape or atari_network or jax_variants or learner_steps or td_kernel or adam" > gpurun_out/run2_learner.log 2>&1; echo "learner rc=$?"; tail -25 gpurun_out/run2_learner.log
timeout 1200 python -m pytest tests -m gpu -q --tb=line > gpurun_out/run2_all.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/run2_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/run2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/run2_smoke.log
for prec in bf16 tf32; do
  timeout 600 python bench.py --steps 500 --warmup 20 --precision $prec --no-cpu-baseline > gpurun_out/run2_bench_$prec.json 2> gpurun_out/run2_bench_$prec.err; echo "bench $prec rc=$?"; cut -c1-400 gpurun_out/run2_bench_$prec.json; tail -3 gpurun_out/run2_bench_$prec.err
done
ls gpurun_out | head -40
