#!/bin/bash
# session 5: the training-shape LayerNorm backward kernel (csrc/norm.cu: ln_bwd_fast_kernel) + four loads in flight in grad_sqnorm:
# new kernel test first, then the full GPU suite, the N = 1 bench, the in-situ breakdown and memcheck of the LayerNorm tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ln_bwd_training_shape or dropout_add_ln" > gpurun_out/s5_ln_tests.log 2>&1; echo "ln tests rc=$?"; tail -15 gpurun_out/s5_ln_tests.log | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s5_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/s5_bench_n1.err | cut -c1-300
PB200_WGRAD_STREAM=0 timeout 300 python tools/step_breakdown.py > gpurun_out/s5_step_breakdown.txt 2>&1; echo "breakdown rc=$?"; head -24 gpurun_out/s5_step_breakdown.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "ln_bwd_training_shape or dropout_add_ln or adam" > gpurun_out/s5_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/s5_memcheck.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s5_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','gemm_ms_per_step','single_stream_step_ms')})
print('layernorm', d['families']['layernorm'])
print('logs', d['last_step_logs'])
PY
