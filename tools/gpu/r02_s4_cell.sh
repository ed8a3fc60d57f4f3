#!/bin/bash
# session 4: the shared encoder-output gradient cell (sum of the decoder layers' key / value gradients in the dgrad epilogues)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_real_shapes_gpu.py tests/test_parity_gpu.py tests/test_fp16_gpu.py -m gpu -q -x > gpurun_out/s4_cell_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s4_cell_tests.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --no-eager > gpurun_out/s4_bench_cell.json 2> gpurun_out/s4_bench_cell.err; echo "bench rc=$?"; tail -2 gpurun_out/s4_bench_cell.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/s4_bench_cell.json').read().strip().splitlines()[-1]); print('N=1:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), 'gemm', round(d['roofline']['frac'],4), round(d['roofline']['gemm_ms_per_step'],3), 'single', round(d['roofline']['single_stream_step_ms'],3), 'launches', d['gpu_launches'])"
