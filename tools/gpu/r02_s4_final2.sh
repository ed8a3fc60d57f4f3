#!/bin/bash
# session 4, last tree: N = 2 gradient-sync check for per-language adapters + N = 2 bench, N = 1 bench, the new tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "gradient_cell or shared_operand or adapters_by" > gpurun_out/s4_new_tests.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/s4_new_tests.log | cut -c1-100)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py adapters_by > gpurun_out/s4_ddp_n2_adapters_by.log 2>&1; echo "ddp adapters_by rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/s4_ddp_n2_adapters_by.log | head -8 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s4_bench_n2_final.json 2> gpurun_out/s4_bench_n2_final.err; echo "bench n2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s4_bench_n2_final.json').read().strip().splitlines()[-1]); print('N=2:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/s4_bench_n1_final2.json 2> gpurun_out/s4_bench_n1_final2.err; echo "bench n1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s4_bench_n1_final2.json').read().strip().splitlines()[-1]); print('N=1:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), 'gemm', round(d['roofline']['frac'],4), round(d['roofline']['gemm_ms_per_step'],3), 'single', round(d['roofline']['single_stream_step_ms'],3), 'ncu frac', round(d['roofline']['frac_from_ncu_launch_list'],4))"
