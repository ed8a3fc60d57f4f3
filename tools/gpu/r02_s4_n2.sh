#!/bin/bash
# session 4, call 2 (N = 2): gradient-sync checks (bfloat16, float16, per-language adapters) and the N = 2 bench with the gradient
# buckets flushed on the weight-gradient stream (no join of the main stream per bucket); the new float16 large-vocabulary tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fp16_gpu.py -m gpu -q -x -k "chunked or large_vocabulary" > gpurun_out/s4_fp16_chunk.log 2>&1; echo "fp16 chunk rc=$?"; tail -15 gpurun_out/s4_fp16_chunk.log | cut -c1-300
for mode in "" float16 adapters_by; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py $mode > gpurun_out/s4_ddp_n2_${mode:-bf16}.log 2>&1; echo "ddp ${mode:-bf16} rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/s4_ddp_n2_${mode:-bf16}.log | head -8 | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s4_bench_n2.json 2> gpurun_out/s4_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/s4_bench_n2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/s4_bench_n2.json').read().strip().splitlines()[-1]); print('N=2:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err; echo "bench n1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s4_bench_n1.json').read().strip().splitlines()[-1]); print('N=1:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
