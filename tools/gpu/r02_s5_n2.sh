#!/bin/bash
# session 5 (N = 2): gradient-sync checks on the last tree (bfloat16, float16) and the N = 2 / N = 1 bench in the same call
mkdir -p gpurun_out
for mode in "" float16; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py $mode > gpurun_out/s5_ddp_n2_${mode:-bf16}.log 2>&1; echo "ddp ${mode:-bf16} rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/s5_ddp_n2_${mode:-bf16}.log | head -8 | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s5_bench_n2.json 2> gpurun_out/s5_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/s5_bench_n2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/s5_bench_n2.json').read().strip().splitlines()[-1]); print('N=2:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/s5_bench_n1_same_box.json 2> gpurun_out/s5_bench_n1_same_box.err; echo "bench n1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/s5_bench_n1_same_box.json').read().strip().splitlines()[-1]); print('N=1:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
