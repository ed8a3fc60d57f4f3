#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/s2_ddp_n2.log 2>&1; echo "ddp bf16 rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/s2_ddp_n2.log | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/ddp_check.py float16 > gpurun_out/s2_ddp_n2_fp16.log 2>&1; echo "ddp fp16 rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/s2_ddp_n2_fp16.log | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s2_bench_n2.json 2> gpurun_out/s2_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/s2_bench_n2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/s2_bench_n2.json').read().strip().splitlines()[-1]); print('N=2:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
