#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/ddp_check.py > gpurun_out/r02_ddp_n8.log 2>&1; echo "ddp rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/r02_ddp_n8.log | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n8.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]); print('N=8 overlap on :', d['value'], d['ms_per_step'], d['e2e']['value'])"
PB200_GRAPH_OVERLAP=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02_bench_n8_nooverlap.json 2> gpurun_out/r02_bench_n8_nooverlap.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n8_nooverlap.json').read().strip().splitlines()[-1]); print('N=8 overlap off:', d['value'], d['ms_per_step'], d['e2e']['value'])"
