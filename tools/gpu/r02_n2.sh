#!/bin/bash
# N = 2: gradient-sync check (eager + graph with the event-ordered NCCL overlap; replicas identical, update == single-process
# accumulation) and bench lines with and without the overlap
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/r02_ddp_n2.log 2>&1; echo "ddp rc=$?"; grep "eager\|graph\|DDP\|Error\|OK\|FAIL" gpurun_out/r02_ddp_n2.log | head -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1]); print('overlap on :', d['value'], d['ms_per_step'], d['e2e']['value'])"
PB200_GRAPH_OVERLAP=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench_n2_nooverlap.json 2> gpurun_out/r02_bench_n2_nooverlap.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n2_nooverlap.json').read().strip().splitlines()[-1]); print('overlap off:', d['value'], d['ms_per_step'], d['e2e']['value'])"
