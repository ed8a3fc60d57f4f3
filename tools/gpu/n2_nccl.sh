#!/bin/bash
# N = 2: gradient-sync check (eager + graph; replicas identical, update == single-process accumulation) and a bench line
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_n2_v3.log 2>&1; echo "ddp rc=$?"; grep "eager\|graph:\|DDP\|Error\|OK\|FAIL" gpurun_out/ddp_n2_v3.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_v3.json 2> gpurun_out/bench_n2_v3.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n2_v3.err | cut -c1-300; cut -c90-260 gpurun_out/bench_n2_v3.json
