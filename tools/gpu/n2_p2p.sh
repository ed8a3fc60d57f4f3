#!/bin/bash
# N = 2: peer-memory allreduce — gradient-sync check (eager + graph) and a bench line
PB200_P2P_ALLREDUCE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/p2p_ddp.log 2>&1; echo "ddp rc=$?"; grep "eager\|graph:\|DDP\|Error\|error\|peer-memory" gpurun_out/p2p_ddp.log | head
PB200_P2P_ALLREDUCE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/p2p_bench_n2.json 2> gpurun_out/p2p_bench_n2.err; echo "bench rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p2p_bench_n2.err | tail -4 | cut -c1-300; cut -c90-260 gpurun_out/p2p_bench_n2.json
