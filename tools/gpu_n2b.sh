timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_n2.log 2>&1; echo "ddp rc=$?"; grep "eager\|graph:\|DDP\|Error" gpurun_out/ddp_n2.log | head
for v in 1 0; do
PB200_GRAPH_NCCL=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_g$v.json 2> gpurun_out/bench_n2_g$v.err; echo "bench graph_nccl=$v rc=$?"; tail -2 gpurun_out/bench_n2_g$v.err | cut -c1-300; cut -c90-260 gpurun_out/bench_n2_g$v.json
done
