timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_n2.log 2>&1; echo "ddp rc=$?"; tail -5 gpurun_out/ddp_n2.log
echo skip
echo skip
echo skip
