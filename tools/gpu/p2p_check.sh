#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu/p2p_check.sh N': stand-alone check of the peer-memory allreduce kernel (bounded)
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    tools/p2p_check.py > gpurun_out/p2p_check_n$N.log 2>&1
echo "p2p_check N=$N rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p2p_check_n$N.log | tail -20
PB200_P2P_MAP=symm timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    tools/p2p_check.py > gpurun_out/p2p_check_symm_n$N.log 2>&1
echo "p2p_check (symmetric memory) N=$N rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p2p_check_symm_n$N.log | tail -20
PB200_P2P_MAP=symm PB200_P2P_NVLS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    tools/p2p_check.py > gpurun_out/p2p_check_nvls_n$N.log 2>&1
echo "p2p_check (NVLS multimem) N=$N rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/p2p_check_nvls_n$N.log | tail -20
