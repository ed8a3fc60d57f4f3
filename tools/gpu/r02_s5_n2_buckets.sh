#!/bin/bash
# session 5 (N = 2): bucket-size sweep of the gradient allreduce on the last tree
mkdir -p gpurun_out
for mb in 25 50 80 150 12 25; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 --bucket-mb $mb --no-cpu-baseline --no-eager > gpurun_out/s5_n2_b$mb.json 2> gpurun_out/s5_n2_b$mb.err
  python -c "
import json; d=json.loads(open('gpurun_out/s5_n2_b$mb.json').read().strip().splitlines()[-1]); print('bucket $mb MB: N=2', round(d['value']), round(d['ms_per_step'],3))"
done
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/s5_n1_ref.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/s5_n1_ref.json').read().strip().splitlines()[-1]); print('N=1:', round(d['value']), round(d['ms_per_step'],3))"
