#!/bin/bash
mkdir -p gpurun_out
run() {  # tag, port, extra args...
  tag=$1; port=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 30 --warmup 5 "$@" > gpurun_out/s2_bench_n8_$tag.json 2> gpurun_out/s2_bench_n8_$tag.err; echo "bench $tag rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/s2_bench_n8_$tag.json').read().strip().splitlines()[-1]); print('N=8 $tag:', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
}
PB200_NCCL_MAX_CTAS=2 run c2 29541
PB200_NCCL_MAX_CTAS=4 run c4 29542
PB200_NCCL_MAX_CTAS=8 run c8 29543
PB200_NCCL_MAX_CTAS=4 run c4b50 29544 --bucket-mb 50
