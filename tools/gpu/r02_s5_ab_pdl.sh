#!/bin/bash
# session 5: same-box A/B of two builds of the library (build/lib_old.so: SIMT glue kernels plain-launched; build/lib_new.so:
# PDL-launched), alternating
mkdir -p gpurun_out
cp pasero_b200/libpasero_b200.so /tmp/lib_keep.so
for i in 1 2 3; do
  for v in new old; do
    cp build/lib_$v.so pasero_b200/libpasero_b200.so
    timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/s5_ab_$v$i.json 2> gpurun_out/s5_ab_$v$i.err
    python -c "
import json
d=json.loads(open('gpurun_out/s5_ab_$v$i.json').read().strip().splitlines()[-1])
print('$v $i ms', round(d['ms_per_step'],4), 'gemm', round(d['roofline']['gemm_ms_per_step'],4), 'single', round(d['roofline']['single_stream_step_ms'],4), 'clk', d['clocks']['sm_mhz'])
"
  done
done
cp /tmp/lib_keep.so pasero_b200/libpasero_b200.so
