for v in 0 1; do echo "PB200_PDL=$v"; PB200_PDL=$v timeout 300 python tools/mb_small.py 2>&1 | grep "tn256 cl2\|wgrad32" ; done
PB200_PDL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k gemm 2>&1 | tail -2
