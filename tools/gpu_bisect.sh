for flags in "-DPB_OLD_CDE" ""; do
PB200_NVCC_EXTRA="$flags" python -c "
from pasero_b200 import build; build.build(force=True)" > gpurun_out/bis_build.log 2>&1; echo "build [$flags] rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -3
done
