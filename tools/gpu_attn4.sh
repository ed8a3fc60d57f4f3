timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k attention > gpurun_out/attn_tc.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/attn_tc.log
timeout 300 python tools/microbench.py attn > gpurun_out/attn_mb_tc.log 2>&1; echo "mb rc=$?"; grep attn gpurun_out/attn_mb_tc.log
PB200_NVCC_EXTRA=-DPB_ATTN_TIMING python -c "
from pasero_b200 import build; build.build(force=True)" > gpurun_out/tl_build.log 2>&1; echo "build rc=$?"
timeout 300 python tools/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/attn_timeline.log
