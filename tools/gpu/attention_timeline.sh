cp pasero_b200/libpasero_b200.so /tmp/lib_backup.so
PB200_NVCC_EXTRA=-DPB_ATTN_TIMING python -c "
from pasero_b200 import build; build.build(force=True)" > gpurun_out/tl_build.log 2>&1; echo "build rc=$?"
timeout 300 python tools/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/attn_timeline.log
