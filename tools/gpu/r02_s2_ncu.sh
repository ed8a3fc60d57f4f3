#!/bin/bash
# one `ncu --set full` capture per hot kernel of the current tree (each after the same command ran clean without ncu);
# the text summaries are produced on the box (the GEMM reports are ~19 MB each, gpurun_out/ carries 64 MB at most)
mkdir -p gpurun_out /tmp/rep
N="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 600 $N -k regex:"$rx" -s $skip -c $cnt -o /tmp/rep/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  bash tools/ncu_summary.sh /tmp/rep/$name.ncu-rep > gpurun_out/r02_ncu_$name.txt 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv > /tmp/rep/$name.src.csv 2>/dev/null; echo '# top SASS lines by stall samples' >> gpurun_out/r02_ncu_$name.txt; python tools/ncu_hot.py /tmp/rep/$name.src.csv 25 >> gpurun_out/r02_ncu_$name.txt 2>&1
  sz=$(stat -c %s /tmp/rep/$name.ncu-rep); if [ "$sz" -lt 6000000 ]; then cp /tmp/rep/$name.ncu-rep gpurun_out/; fi
}
timeout 120 python tools/fbank_one.py 20 2>&1 | tail -1
cap fbank "fbank|cmvn" 4 2 python tools/fbank_one.py 2
timeout 300 python tools/microbench.py attn > gpurun_out/s2_microbench_attn.log 2>&1; echo "attn plain rc=$?"
cap attn_fwd attn_fwd_tc 7 1 python tools/microbench.py attn
cap attn_bwd attn_bwd_tc 7 1 python tools/microbench.py attn
timeout 300 python tools/ln_one.py > gpurun_out/s2_ln_plain.log 2>&1; echo "ln plain rc=$?"
cap ln_bwd dropout_add_ln_bwd 6 1 python tools/ln_one.py
cap ln_fwd dropout_add_ln_fwd 6 1 python tools/ln_one.py
timeout 300 python tools/gemm_one.py wgrad32b 16000 2048 512 > gpurun_out/g_wgrad_plain.log 2>&1; echo "wgrad plain rc=$?"
cap gemm_wgrad_epi7 gemm_kernel 3 1 python tools/gemm_one.py wgrad32b 16000 2048 512
cap gemm_fwd gemm_kernel 3 1 python tools/gemm_one.py fwd 16000 2048 512
cap gemm_fwd_small gemm_kernel 3 1 python tools/gemm_one.py fwd 3072 512 512
du -sh gpurun_out
