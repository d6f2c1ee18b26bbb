#!/bin/bash
# One `ncu --set full` capture per hot kernel of the configs[1] step (GPU box, one GPU). The second step of profiles/prof_step.py is
# the one profiled (launch-skip = launches of that kernel in the first step). Reports land in gpurun_out/, summaries are made in
# the build container with profiles/ncu_tools.py and committed under profiles/rNN/.
set -u
tag=${1:-r1k}
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip "$3" --launch-count 1 \
      -o "gpurun_out/${tag}_$1" -f python profiles/prof_step.py > "gpurun_out/ncu_${tag}_$1.log" 2>&1
  echo "$1 rc=$?"
}
# GEMM launch order inside one step: conv1..6 (0-5), feature projection (6), 4 pos-conv blocks (7-10), per encoder layer l:
# qkv 11+4l, out-proj 12+4l, ffn1 13+4l, ffn2 14+4l; vertex head 59. Second step = +60.
cap gemm_conv1 gemm_tc2_kernel 60
cap gemm_qkv gemm_tc2_kernel 71
cap gemm_oproj gemm_tc2_kernel 72
cap gemm_ffn1 gemm_tc2_kernel 73
cap gemm_ffn2 gemm_tc2_kernel 74
cap gemm_vhead gemm_tc2_kernel 119
cap conv0 conv0_tc 1
cap attn attn_tc 12
cap ar_decoder ff_decoder_ar64 1
cap layernorm layernorm_vec 24
cap flame flame_tc 1
ls -la gpurun_out/${tag}_*.ncu-rep
