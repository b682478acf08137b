#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / TMA pipelines (run on the GPU box through gpurun):
#   tools/sanitize.sh memcheck|racecheck|synccheck  ->  gpurun_out/sanitizer_<tool>.log
# The selected tests cover gemm_bf16_tc_kernel (single CTA and CTA-pair, all operand majors, split variant),
# attn_fwd_tc_kernel / attn_bwd_tc_kernel (bf16 and split variants) and the MAGNO gather kernels at small shapes.
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout ${2:-900} compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 \
  python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -q -x --timeout 800 \
  -k "(gemm_bf16_vs_fp64 and 128-128-64) or (gemm_bf16_vs_fp64 and 1000-200-136) or (gemm_bf16_tn and 64-64-512) or (gemm_bf16_nn and 128-64-64) or (gemm_split_forward and 256-64-256) or (gemm_split_weight and 64-64-512) or (attention_fwd_bf16 and 2-256-2-2-32) or (attention_bwd_bf16 and 2-256-2-2-32) or (attention_split and 3-64-2-2-16 and 1.0) or (attention_split and 2-256-2-2-32 and 1.0) or (agno_dot_both and rand2d) or fused_lift" \
  > gpurun_out/sanitizer_$tool.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_$tool.log
tail -15 gpurun_out/sanitizer_$tool.log
