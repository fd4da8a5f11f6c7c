#!/bin/bash
# Run the GPU kernel parity tests group by group, each in its own process under `timeout`, so a
# hang or a poisoned CUDA context in one kernel family does not hide the others.
# Usage (on the GPU box):  bash tools/run_gpu_checks.sh [groups...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GROUPS_DEFAULT="layernorm rmsnorm im2col space_to_depth s2_merge tsp embed_splice rope gemv decode_attention linear_plain linear_epilogues linear_posemb linear_swiglu fmha_noncausal fmha_causal fmha_paged"
GROUPS_TO_RUN="${@:-$GROUPS_DEFAULT}"
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
summary=gpurun_out/checks_summary.txt
: > $summary
for g in $GROUPS_TO_RUN; do
  log=gpurun_out/check_$g.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "$g" -p no:cacheprovider \
      --timeout 300 --timeout-method=thread --tb=short > $log 2>&1
  rc=$?
  echo "== $g rc=$rc :: $(tail -n 1 $log)" | tee -a $summary
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" $log | head -n 12 | tee -a $summary; fi
done
