#!/bin/bash
# One GPU session: e2e parity tests, smoke, bench, ncu launch list + one full capture of the
# dominant kernel.  Everything lands in gpurun_out/.  Usage: bash tools/run_gpu_round.sh [stages...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${@:-model smoke bench launches ncu}"
for st in $STAGES; do
  case $st in
    kernels)
      bash tools/run_gpu_checks.sh ;;
    model)
      timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --timeout 600 \
          --timeout-method=thread --tb=short > gpurun_out/model_tests.log 2>&1
      echo "== model tests rc=$? :: $(tail -n 1 gpurun_out/model_tests.log)"
      grep -E "^(FAILED|ERROR|E  )" gpurun_out/model_tests.log | head -n 30 ;;
    smoke)
      timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
      echo "== smoke rc=$? :: $(tail -n 2 gpurun_out/smoke.log)" ;;
    bench)
      timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "== bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "== bench ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json ;;
    launches)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none \
          --kernel-name-base demangled -k regex:vb:: -s 1450 -c 1500 --csv \
          --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 > gpurun_out/launches.log 2>&1
      echo "== launches rc=$? lines=$(wc -l < gpurun_out/launches.csv)" ;;
    ncu)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemv_tma_kernel -s 60 -c 4 \
          -f -o gpurun_out/prof_gemv python bench.py --profile --steps 1 > gpurun_out/ncu_gemv.log 2>&1
      echo "== ncu gemv rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null ;;
    ncu_gemm)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 3 \
          -f -o gpurun_out/prof_gemm python bench.py --profile --steps 1 > gpurun_out/ncu_gemm.log 2>&1
      echo "== ncu gemm rc=$?" ;;
    ncu_skinny)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 2 -c 1 \
          -f -o gpurun_out/prof_skinny python tools/profile_skinny.py > gpurun_out/ncu_skinny.log 2>&1
      echo "== ncu skinny rc=$?" ;;
    ncu_gemm_big)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 \
          -f -o gpurun_out/prof_gemm_big python tools/profile_gemm_big.py > gpurun_out/ncu_gemm_big.log 2>&1
      echo "== ncu gemm big rc=$?" ;;
    gemm_table)
      timeout 600 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1
      echo "== gemm table rc=$?"; tail -n 3 gpurun_out/bench_gemm.log | cut -c1-300 ;;
    ncu_fmha)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 10 -c 2 \
          -f -o gpurun_out/prof_fmha python bench.py --profile --steps 1 > gpurun_out/ncu_fmha.log 2>&1
      echo "== ncu fmha rc=$?" ;;
  esac
done
