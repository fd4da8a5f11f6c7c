#!/bin/bash
# Round-2 GPU session driver: stages write logs into gpurun_out/.  Usage: bash tools/run_gpu_r2.sh [stages...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${@:-tests smoke bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
for st in $STAGES; do
  case $st in
    tests)
      rm -f gpurun_out/parity_report.jsonl
      timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --timeout-method=thread \
          --tb=short -x > gpurun_out/gpu_tests.log 2>&1
      echo "== gpu tests rc=$? :: $(tail -n 1 gpurun_out/gpu_tests.log)"
      grep -E "^(FAILED|ERROR|E  )" gpurun_out/gpu_tests.log | head -n 40 ;;
    tests_all)
      rm -f gpurun_out/parity_report.jsonl
      timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --timeout-method=thread \
          --tb=short > gpurun_out/gpu_tests.log 2>&1
      echo "== gpu tests rc=$? :: $(tail -n 1 gpurun_out/gpu_tests.log)"
      grep -E "^(FAILED|ERROR|E  )" gpurun_out/gpu_tests.log | head -n 60 ;;
    newtests)
      rm -f gpurun_out/parity_report.jsonl
      timeout 2400 python -m pytest tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider \
          -k "fmha2 or decode_attention or cfg or generate_content" --timeout 900 --timeout-method=thread --tb=short \
          > gpurun_out/gpu_newtests.log 2>&1
      echo "== new gpu tests rc=$? :: $(tail -n 1 gpurun_out/gpu_newtests.log)"
      grep -E "^(FAILED|ERROR|E  )" gpurun_out/gpu_newtests.log | head -n 60 ;;
    sptests)
      timeout 1200 python -m pytest tests/test_sp_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 \
          --timeout-method=thread --tb=short > gpurun_out/sp_tests.log 2>&1
      echo "== sp tests rc=$? :: $(tail -n 1 gpurun_out/sp_tests.log)"
      grep -E "^(FAILED|ERROR|E  )" gpurun_out/sp_tests.log | head -n 30 ;;
    smoke)
      timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
      echo "== smoke rc=$? :: $(tail -n 2 gpurun_out/smoke.log)" ;;
    bench)
      timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "== bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    bench_quick)
      timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-sp --no-video > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
      echo "== bench quick rc=$?"; tail -c 4000 gpurun_out/bench_quick.json; tail -n 5 gpurun_out/bench_quick.err ;;
    benchN)
      N=$(nvidia-smi -L | wc -l)
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
          --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
      echo "== bench N=$N rc=$?"; tail -c 5000 gpurun_out/bench_n$N.json; tail -n 5 gpurun_out/bench_n$N.err ;;
    benchref)
      timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "== bench ref rc=$?"; tail -c 2500 gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err ;;
    benchrefgpu)
      timeout 900 python bench.py --impl reference_gpu --steps 3 --warmup 2 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err
      echo "== bench ref gpu rc=$?"; tail -c 2500 gpurun_out/bench_refgpu.json; tail -n 3 gpurun_out/bench_refgpu.err ;;
    launches)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none \
          --kernel-name-base demangled -k regex:vb:: -s 1450 -c 1500 --csv \
          --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 > gpurun_out/launches.log 2>&1
      echo "== launches rc=$? lines=$(wc -l < gpurun_out/launches.csv)" ;;
    ledger)
      # ncu --set full of one launch per hot kernel; the raw page is exported on the box (csv, small) and
      # the report itself kept only if it fits the 64 MiB return budget
      timeout 1500 ncu --set full --clock-control none --kernel-name-base demangled \
          --profile-from-start off -k regex:vb:: -f -o /tmp/ledger python tools/ncu_ledger.py > gpurun_out/ledger.log 2>&1
      echo "== ledger rc=$?"; ls -la /tmp/ledger.ncu-rep
      ncu -i /tmp/ledger.ncu-rep --page raw --csv > gpurun_out/ledger_raw.csv 2>/dev/null
      sz=$(stat -c %s /tmp/ledger.ncu-rep 2>/dev/null || echo 0)
      if [ "$sz" -lt 30000000 ]; then cp /tmp/ledger.ncu-rep gpurun_out/ledger.ncu-rep; fi
      wc -c gpurun_out/ledger_raw.csv ;;
    ncu_one)
      # full capture WITH source of one kernel: NCU_K=<regex> NCU_SCRIPT=<tools/x.py> NCU_ARGS=...
      timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
          --profile-from-start off -k "regex:${NCU_K}" -c ${NCU_C:-1} -f -o gpurun_out/prof_${NCU_NAME:-one} \
          python ${NCU_SCRIPT} ${NCU_ARGS} > gpurun_out/ncu_${NCU_NAME:-one}.log 2>&1
      echo "== ncu_one rc=$?"; ls -la gpurun_out/prof_${NCU_NAME:-one}.ncu-rep ;;
    *)
      if [ -f "tools/$st" ]; then
        timeout 1200 python "tools/$st" > "gpurun_out/${st%.py}.log" 2>&1
        echo "== $st rc=$? :: $(tail -n 3 gpurun_out/${st%.py}.log)"
      else
        echo "unknown stage $st"
      fi ;;
  esac
done
