#!/bin/bash
# Multi-GPU session (gpurun --gpus N): SP parity test, SP prefill bench at 1..N ranks, replica bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
FRAMES=${2:-64}
timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider --timeout 600 --timeout-method=thread \
    --tb=short > gpurun_out/sp_tests.log 2>&1
echo "== sp tests rc=$? :: $(tail -n 1 gpurun_out/sp_tests.log)"; grep -E "^(FAILED|ERROR|E  )" gpurun_out/sp_tests.log | head -n 20
for n in 1 $N; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
      bench.py --gpus $n --workload sp_prefill --frames $FRAMES --steps 3 --warmup 3 > gpurun_out/bench_sp_n$n.json 2> gpurun_out/bench_sp_n$n.err
  echo "== sp bench n=$n rc=$?"; tail -c 1500 gpurun_out/bench_sp_n$n.json; tail -n 3 gpurun_out/bench_sp_n$n.err
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_replicas_n$N.json 2> gpurun_out/bench_replicas_n$N.err
echo "== replica bench n=$N rc=$?"; tail -c 600 gpurun_out/bench_replicas_n$N.json; tail -n 3 gpurun_out/bench_replicas_n$N.err
