#!/bin/bash
# The measurement commands behind profiles/r02_* (run on a GPU box from the repo root; outputs to gpurun_out/).
#   bash scripts/measure_r02.sh n1            tests, smoke, reference arm, default bench, launch list   (1 GPU)
#   bash scripts/measure_r02.sh ngpu N [trace] bench under torchrun on N GPUs (+ KXPU_TRACE_MERGE phase trace)
#   bash scripts/measure_r02.sh sanitize      memcheck / racecheck / synccheck of scripts/sanitizer_check.py, GPU tests under memcheck
#   bash scripts/profile_r02.sh               ncu --set full captures (separate script)
set -x
mkdir -p gpurun_out
case "$1" in
n1)
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu.log
  timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json
  timeout 400 python bench.py > gpurun_out/r02_bench_n1.json
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --settle-ms 50 > /dev/null 2>&1
  KXPU_TRACE_SMALL=1 timeout 60 python scripts/small_trace.py 2>&1 | tail -4 > gpurun_out/r02_small_trace.txt
  SIZES=40000x24,65536x190 ITERS=3 timeout 200 python scripts/full_path_check.py > gpurun_out/r02_full_path.log 2>&1
  ;;
ngpu)
  N=$2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
  if [ "$3" = "trace" ]; then
    KXPU_TRACE_MERGE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-second-mode --no-aux --settle-ms 100 2>&1 >/dev/null | grep "shard trace" > gpurun_out/r02_shard_trace_n${N}.txt
  fi
  ;;
sanitize)
  for tool in memcheck racecheck synccheck; do
    COPIES=2 timeout 400 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python scripts/sanitizer_check.py > gpurun_out/r02_sanitizer_${tool}_run.log 2>&1
    tail -2 gpurun_out/r02_sanitizer_$tool.log
  done
  timeout 800 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_pytest_memcheck.log python -m pytest \
      tests/test_gpu_pciids.py tests/test_sharding.py tests/test_gpu_full.py tests/test_gpu_discovery.py -m gpu -x -q \
      -k "not x1000 and not cfg3_one_million and not hypothesis" > gpurun_out/r02_sanitizer_pytest_run.log 2>&1
  tail -3 gpurun_out/r02_sanitizer_pytest_memcheck.log gpurun_out/r02_sanitizer_pytest_run.log
  ;;
*) echo "usage: $0 n1 | ngpu N [trace] | sanitize"; exit 2 ;;
esac
