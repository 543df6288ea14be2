timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; tail -c 600 gpurun_out/r02_bench_reference_final.json
timeout 400 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 1500 gpurun_out/r02_bench_n1_final.json; tail -3 gpurun_out/r02_bench_n1_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --settle-ms 50 > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck.log python scripts/sanitizer_check.py > gpurun_out/r02_sanitizer_run.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log; tail -2 gpurun_out/r02_sanitizer_run.log
