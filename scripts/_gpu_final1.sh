timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; tail -c 300 gpurun_out/r02_bench_reference_final.json
timeout 400 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
print(d["aux"]["cfg2_pci_ids_once"]); print(d["cpu_best"]["job_gbs_all_cores"], d["cpu_best"]["cores"], d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/r02_bench_n1_final.err
