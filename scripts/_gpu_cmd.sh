timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_o.json 2> gpurun_out/r02_bench_n1_o.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_o.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"]["frac"])
print(d["aux"]["cfg2_pci_ids_once"])
PY
tail -3 gpurun_out/r02_bench_n1_o.err
