timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python scripts/small_trace.py 2>&1 | grep "small trace" | tail -1 | cut -c1-330
timeout 250 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_n.json 2> gpurun_out/r02_bench_n1_n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_n.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"]["frac"])
print(d["aux"]["cfg2_pci_ids_once"]["device_us_parse_resolve_finalize_join"], d["aux"]["cfg2_pci_ids_once"]["e2e_us_host_text_to_rows"], d["aux"]["cfg3_classify"]["kernel_ms"])
PY
SIZES=65536x190 ITERS=3 timeout 100 python scripts/full_path_check.py 2>&1 | tail -3
