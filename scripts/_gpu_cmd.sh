timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python scripts/small_trace.py 2>&1 | grep "small trace" | tail -2 | cut -c1-340
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_q.json 2> gpurun_out/r02_bench_n1_q.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_q.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"]["frac"])
c=d["aux"]["cfg2_pci_ids_once"]; print(c["device_us_parse_resolve_finalize_join"], c["e2e_us_host_text_to_rows"], c["kernel_us_inside_e2e_call"])
PY
tail -3 gpurun_out/r02_bench_n1_q.err
