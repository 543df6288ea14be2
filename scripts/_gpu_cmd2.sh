timeout 400 python -m pytest tests/test_sharding.py tests/test_gpu_pciids.py -m gpu -x -q 2>&1 | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2_k.json 2> gpurun_out/r02_bench_n2_k.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n2_k.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_checked"], d["e2e"]["value"])
w=d["weak_scaling"]; print("weak", w["value"], w["ms_per_step"], w["kernel_ms"], w["e2e"]["value"])
PY
tail -3 gpurun_out/r02_bench_n2_k.err
KXPU_TRACE_MERGE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "shard trace" | tail -2
