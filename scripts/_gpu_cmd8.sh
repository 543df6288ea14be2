N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${N}_final.json 2> gpurun_out/r02_bench_n${N}_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n${N}_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_checked"], d["e2e"]["value"], d["limiting_stage"])
w=d["weak_scaling"]; print("weak", w["value"], w["ms_per_step"], w["kernel_ms"], w["e2e"]["value"])
PY
tail -2 gpurun_out/r02_bench_n${N}_final.err
[ "$2" = "trace" ] && KXPU_TRACE_MERGE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-second-mode --no-aux --settle-ms 100 2>&1 >/dev/null | grep "shard trace" | tail -$N > gpurun_out/r02_shard_trace_n${N}.txt
[ "$2" = "trace" ] && tail -3 gpurun_out/r02_shard_trace_n${N}.txt; true
