#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU).  Reports land in gpurun_out/, summaries are made
# on the build box with scripts/ncu_summary.py and committed under profiles/.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
cd "$(dirname "$0")/.."
# aux workloads: classify kernels, fused CDI emit (JSON, YAML), the small-text cooperative kernel
ROUNDS=2 timeout 300 $NCU -k regex:'k_cdi_fused|k_candidates|k_accept_scan|k_groups|k_devfirst_scan|k_pairs|k_onesweep|k_bounds|small_load' \
  --launch-skip 13 -c 12 -o gpurun_out/r02_aux python scripts/aux_launches.py > gpurun_out/r02_ncu_aux.log 2>&1
# text without repeated blocks: every kernel of the load on its full path (steady state: third load or later)
SIZES=65536x190 ITERS=6 NO_ORACLE=1 timeout 300 $NCU -k regex:'parse_kernel_v5|resolve_ranges|resolve_chunks|select_finalize' \
  --launch-skip 16 -c 4 -o gpurun_out/r02_allalive python scripts/full_path_check.py > gpurun_out/r02_ncu_allalive.log 2>&1
# cfg4: the kernels behind the parse
timeout 300 $NCU -k regex:'resolve_ranges|resolve_chunks|select_finalize|lookup_kernel' \
  --launch-skip 40 -c 4 -o gpurun_out/r02_cfg4_tail python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_cfg4.log 2>&1
# launch lists (gpu__time_duration only)
ROUNDS=3 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_aux_launches.csv python scripts/aux_launches.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
