"""kata-xpu-device-plugin_b200 -- B200-native discovery hot path of the Kata xPU device plugin.

This package holds only what the hot path needs:
  csrc/      CUDA kernels for sm_100a + the C ABI (include/kxpu.h) -> lib/libkxpu.so
  host/      host-side mirror of the reference's discovery / CDI / Allocate logic (C++)
  binding.py ctypes binding of the C ABI (what the Go cgo shim of INTEGRATION.md does)
  sharding.py the shard-planning rule in Python (the product planner is kxpu_plan_shards behind the ABI;
             tests hold the two equal)
  workloads.py synthetic inputs of BASELINE.json configs[0..4]

The directory name contains '-', so import it through the repo-root shim `kxpu_b200`.
There is NO CPU fallback: if lib/libkxpu.so or a B200 is missing, every compute call
raises KxpuError.
"""
from .binding import Kxpu, KxpuError, KxpuMulti, lib_path, load_library, plan_shards  # noqa: F401
