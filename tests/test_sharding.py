"""N > 1 path: shard planning (CPU), the shard/merge protocol over gloo with world_size 2 (CPU),
and the real NCCL path when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def shard_candidates(text, base):
    """What one rank contributes (pure Python restatement of the parse kernel's fold rule): for every
    key the smallest (line, anchor) over ALL blocks of the shard, and per vendor the first anchor."""
    cand, vfirst = {}, {}
    cur = None  # (vendor or None, anchor offset)
    pos = 0
    for line in text.split(b"\n"):
        off = pos
        pos += len(line) + 1
        if off >= len(text):
            break
        if line.endswith(b"\r"):
            line = line[:-1]
        if line.startswith(b"#"):
            continue
        hexok = lambda s: len(s) == 4 and all(c in b"0123456789abcdef" for c in s)
        if line.startswith(b"\t"):
            if cur and cur[0] is not None and hexok(line[1:5]):
                k = (cur[0] << 16) | int(line[1:5], 16)
                v = (base + off, cur[1])
                if k not in cand or v < cand[k]:
                    cand[k] = v
            continue
        v = int(line[:4], 16) if hexok(line[:4]) else None
        cur = (v, base + off)
        if v is not None and v not in vfirst:
            vfirst[v] = base + off
    return cand, vfirst


def merge(parts):
    cand, vfirst = {}, {}
    for c, vf in parts:
        for k, v in c.items():
            if k not in cand or v < cand[k]:
                cand[k] = v
        for v, o in vf.items():
            vfirst[v] = min(o, vfirst.get(v, o))
    rows = sorted((line, k) for k, (line, anchor) in cand.items() if vfirst.get(k >> 16) == anchor)
    return np.array([k for _, k in rows], np.uint32), np.array([l for l, _ in rows], np.uint64)


def test_plan_shards_cuts_at_vendor_lines(pci_text):
    from kxpu_b200.sharding import plan_shards
    for n in (1, 2, 3, 4, 8):
        sh = plan_shards(pci_text, n)
        assert sh[0][0] == 0 and sh[-1][1] == len(pci_text)
        for (a, b), (c, d) in zip(sh, sh[1:]):
            assert b == c
        for a, _ in sh[1:]:
            assert pci_text[a - 1:a] == b"\n" and pci_text[a:a + 1] not in (b"\t", b"#")
    assert plan_shards(b"", 4) == [(0, 0)] * 4
    assert plan_shards(b"\tonly\n\tdevice lines\n", 2) == [(0, 20), (20, 20)]


def test_shard_merge_equals_whole_text(oracle, pci_text):
    from kxpu_b200.sharding import plan_shards
    text = pci_text[:300000]
    text = text[:text.rfind(b"\n") + 1] * 2
    want = oracle.table_build(text)
    for n in (1, 2, 5):
        parts = [shard_candidates(text[a:b], a) for a, b in plan_shards(text, n)]
        k, o = merge(parts)
        assert np.array_equal(k, want["key"]) and np.array_equal(o, want["line_off"])


def _gloo_worker(rank, world, port, text, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import kxpu_b200  # noqa: F401
    from kxpu_b200.sharding import plan_shards
    a, b = plan_shards(text, world)[rank]
    mine = shard_candidates(text[a:b], a)
    parts = [None] * world
    dist.all_gather_object(parts, mine)  # the one exchange of the path
    k, o = merge(parts)
    q.put((rank, k.tobytes(), o.tobytes()))
    dist.destroy_process_group()


def test_world_size_2_gloo(oracle, pci_text):
    """One process per rank over gloo: both ranks end with the table of the whole text."""
    import torch.multiprocessing as mp
    text = pci_text[:200000]
    text = text[:text.rfind(b"\n") + 1] + b"10de  dup\n\t2330  not the first\n" + pci_text[300000:400000]
    want = oracle.table_build(text)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, text, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, kb, ob in got:
        assert np.array_equal(np.frombuffer(kb, np.uint32), want["key"])
        assert np.array_equal(np.frombuffer(ob, np.uint64), want["line_off"])


@pytest.mark.gpu
def test_two_gpu_nccl_sharded_load():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "scripts", "multi_rank_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_RANK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
