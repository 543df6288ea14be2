"""Multi-GPU parity check (run under torchrun, one rank per GPU): sharded load + join == oracle on the whole text."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kxpu_b200 as K
from kxpu_b200 import workloads as W
plan_shards = K.plan_shards

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
kx = K.Kxpu(lr)
uid = torch.zeros(K.binding.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(kx.comm_unique_id()))
dist.broadcast(uid, 0)
kx.comm_init(world, rank, uid.cpu().numpy())

base = W.load_pci_ids()
texts = {
    "pci.ids": base,
    "pci.ids x3 (first occurrence wins across shards)": base * 3,
    "reversed halves (later shard holds the earlier vendors' duplicates)": base[700000:base.find(b"\n", 1400000) + 1] + base,
    "tiny": b"10de  NV\n\t0001  a\n10df  x\n\t0002  b\n" * 3,
}
ok = True
for name, text in texts.items():
    shards = plan_shards(text, world)
    s, e = shards[rank]
    n = e - s
    d = kx.dev_alloc(max(n, 16))
    if n:
        kx.upload(d, np.frombuffer(text[s:e], np.uint8))
    # the join: every rank probes its slice of the same 4096 keys, all hits land on every rank
    from oracle import oracle as O
    orows = O.table_build(text)
    q = W.make_queries(orows["key"], 4096, 7) if len(orows) > 64 else np.array([0x10de0001, 0x10df0002, 3, 4] * world, np.uint32)
    per = len(q) // world
    dq = kx.dev_alloc(max(per * 4, 16))
    kx.upload(dq, q[rank * per:(rank + 1) * per])
    dr = kx.dev_alloc(len(q) * 4)
    tab = kx.pciids_join_sharded(d, n, s, dq, per, rank * per, per * world, dr)
    keys, offs, rows = kx.table_export(tab)
    got = kx.download(dr, per * world * 4, np.int32)
    line_of_row = np.full(tab.rows + 1, -1, np.int64)
    line_of_row[rows] = offs.astype(np.int64)
    got_line = np.where(got >= 0, line_of_row[np.maximum(got, 0)], -1)
    want_line = np.array([O.device_name(text, int(k))[0] for k in q[:per * world][:256]], np.int64)
    join_ok = np.array_equal(got_line[:256], want_line)
    tab2 = kx.pciids_load_sharded(d, n, s)  # load only (no join phase)
    k2, o2, _ = kx.table_export(tab2)
    join_ok = join_ok and np.array_equal(k2, keys) and np.array_equal(o2, offs)
    tab2.free()
    kx.dev_free(dq); kx.dev_free(dr)
    names, _, _ = kx.names(tab, rows[:300])
    res = None
    if rank == 0:
        same = np.array_equal(keys, orows["key"]) and np.array_equal(offs, orows["line_off"])
        nm_ok = all(nm == O.row_name(text, int(o)) for nm, o in zip(names, offs[:300]))
        print("%-70s shards %s rows %d table %s names %s join %s" % (name, shards, tab.rows, same, nm_ok, join_ok))
        ok = ok and same and nm_ok and join_ok
    # every rank holds the same table
    hv = int(np.bitwise_xor.reduce(keys.astype(np.uint64) * 31 + offs)) if len(keys) else 0
    hv ^= int(np.bitwise_xor.reduce(got.astype(np.int64).view(np.uint64) * np.arange(1, len(got) + 1, dtype=np.uint64)))
    h = torch.tensor([(hv & 0x7FFFFFFFFFFFFFFF) if join_ok else -1 - rank], dtype=torch.int64, device="cuda")
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    if rank == 0:
        same_all = all(int(x.item()) == int(hs[0].item()) for x in hs)
        print("   identical on all ranks:", same_all)
        ok = ok and same_all
    tab.free()
    kx.dev_free(d)
kx.comm_destroy()
dist.destroy_process_group()
if rank == 0:
    print("MULTI_RANK_OK" if ok else "MULTI_RANK_FAIL")
    sys.exit(0 if ok else 1)
