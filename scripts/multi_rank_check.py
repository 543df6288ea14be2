"""Multi-GPU parity check (run under torchrun, one rank per GPU): sharded load == oracle on the whole text."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import kxpu_b200 as K
from kxpu_b200 import workloads as W
from kxpu_b200.sharding import plan_shards

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
    tab = kx.pciids_load_sharded(d, n, s)
    keys, offs, rows = kx.table_export(tab)
    names, _, _ = kx.names(tab, rows[:300])
    res = None
    if rank == 0:
        from oracle import oracle as O
        orows = O.table_build(text)
        same = np.array_equal(keys, orows["key"]) and np.array_equal(offs, orows["line_off"])
        nm_ok = all(nm == O.row_name(text, int(o)) for nm, o in zip(names, offs[:300]))
        print("%-70s shards %s rows %d table %s names %s merge_ms %.3f" % (name, shards, tab.rows, same, nm_ok, kx.timings()[K.binding.T_MERGE]))
        ok = ok and same and nm_ok
    # every rank holds the same table
    h = torch.tensor([int(np.bitwise_xor.reduce(keys.astype(np.uint64) * 31 + offs)) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
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
