#!/usr/bin/env python
"""Row-sharded predict + the library's NCCL all-gather on W GPUs (run under torchrun; W = WORLD_SIZE).
Every rank classifies its block, gathers all labels through tcsdn_allgather_labels_u8 / tcsdn_allgather_labels, and checks them against a
full single-GPU predict of the same rows.  Prints 'gather ok' on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from traffic_classifier_sdn_b200 import from_spec
from traffic_classifier_sdn_b200.parallel import Communicator, predict_sharded
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = Communicator()
ok = True
for name, n in (("gnb", 1_000_003), ("forest", 200_001), ("gnb", 5)):
    w = bench.build_workload(name)
    est = from_spec(w["spec"])
    X = bench.synth_rows(n, w["d"], seed=77, device=dev)
    full = est.predict_indices(X)
    got = predict_sharded(est, X, gather=True, comm=comm)      # library communicator, labels as bytes on the wire
    per0 = -(-n // world)
    a0, b0 = min(n, rank * per0), min(n, (rank + 1) * per0)
    got32 = comm.allgather_labels(full[a0:b0].contiguous(), per0)[:n]   # the int32 wire format of the same gather
    ok &= bool(torch.equal(full, got32))
    got_torch = predict_sharded(est, X, gather=True)            # torch.distributed path, same answer
    mine = predict_sharded(est, X, gather=False)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(full, got)) and bool(torch.equal(full, got_torch)) and got.numel() == n
    per = -(-n // world)
    ok &= bool(torch.equal(mine, full[min(n, rank * per):min(n, (rank + 1) * per)]))
# the exchange fused into the classification: labels of all ranks land in every rank's peer-memory buffer
for name, n in (("gnb", 1_000_000), ("forest", 50_001)):
    w = bench.build_workload(name)
    est = from_spec(w["spec"])
    per = -(-n // world)
    X = bench.synth_rows(per * world, w["d"], seed=99, device=dev)[: n]
    full = est.predict_indices(X)
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    comm.gather_buffer(per)                     # collective: the same block size on every rank
    for rep in range(3):
        got = comm.predict_gathered(est, X[a:b])
        torch.cuda.synchronize()
        flat = torch.cat([got[r, : min(per, max(0, n - r * per))] for r in range(world)]).to(torch.int32)
        ok &= bool(torch.equal(flat, full))
        pad_ok = all(bool((got[r, min(per, max(0, n - r * per)):] == 255).all()) for r in range(world))
        ok &= pad_ok
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
comm.close()
if rank == 0:
    print("gather ok" if int(flag.item()) == 1 else "gather MISMATCH", "world", world)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
