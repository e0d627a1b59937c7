"""Worker for the world_size-2 gloo test: the host-side multi-rank plumbing of bench.py
(rendezvous, NCCL-id broadcast, env sharding, max-over-ranks reduction) without any GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import dependence_free_rl_b200 as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    obj = [D.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    nccl_id = obj[0]
    assert isinstance(nccl_id, bytes) and len(nccl_id) == 128
    # every rank must hold the same id
    t = torch.tensor(list(nccl_id), dtype=torch.int64)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    # env sharding: contiguous, disjoint, covering
    per = 1000
    off = torch.tensor([rank * per], dtype=torch.int64)
    offs = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(offs, off)
    assert [int(o) for o in offs] == [r * per for r in range(world)]
    # max-over-ranks timing reduction
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert float(ms) == 10.0 + world - 1
    # without a GPU the compute path must fail loudly, never fall back to the CPU
    failed = False
    try:
        D.Context(0, world, rank, nccl_id)
    except D._lib.DfrlError as e:
        failed = "no CUDA device" in str(e) or "sm_" in str(e)
    if not torch.cuda.is_available():
        assert failed
    dist.barrier()
    if rank == 0:
        print(json.dumps({"ok": True, "world": world}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
