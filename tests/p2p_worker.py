"""2-rank GPU worker (torchrun): the fused peer-memory gradient exchange must give bit-identical
parameters on both ranks and the same parameters as the NCCL path within fp32 rounding."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dependence_free_rl_b200 as D  # noqa: E402


CONV = os.environ.get("P2P_NETS", "dense") == "conv"   # the reference's conv1d policy on the table path (conv_table.cuh)


def run(ctx, rank, world, n, iters):
    T = 4
    if CONV:
        policy = D.Model(ctx, D.conv_layers([4, 128, 64, 1], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 32, 1]), 32)
    else:
        policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
    policy.init_parameters(1234)
    value.init_parameters(1235)
    env = D.Environment(ctx, n, seed=1234, env_offset=rank * n)
    rows = n * world * T
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-4 * 32 / rows, value_lr=1e-5 * 32 / rows)
    tr.iterate(iters)
    out = (policy.parameters().copy(), value.parameters().copy(), tr.stats())
    tr.close(); env.close(); policy.close(); value.close()
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the result line
    dist.init_process_group("gloo", rank=rank, world_size=world)
    obj = [D.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx = D.Context(int(os.environ.get("LOCAL_RANK", rank)), world, rank, obj[0])
    n, iters = (16384, 4) if CONV else (4096, 6)
    p_nccl, v_nccl, _ = run(ctx, rank, world, n, iters)       # NCCL all-reduce + optimizer kernel
    handles = [None] * world
    dist.all_gather_object(handles, ctx.p2p_export())
    ctx.p2p_attach(handles)
    assert ctx.p2p_attached()
    p_p2p, v_p2p, stats = run(ctx, rank, world, n, iters)     # fused peer-memory exchange
    allp = [None] * world
    dist.all_gather_object(allp, (p_p2p.tobytes(), v_p2p.tobytes()))
    ok_identical = all(a == allp[0] for a in allp)
    scale = np.max(np.abs(p_nccl))
    err = float(np.max(np.abs(p_p2p - p_nccl)) / scale)
    verr = float(np.max(np.abs(v_p2p - v_nccl)) / np.max(np.abs(v_nccl)))
    moved = bool(np.any(p_p2p != D_init(ctx)))
    if rank == 0:
        print(json.dumps({"ok": bool(ok_identical and err < 1e-5 and verr < 1e-5 and moved), "identical_across_ranks": ok_identical,
                          "p2p_vs_nccl_rel": err, "value_rel": verr, "env_steps": stats["env_steps"]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def D_init(ctx):
    m = D.Model(ctx, D.conv_layers([4, 128, 64, 1], D.SOFTMAX) if CONV else D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
    m.init_parameters(1234)
    p = m.parameters().copy()
    m.close()
    return p


if __name__ == "__main__":
    main()
