"""World-size-2 gloo test of the data-parallel host logic (causal_vae_b200/parallel.py), with the
oracle as the per-rank compute: sharding, replica broadcast, SUM all-reduce of the flat gradient,
and that every rank ends the step with identical parameters."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from causal_vae_b200 import parallel
    from oracle import cvae_oracle as O
    H = W = 64
    B = 4
    # rank 1 starts from different weights: the broadcast must make the replicas identical
    P = O.fill_state_dict(O.vessel_shapes(H, W), seed=rank)
    holder = torch.nn.Module()
    names = []
    for k, v in P.items():
        nm = k.replace(".", "__")
        names.append((k, nm))
        if v.is_floating_point() and "running" not in k:
            holder.register_parameter(nm, torch.nn.Parameter(v))
        else:
            holder.register_buffer(nm, v)
    parallel.broadcast_module(holder, src=0)
    P = {k: getattr(holder, nm).data for k, nm in names}
    ref0 = O.fill_state_dict(O.vessel_shapes(H, W), seed=0)
    assert all(torch.equal(P[k], ref0[k]) for k in P)

    x, m, t, eps = O.vessel_inputs(B, H, W, seed=0)
    xs, ms, ts, es = parallel.shard_batch([x, m, t, eps], rank, world)
    assert xs.shape[0] == B // world
    _, grads, _ = O.vessel_train_step({k: v.clone() for k, v in P.items()}, {}, 1, xs, ms, ts, es)
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in keys])
    local = flat.clone()
    parallel.allreduce_gradients(flat)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat, sum(gathered), rtol=0, atol=0)          # SUM, not MEAN
    mean = local.clone()
    parallel.allreduce_gradients(mean, mean=True)
    assert torch.allclose(mean, flat / world)
    if rank == 0:
        torch.save({"flat": flat, "keys": keys, "shard0": local}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_gloo_gradient_exchange(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # the summed shard gradients are not the single-device gradient of the concatenated batch
    # (BatchNorm statistics and pos_weight are per shard), but they are the sum of two valid
    # per-shard reference steps: recompute shard 0 here and check it is one of the addends.
    sys.path.insert(0, ROOT)
    from causal_vae_b200 import parallel
    from oracle import cvae_oracle as O
    P = O.fill_state_dict(O.vessel_shapes(64, 64), seed=0)
    x, m, t, eps = O.vessel_inputs(4, 64, 64, seed=0)
    xs, ms, ts, es = parallel.shard_batch([x, m, t, eps], 0, 2)
    _, g0, _ = O.vessel_train_step(P, {}, 1, xs, ms, ts, es)
    flat0 = torch.cat([g0[k].reshape(-1) for k in got["keys"]])
    # (different thread counts reorder the fp32 reductions: compare at 1e-4 of max |g|)
    assert (flat0 - got["shard0"]).abs().max() <= 1e-4 * flat0.abs().max()
    assert (got["flat"] - flat0).abs().max() > 0


def test_shard_bounds_cover_batch():
    from causal_vae_b200 import parallel
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
