"""world_size-2 gloo test (CPU) of the data-parallel logic: shards with global loss denominators, summed by
all-reduce, reproduce the single-process full-batch gradients (G per task and private grads) -- the property the
NCCL path relies on.  The per-shard gradients come from the CPU oracle; the sharding, the
denominators and the reduction go through the package's dist helpers and torch.distributed."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _grads(p, xs, ys, denoms, cls_w, O):
    """unnormalised-by-shard oracle gradients: loss_k = sum_shard w nll / denom_global"""
    logits = O.weargait_forward(O._HeadAlias(p), *xs)
    losses = []
    for lg, y, d, w in zip(logits, ys, denoms, cls_w):
        z = 25.0 * torch.where(torch.nn.functional.one_hot(y, 2).bool(), lg - 0.2, lg)
        nll = -torch.log_softmax(z, 1).gather(1, y.view(-1, 1)).squeeze(1)
        losses.append((w[y] * nll).sum() / d)
    keys = list(p.keys()); leaves = [p[k] for k in keys]
    out = []
    for L in losses:
        gs = torch.autograd.grad(L, leaves, retain_graph=True, allow_unused=True)
        out.append(torch.cat([(torch.zeros_like(v) if g is None else g).reshape(-1) for v, g in zip(leaves, gs)]))
    return torch.stack(out), torch.stack([l.detach() for l in losses])


def _worker(rank, world, initfile, result_dir):
    for q in (ROOT, ROOT / "oracle"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import torch.distributed as dist
    import gait_oracle as O
    import gaitk
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = gaitk.WearGaitThreeModal()
    p = O.canonical_params({k: v.detach().numpy().copy() for k, v in m.state_dict().items()}, True)
    B = 24
    xs, y = O.synth_weargait_batch(B, seed=3)
    xs = [torch.from_numpy(x) for x in xs]; y = torch.from_numpy(y)
    cls_w = [torch.tensor([1.7, 0.3])] * 3
    denoms = gaitk.dist.global_denominators([y] * 3, cls_w)           # from the GLOBAL labels: no communication
    lo, hi = gaitk.dist.shard_bounds(B, rank, world)
    g, l = _grads(p, [x[lo:hi] for x in xs], [y[lo:hi]] * 3, denoms, cls_w, O)
    buf = torch.cat([g.reshape(-1), l])
    gaitk.dist.all_reduce_sum_(buf)
    if rank == 0:
        full_g, full_l = _grads(p, xs, [y] * 3, denoms, cls_w, O)
        np.save(os.path.join(result_dir, "sharded.npy"), buf.numpy())
        np.save(os.path.join(result_dir, "full.npy"), torch.cat([full_g.reshape(-1), full_l]).numpy())
    dist.destroy_process_group()


def test_two_rank_shards_add_up_gloo():
    with tempfile.TemporaryDirectory() as td:
        initfile = os.path.join(td, "init")
        mp.spawn(_worker, args=(2, initfile, td), nprocs=2, join=True)
        a = np.load(os.path.join(td, "sharded.npy")); b = np.load(os.path.join(td, "full.npy"))
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()


def test_shard_bounds():
    import gaitk
    assert gaitk.dist.shard_bounds(10, 0, 4) == (0, 2) and gaitk.dist.shard_bounds(10, 3, 4) == (6, 8)
    assert gaitk.dist.shard_bounds(32768, 7, 8) == (28672, 32768)
