"""Multi-GPU check of the peer-memory exchange (gaitk_p2p_allreduce) against the NCCL all-reduce path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_p2p_check.py

Every rank trains two copies of the model on its shard of the same global batches -- one exchanging gbuf with NCCL,
one with the fused peer-memory kernel, eagerly and through CUDA graphs -- and checks that (a) the two exchanges give the
same parameters up to fp32 summation order, (b) with the peer-memory exchange all ranks hold bit-identical parameters.
Not collected by pytest (needs >= 2 GPUs); tests/test_gpu_parity.py launches it when two devices are visible."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))


def main():
    import gaitk
    import gait_oracle as O
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Bg = 64 * world
    models, steps = {}, {}
    for mode in ("nccl", "p2p", "p2p_graph", "single"):
        torch.manual_seed(0)
        m = gaitk.WearGaitThreeModal().cuda()
        crit = [gaitk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
        steps[mode] = gaitk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gaitk.DTYPE_F32,
                                           p2p=mode in ("p2p", "p2p_graph"), use_graph=mode == "p2p_graph",
                                           process_group=False if mode == "single" else None)
        models[mode] = m
    batches = []
    for i in range(2):
        xs, y = O.synth_weargait_batch(Bg, seed=100 + i)
        sl = slice(rank * 64, (rank + 1) * 64)
        batches.append(([torch.from_numpy(x[sl]).cuda() for x in xs], torch.from_numpy(y[sl]).cuda(), torch.from_numpy(y).cuda(),
                        [torch.from_numpy(x).cuda() for x in xs]))
    for it in range(9):
        xs, y, yg, xs_all = batches[it % 2]
        for mode in ("nccl", "p2p", "p2p_graph"):
            loss, _ = steps[mode].step(xs, [y] * 3, ys_global=[yg] * 3)
            if mode != "nccl":
                assert not steps[mode].exchange_failed(), "peer did not arrive"
        steps["single"].step(xs_all, [yg] * 3)                # the whole global batch on one device, no exchange
    torch.cuda.synchronize()
    flat = {k: models[k].flat_params().detach().clone() for k in models}
    # N ranks x B/N windows == one rank x B windows: same parameters up to fp32 summation order
    e1 = float((flat["nccl"] - flat["single"]).abs().max())
    assert e1 <= 5e-6 * float(flat["single"].abs().max()) + 1e-7, ("sharded vs single-device global batch", e1)
    err = float((flat["nccl"] - flat["p2p"]).abs().max()); scale = float(flat["nccl"].abs().max())
    assert err <= 2e-6 * scale + 1e-7, ("p2p vs nccl", err, scale)
    assert torch.equal(flat["p2p"], flat["p2p_graph"]), "graph replay differs from eager"
    gathered = [torch.empty_like(flat["p2p"]) for _ in range(world)]
    dist.all_gather(gathered, flat["p2p"])
    for r in range(world):
        assert torch.equal(gathered[r], gathered[0]), f"rank {r} diverged from rank 0"
    l_n, _ = steps["nccl"].stats(); l_p, _ = steps["p2p"].stats()
    assert torch.allclose(l_n, l_p, rtol=1e-5, atol=1e-6), (l_n, l_p)
    # the consistency-coupled FoG step shards the same way (KL 'batchmean' over the GLOBAL batch)
    fog = {}
    sk_all, se_all, y_all = O.synth_fog_batch(32 * world, seed=7)
    for mode in ("dp", "single"):
        torch.manual_seed(1)
        m = gaitk.MultiModalMultiTaskModel(21, 6, 6, 6, 426, 16, 8, 128, 3, synchronized_loading=True).cuda()
        crit = [gaitk.GCLLoss(cls_num_list=c, m=0.2, s=25.0, noise_mul=0.0) for c in ([50, 30, 20], [45, 33, 22])]
        st = gaitk.FusedTrainStep(m, crit, cagrad_c=0.1, private_mult=1.0, consistency_lambda=1.0,
                                  process_group=False if mode == "single" else None)
        yg = torch.from_numpy(y_all).cuda()
        for it in range(3):
            if mode == "dp":
                sl = slice(rank * 32, (rank + 1) * 32)
                st.step([torch.from_numpy(sk_all[sl]).cuda(), torch.from_numpy(se_all[sl]).cuda()], [yg[sl].contiguous()] * 2, ys_global=[yg, yg])
            else:
                st.step([torch.from_numpy(sk_all).cuda(), torch.from_numpy(se_all).cuda()], [yg, yg])
        fog[mode] = m.flat_params().detach().clone()
    e2 = float((fog["dp"] - fog["single"]).abs().max())
    assert e2 <= 5e-6 * float(fog["single"].abs().max()) + 1e-7, ("FoG consistency step: sharded vs single", e2)
    if rank == 0:
        print(f"P2P_CHECK_OK world={world} max|nccl-p2p|={err:.2e} max|sharded-single|={e1:.2e} fog_sync max|sharded-single|={e2:.2e} "
              f"losses={l_p.cpu().tolist()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
