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
    for mode in ("nccl", "p2p", "p2p_graph"):
        torch.manual_seed(0)
        m = gaitk.WearGaitThreeModal().cuda()
        crit = [gaitk.GCLLoss(cls_num_list=[40, 60], m=0.2, s=25.0, noise_mul=0.0) for _ in range(3)]
        steps[mode] = gaitk.FusedTrainStep(m, crit, cagrad_c=0.5, private_mult=2.0, dtype=gaitk.DTYPE_F32,
                                           p2p=mode != "nccl", use_graph=mode == "p2p_graph")
        models[mode] = m
    batches = []
    for i in range(2):
        xs, y = O.synth_weargait_batch(Bg, seed=100 + i)
        sl = slice(rank * 64, (rank + 1) * 64)
        batches.append(([torch.from_numpy(x[sl]).cuda() for x in xs], torch.from_numpy(y[sl]).cuda(), torch.from_numpy(y).cuda()))
    for it in range(9):
        xs, y, yg = batches[it % 2]
        for mode in ("nccl", "p2p", "p2p_graph"):
            loss, _ = steps[mode].step(xs, [y] * 3, ys_global=[yg] * 3)
            if mode != "nccl":
                assert not steps[mode].exchange_failed(), "peer did not arrive"
    torch.cuda.synchronize()
    flat = {k: models[k].flat_params().detach().clone() for k in models}
    err = float((flat["nccl"] - flat["p2p"]).abs().max()); scale = float(flat["nccl"].abs().max())
    assert err <= 2e-6 * scale + 1e-7, ("p2p vs nccl", err, scale)
    assert torch.equal(flat["p2p"], flat["p2p_graph"]), "graph replay differs from eager"
    gathered = [torch.empty_like(flat["p2p"]) for _ in range(world)]
    dist.all_gather(gathered, flat["p2p"])
    for r in range(world):
        assert torch.equal(gathered[r], gathered[0]), f"rank {r} diverged from rank 0"
    l_n, _ = steps["nccl"].stats(); l_p, _ = steps["p2p"].stats()
    assert torch.allclose(l_n, l_p, rtol=1e-5, atol=1e-6), (l_n, l_p)
    if rank == 0:
        print(f"P2P_CHECK_OK world={world} max|nccl-p2p|={err:.2e} losses={l_p.cpu().tolist()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
