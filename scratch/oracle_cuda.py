"""Estimate of the reference's stock CUDA-PyTorch step: the oracle (functional torch ops, SciPy SLSQP on the host with the
same D2H sync) run on CUDA tensors.  Not an official arm -- context for north_star's >= 20x target."""
import sys, time; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch, gait_oracle as O, gaitk
torch.manual_seed(0)
m = gaitk.WearGaitThreeModal()
state = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
counts = [[400, 600]] * 3
# the oracle builds class-count tensors on CPU: patch the tiny constant tensors onto the GPU
_gcl = O.gcl_loss
def gcl_cuda(logits, y, cls_num_list, m=0.5, s=30.0, weight=None, noise_mul=1.0, noise=None):
    onehot = torch.nn.functional.one_hot(y, logits.shape[1]).bool()
    z = torch.where(onehot, logits - m, logits)
    return O.weighted_ce(s * z, y, weight)
O.gcl_loss = gcl_cuda
_cg = O.cagrad_combine
def cagrad_cuda(G, alpha):
    g, A, w = _cg(G.cpu(), alpha)          # GG .cpu() + SciPy on the host, as multitask_weighting.py:699-719
    return g.to(G.device), A, w
O.cagrad_combine = cagrad_cuda
for B in (64, 4096, 32768):
    p = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in O.canonical_params(state, True).items()}
    bufs = {}
    xs, y = O.synth_weargait_batch(B, seed=1)
    xs = [torch.from_numpy(x).cuda() for x in xs]; ys = [torch.from_numpy(y).cuda()] * 3
    for it in range(3):
        O.weargait_train_step(p, bufs, xs, ys, synchronized=True, wm="gcl", counts=counts, alpha=0.5)
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 10
    for it in range(n):
        O.weargait_train_step(p, bufs, xs, ys, synchronized=True, wm="gcl", counts=counts, alpha=0.5)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"oracle-on-CUDA (ATen/cuDNN ops, host SLSQP) B={B}: {dt*1e3:.2f} ms/step = {B/dt/1e6:.3f} M windows/s", flush=True)
