"""Drop-in for ``CAGrad`` of ``train/learning/optimizers/multitask_weighting.py`` :653-776.

``backward(losses=[...], shared_parameters=[...])`` keeps the reference's observable contract: one
``losses[i].backward(retain_graph=True)`` per task (so every non-shared leaf accumulates
sum_i dL_i/dtheta, :680-688), the shared ``.grad`` overwritten with n * cagrad(G) and clipped to
``max_norm`` (:748-759, :775), and the return value ``(None, {"GTG": ndarray, "weights": ndarray})``.
The Gram matrix, the simplex solve (SciPy SLSQP on the host in the reference), the combination and the
clip run in ONE single-CTA CUDA kernel (gaitk_cagrad) -- no device->host round trip inside the step; the
returned dict converts to numpy lazily, only if somebody reads it."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, stream_handle


class _LazyExtras(dict):
    """{"GTG", "weights"} backed by the device diag vector; numpy conversion (a sync) happens on access."""
    def __init__(self, diag: torch.Tensor, n: int):
        super().__init__(); self._diag = diag; self._n = n
        dict.__setitem__(self, "GTG", None); dict.__setitem__(self, "weights", None)

    def __getitem__(self, k):
        d = self._diag.detach().cpu().numpy().astype(np.float64)
        if k == "weights":
            return d[:self._n].copy()
        if k == "GTG":
            return d[3:12].reshape(3, 3)[:self._n, :self._n].astype(np.float32)
        raise KeyError(k)

    def get(self, k, default=None):
        try:
            return self[k]
        except KeyError:
            return default


class CAGrad:
    def __init__(self, n_tasks, device: torch.device = None, c=0.4, max_norm=1.0, solver: int = _lib.SOLVER_SLSQP):
        self.n_tasks = n_tasks; self.device = device; self.c = c; self.max_norm = max_norm; self.solver = solver
        self._G = None; self._g = None; self._diag = None

    def parameters(self) -> List[torch.Tensor]:
        return []

    def _buffers(self, P: int, dev):
        if self._G is None or self._G.shape[1] != P or self._G.device != dev:
            self._G = torch.zeros(3, P, dtype=torch.float32, device=dev)
            self._g = torch.empty(P, dtype=torch.float32, device=dev)
        self._diag = torch.zeros(16, dtype=torch.float32, device=dev)
        return self._G, self._g, self._diag

    def cagrad_device(self, G: torch.Tensor, alpha: float = None, max_norm: float = None):
        """G (n_tasks, P) fp32 on CUDA -> (n * cagrad(G) clipped to max_norm, diag[16])."""
        n, P = G.shape
        if n > 3:
            raise _lib.GaitkError("the on-device CAGrad solve supports n_tasks <= 3 (all the reference trainers use)")
        g = torch.empty(P, dtype=torch.float32, device=G.device); diag = torch.zeros(16, dtype=torch.float32, device=G.device)
        Gc = G.contiguous().float()
        check(lib().gaitk_cagrad(Gc.data_ptr(), P, n, float(self.c if alpha is None else alpha),
                                 float(self.max_norm if max_norm is None else max_norm), g.data_ptr(), diag.data_ptr(),
                                 int(self.solver), stream_handle()), "gaitk_cagrad")
        return g, diag

    def get_weighted_loss(self, losses, shared_parameters, **kwargs):
        shared = list(shared_parameters)
        dims = [p.numel() for p in shared]
        P = sum(dims)
        G, g, diag = self._buffers(P, shared[0].device)
        if self.n_tasks > 3:
            raise _lib.GaitkError("n_tasks > 3 not supported")
        G.zero_()
        for i in range(self.n_tasks):
            losses[i].backward(retain_graph=True)
            off = 0
            for p, n in zip(shared, dims):
                if p.grad is not None:
                    G[i, off:off + n].copy_(p.grad.reshape(-1))
                off += n
                p.grad = None
        check(lib().gaitk_cagrad(G.data_ptr(), P, self.n_tasks, float(self.c), float(self.max_norm), g.data_ptr(),
                                 diag.data_ptr(), int(self.solver), stream_handle()), "gaitk_cagrad")
        off = 0
        for p, n in zip(shared, dims):
            p.grad = g[off:off + n].view_as(p).clone()
            off += n
        return diag

    def backward(self, losses, parameters=None, shared_parameters=None, task_specific_parameters=None, **kwargs):
        diag = self.get_weighted_loss(losses, shared_parameters)
        return None, _LazyExtras(diag, self.n_tasks)

    __call__ = backward
