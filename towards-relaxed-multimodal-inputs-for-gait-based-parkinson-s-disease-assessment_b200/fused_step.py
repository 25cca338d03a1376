"""FusedTrainStep -- the fast path for one training step of the reference trainers.

Equivalent, on identical inputs, to ``forward_batch`` -> three criteria -> ``step_cagrad_three`` ->
``optimizer.step()`` of ``train/weargait_train.py`` :163-248,305-311 (private_mult = 2) or to
``process_batch(train=True)`` of ``train/fbg_fog_train.py`` :46-152 (private_mult = 1), but as

    gaitk_step_grads   one fused kernel per stream (forward + loss + whole backward, inputs read once)
    [all-reduce]       one NCCL all-reduce of the ~26 KB gradient buffer when data-parallel
    gaitk_step_update  single-CTA Gram + simplex solve + combine + clip + SGD

with no host synchronisation; loss / accuracy stay on the device until the caller reads them.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import LossDesc, check, lib, ptr_array, stream_handle
from .classification_losses import criterion_spec
from .plan import FlatParamModule


class _PendingResult:
    """(loss[n], correct[n]) of one step on its way to a pinned host buffer (FusedTrainStep.step_indices_async)."""
    __slots__ = ("_buf", "_ready")

    def __init__(self, buf, ready):
        self._buf = buf; self._ready = ready

    def result(self) -> torch.Tensor:
        self._ready.synchronize()
        return self._buf.clone()


class FusedTrainStep:
    def __init__(self, model: FlatParamModule, criterions: Sequence, *, cagrad_c: float, max_norm: float = 1.0,
                 lr: float = 1e-3, momentum: float = 0.9, weight_decay: float = 1e-4, private_mult: float = 2.0,
                 process_group=None, consistency_lambda: float = 0.0, solver: int = _lib.SOLVER_SLSQP,
                 dtype: int = None, use_graph: bool = False, p2p: bool = False):
        self.model = model; self.criterions = list(criterions)
        self.cagrad_c = float(cagrad_c); self.max_norm = float(max_norm)
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.private_mult = float(private_mult); self.pg = process_group
        self.consistency_lambda = float(consistency_lambda); self.solver = int(solver)
        self.dtype = int(getattr(model, 'compute_dtype', _lib.DTYPE_F32) if dtype is None else dtype)
        self._mom = None; self._gbuf = None; self._denom = None; self._diag = None
        self._pinned = {}; self._dev_in = {}
        self._copy_stream = None; self._staged = {}; self._slot_free = {}; self._host_out = {}
        self.use_graph = bool(use_graph); self._graphs = {}
        # p2p: data-parallel exchange by gaitk_p2p_allreduce over symmetric (peer-mapped) memory instead of NCCL
        self.p2p = bool(p2p); self._p2p = None; self._red = None

    # ------------------------------------------------------------------ buffers
    def _buffers(self, plan):
        dev = plan.device
        if self._gbuf is None or self._gbuf.numel() != plan.gbuf_floats or self._gbuf.device != dev:
            self._gbuf = torch.zeros(plan.gbuf_floats, dtype=torch.float32, device=dev)
            self._denom = torch.ones(_lib.DENOM_FLOATS, dtype=torch.float32, device=dev)
            self._diag = torch.zeros(_lib.DIAG_FLOATS, dtype=torch.float32, device=dev)
        if self._mom is None or self._mom.numel() != plan.NP or self._mom.device != dev:
            self._mom = torch.zeros(plan.NP, dtype=torch.float32, device=dev)
        if self._red is None or self._p2p is None:
            self._red = self._gbuf
        return self._gbuf, self._denom, self._diag, self._mom

    STAGE_LIMIT = 16 << 20            # bytes per tensor: larger inputs are not copied into staging buffers

    def _static(self, name, seq):
        """inputs -> address-stable tensors: tensors of at most STAGE_LIMIT bytes are copied into per-(position, shape) staging
        buffers (aliases among the inputs, e.g. ys = [y, y, y], stay aliases: the label histogram is shared by address)"""
        if seq is None:
            return None
        store = self.__dict__.setdefault("_static_in", {})
        out, first = [], {}
        for i, t in enumerate(seq):
            if t is None or not t.is_cuda or t.numel() * t.element_size() > self.STAGE_LIMIT:
                out.append(t); continue
            j = first.setdefault(id(t), i)
            if j != i:
                out.append(out[j]); continue
            key = (name, i, tuple(t.shape), t.dtype, t.device)
            buf = store.get(key)
            if buf is None:
                buf = store[key] = torch.empty_like(t, memory_format=torch.contiguous_format)
            if buf.data_ptr() != t.data_ptr():
                buf.copy_(t)
            out.append(buf)
        return out

    def _noise_buffers(self, B, K, dev, off_fns):
        """persistent (B, K) logit-offset buffers (stable addresses: CUDA-graph replays read them); refreshed with a new draw
        on every eager call, and by step() before every replay -- never while a capture is in progress"""
        key = (B, K, str(dev))
        if getattr(self, "_noise_key", None) != key:
            self._noise = [None if f is None else torch.zeros(B, K, dtype=torch.float32, device=dev) for f in off_fns]
            self._noise_key = key
        if not torch.cuda.is_current_stream_capturing():
            for b, f in zip(self._noise, off_fns):
                if b is not None:
                    b.copy_(f(b))
        return self._noise

    def _p2p_state(self, plan):
        """Symmetric allocation [gbuf parity 0 | gbuf parity 1 | flag word] mapped by every rank of the group (torch
        symmetric memory does the handle exchange), device arrays of the peers' pointers, the local exchange counter and
        the local reduced buffer.  Collective: every rank must reach it at the same step."""
        if self._p2p is not None:
            return self._p2p
        try:
            return self._p2p_setup(plan)
        except Exception as e:                       # no symmetric memory on this system: every rank falls back to NCCL
            import warnings
            warnings.warn(f"gaitk: peer-memory exchange unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce")
            self.p2p = False
            return None

    def _p2p_setup(self, plan):
        import torch.distributed._symmetric_memory as symm
        dist = torch.distributed
        group = dist.group.WORLD if self.pg in (None, False) else self.pg
        dev = plan.device
        n = plan.gbuf_floats; n_pad = (n + 63) // 64 * 64
        buf = symm.empty(2 * n_pad + 64, dtype=torch.float32, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        ptrs = [int(q) for q in hdl.buffer_ptrs]
        st = dict(buf=buf, hdl=hdl, rank=int(hdl.rank), world=int(hdl.world_size), steps=0,
                  gbufs=[buf[0:n], buf[n_pad:n_pad + n]],
                  peer_gbuf=[torch.tensor([q + par * n_pad * 4 for q in ptrs], dtype=torch.int64, device=dev) for par in (0, 1)],
                  peer_flag=torch.tensor([q + 2 * n_pad * 4 for q in ptrs], dtype=torch.int64, device=dev),
                  counter=torch.zeros(1, dtype=torch.int32, device=dev),
                  gsum=torch.zeros(n, dtype=torch.float32, device=dev))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                     # every rank's buffer is zeroed before anyone publishes a flag
        self._p2p = st; self._red = st["gsum"]
        return st

    def _distributed(self) -> bool:
        """process_group=False: never; a group: always; None: whenever a default group with >1 ranks exists."""
        if self.pg is False:
            return False
        if self.pg is not None:
            return True
        d = torch.distributed
        return d.is_available() and d.is_initialized() and d.get_world_size() > 1

    @property
    def momentum_buffer(self):
        return self._mom

    def stats(self):
        """(losses[n_streams], correct[n_streams]) device tensors of the last step."""
        plan = self.model.plan()
        st = (self._red if self._red is not None else self._gbuf)[3 * plan.P + plan.NP:]
        return st[0:plan.n_streams], st[4:4 + plan.n_streams]

    def diag(self):
        return self._diag

    def exchange_failed(self) -> bool:
        """True once a peer failed to arrive in the peer-memory exchange (sticky; the updates since then were skipped).
        Reads one float from the device."""
        return self._diag is not None and float(self._diag[_lib.DIAG_EXCHANGE]) < 0.0

    def reset_exchange_status(self):
        if self._diag is not None:
            self._diag[_lib.DIAG_EXCHANGE] = 0.0

    # ------------------------------------------------------------------ step
    def _step_impl(self, xs: Sequence[torch.Tensor], ys: Sequence[torch.Tensor], *, enabled: Sequence[bool] = None,
             tasks: Sequence[bool] = None, ys_global: Optional[Sequence[torch.Tensor]] = None,
             win_start: Optional[Sequence[torch.Tensor]] = None, logits_out: Optional[Sequence] = None,
             update: bool = True, grads_out: Optional[torch.Tensor] = None, part: Optional[str] = None,
             par: Optional[int] = None, count: bool = True):
        """xs[s]: (B, T_s, D_s) CUDA fp32 (or frame stores with win_start[s] int64[B]); ys[s]: int64[B].
        ys_global: label vectors of the WHOLE data-parallel batch (defaults to ys) -- they fix the
        weighted-mean denominators so that shards add up to the single-GPU step."""
        model = self.model
        n = len(xs)
        if hasattr(model, "set_window") and (win_start is None or win_start[0] is None):
            model.set_window(xs[0].shape[1])
        plan = model.plan(); flat = model.flat_params()
        gbuf, denom, diag, mom = self._buffers(plan)
        p2p = self.p2p and self._distributed()
        if p2p:
            ps = self._p2p_state(plan)
            p2p = ps is not None
        if p2p:
            par = (ps["steps"] & 1) if par is None else par
            gbuf = ps["gbufs"][par]                     # this step's half of the symmetric buffer
        self._red = ps["gsum"] if p2p else gbuf
        K = plan.K
        descs = (LossDesc * _lib.MAX_STREAMS)()
        off_fns = []
        for s in range(n):
            d, off_fn = criterion_spec(self.criterions[s], K)
            descs[s] = d
            # GCL logit noise (classification_losses.py:99-105, noise_mul != 0): drawn on the host exactly as the criterion's own
            # call does (same RNG stream), uploaded into a persistent (B, K) buffer the kernels subtract from the logits
            off_fns.append(off_fn if getattr(self.criterions[s], "noise_mul", 0) not in (0, 0.0) else None)
        B = plan._check_inputs(xs, win_start)
        off_ptrs = None
        if any(f is not None for f in off_fns):
            bufs = self._noise_buffers(B, K, flat.device, off_fns)
            off_ptrs = ptr_array([0 if b is None else b.data_ptr() for b in bufs])
        enabled_mask = 0b111 if enabled is None else sum(1 << s for s, e in enumerate(enabled) if e)
        task_mask = ((1 << n) - 1) if tasks is None else sum(1 << s for s, e in enumerate(tasks) if e)
        if ys_global is None and self._distributed():
            # each rank normalising by its LOCAL sum of w[y] would make the all-reduced sum `world` times the global-mean
            # gradient (and break class-weighted means outright): the global label vectors are part of the contract
            raise _lib.GaitkError("data-parallel step: pass ys_global (the label vectors of the WHOLE global batch, one per "
                                  "stream) so that the weighted-mean denominators are global; see FusedTrainStep.step")
        yg = ys if ys_global is None else ys_global
        st = stream_handle()
        counts = (C.c_int * n)(*[int(y.numel()) for y in yg])
        check(lib().gaitk_loss_denominators(ptr_array([y.data_ptr() for y in yg]), counts, n, descs, denom.data_ptr(), st),
              "gaitk_loss_denominators")
        ws = plan.workspace(B)
        check(lib().gaitk_step_grads(plan.handle, flat.data_ptr(), ptr_array([x.data_ptr() for x in xs]),
                                     None if win_start is None else ptr_array([0 if w is None else w.data_ptr() for w in win_start]),
                                     ptr_array([y.data_ptr() for y in ys]), B, descs, off_ptrs, denom.data_ptr(), enabled_mask,
                                     task_mask, self.private_mult, self.consistency_lambda,
                                     None if logits_out is None else ptr_array([0 if l is None else l.data_ptr() for l in logits_out]),
                                     gbuf.data_ptr(), ws.data_ptr(), ws.numel(), self.dtype, st), "gaitk_step_grads")
        if part == "grads":
            return None
        if p2p:
            check(lib().gaitk_p2p_allreduce(plan.handle, ps["peer_gbuf"][par].data_ptr(), ps["peer_flag"].data_ptr(),
                                            ps["counter"].data_ptr(), ps["rank"], ps["world"], ps["gsum"].data_ptr(),
                                            diag.data_ptr(), st), "gaitk_p2p_allreduce")
            if count:
                ps["steps"] += 1
            gbuf = ps["gsum"]
        elif self._distributed():
            torch.distributed.all_reduce(gbuf, group=self.pg if self.pg not in (None, False) else None)
        self._update_part(plan, flat, gbuf, mom, diag, task_mask, update, grads_out, st)
        return self.stats()

    def _update_part(self, plan, flat, gbuf, mom, diag, task_mask, update, grads_out, st):
        check(lib().gaitk_step_update(plan.handle, flat.data_ptr() if update else None, mom.data_ptr() if update else None,
                                      gbuf.data_ptr(), task_mask, self.cagrad_c, self.max_norm, self.lr, self.momentum,
                                      self.weight_decay, None if grads_out is None else grads_out.data_ptr(),
                                      diag.data_ptr(), self.solver | _lib.SOLVER_FLAG_CHECK_EXCHANGE, st), "gaitk_step_update")
        return self.stats()

    def step(self, xs, ys, **kw):
        """One fused training step.  With use_graph=True the launch sequence (denominators, stream kernels, reduces,
        update) is captured once per distinct set of buffer addresses / options and replayed as ONE CUDA graph; when
        data-parallel, as TWO graphs (gradient half, update half) around the eager NCCL all-reduce of gbuf."""
        if not self.use_graph or kw.get("grads_out") is not None or kw.get("logits_out") is not None:
            return self._step_impl(xs, ys, **kw)
        ws = kw.get("win_start")
        if hasattr(self.model, "set_window") and (ws is None or ws[0] is None):
            self.model.set_window(xs[0].shape[1])
        # A captured graph bakes buffer ADDRESSES in.  (1) Tensors the caller reuses (resident stores, rotating device batches) are
        # recognised by address: the second time a set of addresses shows up it gets its own graph, replayed without any copy.
        # (2) Trainers hand over fresh tensors every step: on the first sighting of a set of addresses the small inputs (batches up
        # to a few thousand windows, index / label vectors) are copied into static per-shape staging buffers, so the graph is keyed
        # by shapes and options only and is captured once.  (3) Large fresh tensors cannot be staged cheaply: after 8 captures in a
        # row that were never replayed the step stays eager (see below).
        def ptrs(seq):
            return None if seq is None else tuple(0 if t is None else t.data_ptr() for t in seq)
        dist_mode = self._distributed()
        p2p = self.p2p and dist_mode and self._p2p_state(self.model.plan()) is not None
        par = (self._p2p["steps"] & 1) if p2p else 0
        # everything the captured launches bake in: buffer addresses, batch size, and -- BY VALUE -- the loss descriptors
        # (class weights after a DRW update, margins, scale, NaN flag) and every scalar option of the step
        K = self.model.plan().K
        desc_bytes = b"".join(bytes(criterion_spec(c, K)[0]) for c in self.criterions)
        def make_key(xs_, ys_, kw_):
            return (ptrs(xs_), ptrs(ys_), ptrs(kw_.get("ys_global")), ptrs(kw_.get("win_start")), tuple(kw_.get("enabled") or ()),
                    tuple(kw_.get("tasks") or ()), kw_.get("update", True),
                    xs_[0].shape[0] if kw_.get("win_start") is None else kw_["win_start"][0].numel(),
                    desc_bytes, self.model.flat_params().data_ptr(), self.lr, self.momentum, self.weight_decay, self.cagrad_c,
                    self.max_norm, self.private_mult, self.solver, self.consistency_lambda, self.dtype, dist_mode, p2p, par)
        key = make_key(xs, ys, kw)
        if key not in self._graphs:
            seen = self.__dict__.setdefault("_raw_seen", {})
            if len(seen) > 512:
                seen.clear()
            first_sighting = key not in seen
            seen[key] = True
            if first_sighting:
                xs = self._static("x", xs); ys = self._static("y", ys)
                kw = dict(kw)
                for name in ("ys_global", "win_start"):
                    if kw.get(name) is not None:
                        kw[name] = self._static(name, kw[name])
                key = make_key(xs, ys, kw)
        g = self._graphs.get(key)
        if g is None:
            # fresh (unstaged, large) tensors on every call would mean one capture per step: after 8 misses in a row stay eager
            self._miss_run = getattr(self, "_miss_run", 0) + 1
            if self._miss_run > 8:
                return self._step_impl(xs, ys, **kw)
            self._step_impl(xs, ys, **kw)                      # this call's step, eagerly (also allocates buffers / workspace)
            torch.cuda.synchronize()
            # capture records the launch sequence without executing it; later calls replay it
            if len(self._graphs) > 64:
                self._graphs.clear()
            if not dist_mode or p2p:
                # (p2p: the eager call above advanced the parity; the graph is captured for the parity of the key and
                # will be replayed whenever that parity comes round again)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    self._step_impl(xs, ys, par=par, count=False, **kw)
                self._graphs[key] = (g1, None)
            else:
                plan = self.model.plan(); flat = self.model.flat_params()
                gbuf, denom, diag, mom = self._buffers(plan)
                n = len(xs); tasks = kw.get("tasks")
                task_mask = ((1 << n) - 1) if tasks is None else sum(1 << i for i, e in enumerate(tasks) if e)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    self._step_impl(xs, ys, part="grads", **kw)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    self._update_part(plan, flat, gbuf, mom, diag, task_mask, kw.get("update", True), None, stream_handle())
                self._graphs[key] = (g1, g2)
            return self.stats()
        g1, g2 = g
        self._miss_run = 0
        if getattr(self, "_noise", None) is not None:        # a fresh GCL noise draw per step (same buffers the graph reads)
            for b, c in zip(self._noise, self.criterions):
                if b is not None:
                    b.copy_(c.logit_offset(b))
        g1.replay()
        if p2p:
            self._p2p["steps"] += 1
        if g2 is not None:
            torch.distributed.all_reduce(self._gbuf, group=self.pg if self.pg not in (None, False) else None)
            g2.replay()
        return self.stats()

    # ------------------------------------------------------------------ end-to-end (host batch) entries
    def step_host(self, xs_host: Sequence[torch.Tensor], ys_host: Sequence[torch.Tensor], **kw):
        """The call a trainer makes with a DataLoader batch: pinned host tensors in, H2D copies on the
        current stream, fused step, and a device->host read of (loss, correct)."""
        self.stage_host(xs_host, ys_host, slot=0)
        return self.step_staged(slot=0, **kw)

    def stage_host(self, xs_host: Sequence[torch.Tensor], ys_host: Sequence[torch.Tensor], slot: int = 0):
        """Start the H2D copy of a (pinned) host batch into device slot `slot` on the copy stream and return
        immediately.  With two slots the copy of batch i+1 runs under the compute of batch i:

            step.stage_host(b0, y0, 0)
            for i in range(n):
                if i + 1 < n: step.stage_host(b[i+1], y[i+1], (i + 1) % 2)
                out = step.step_staged(i % 2)
        """
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        ev_free = self._slot_free.get(slot)
        with torch.cuda.stream(self._copy_stream):
            if ev_free is not None:
                self._copy_stream.wait_event(ev_free)          # the step that last read this slot has finished
            xs = [self._to_dev((slot, "x", i), x, dev) for i, x in enumerate(xs_host)]
            same = all(y is ys_host[0] for y in ys_host)
            if same:
                y0 = self._to_dev((slot, "y", 0), ys_host[0], dev); ys = [y0] * len(ys_host)
            else:
                ys = [self._to_dev((slot, "y", i), y, dev) for i, y in enumerate(ys_host)]
            ev = torch.cuda.Event(); ev.record(self._copy_stream)
        self._staged[slot] = (xs, ys, ev)

    def step_staged(self, slot: int = 0, **kw):
        xs, ys, ev = self._staged[slot]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        loss, correct = self.step(xs, ys, **kw)
        done = torch.cuda.Event(); done.record(cur); self._slot_free[slot] = done
        return torch.cat([loss, correct]).to("cpu", non_blocking=False)

    # ------------------------------------------------------------------ device-resident dataset entry
    def step_indices(self, stores: Sequence[torch.Tensor], win_start_host: Sequence[torch.Tensor],
                     ys_host: Sequence[torch.Tensor], **kw):
        """B200-first data path: the normalised frame stores (N_frames, D) live in HBM for the whole fold
        (uploaded once, like prepare_split runs once per fold); per step the host sends only the window
        start indices (int64[B], pinned) and the labels; the stream kernels gather the windows themselves."""
        dev = stores[0].device
        same_w = all(w is win_start_host[0] for w in win_start_host)
        if same_w:
            w0 = self._to_dev(("w", 0), win_start_host[0], dev); ws = [w0] * len(win_start_host)
        else:
            ws = [self._to_dev(("w", i), w, dev) for i, w in enumerate(win_start_host)]
        same = all(y is ys_host[0] for y in ys_host)
        if same:
            y0 = self._to_dev(("yi", 0), ys_host[0], dev); ys = [y0] * len(ys_host)
        else:
            ys = [self._to_dev(("yi", i), y, dev) for i, y in enumerate(ys_host)]
        loss, correct = self.step(list(stores), ys, win_start=ws, **kw)
        return torch.cat([loss, correct]).to("cpu", non_blocking=False)

    def step_indices_async(self, stores: Sequence[torch.Tensor], win_start_host: Sequence[torch.Tensor],
                           ys_host: Sequence[torch.Tensor], slot: int = 0, **kw):
        """Pipelined step_indices: the (pinned) index / label vectors go to device slot `slot` on the copy stream, the step
        is launched behind them, its (loss, correct) is copied to a pinned host buffer without blocking, and a handle is
        returned; ``handle.result()`` waits for THIS step's numbers.  Reading step i - 1 after launching step i keeps the
        host one step ahead of the device (no launch gap, the H2D copy runs under the previous step):

            pending = None
            for i in range(n):
                h = step.step_indices_async(stores, w[i], y[i], slot=i % 2)
                if pending is not None: log(pending.result())
                pending = h
            log(pending.result())

        Two slots suffice as long as the result of the step that last used a slot has been read before it is reused."""
        dev = stores[0].device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        ev_free = self._slot_free.get(("idx", slot))
        with torch.cuda.stream(self._copy_stream):
            if ev_free is not None:
                self._copy_stream.wait_event(ev_free)          # the step that last read this slot has finished
            if all(w is win_start_host[0] for w in win_start_host):
                w0 = self._to_dev((slot, "w", 0), win_start_host[0], dev); ws = [w0] * len(win_start_host)
            else:
                ws = [self._to_dev((slot, "w", i), w, dev) for i, w in enumerate(win_start_host)]
            if all(y is ys_host[0] for y in ys_host):
                y0 = self._to_dev((slot, "yi", 0), ys_host[0], dev); ys = [y0] * len(ys_host)
            else:
                ys = [self._to_dev((slot, "yi", i), y, dev) for i, y in enumerate(ys_host)]
            ev = torch.cuda.Event(); ev.record(self._copy_stream)
        cur.wait_event(ev)
        loss, correct = self.step(list(stores), ys, win_start=ws, **kw)
        done = torch.cuda.Event(); done.record(cur); self._slot_free[("idx", slot)] = done
        res = torch.cat([loss, correct])
        out = self._host_out.get(slot)
        if out is None or out.shape != res.shape:
            out = torch.empty(res.shape, dtype=res.dtype).pin_memory(); self._host_out[slot] = out
        out.copy_(res, non_blocking=True)
        ready = torch.cuda.Event(); ready.record(cur)
        return _PendingResult(out, ready)

    def _to_dev(self, key, t, dev):
        buf = self._dev_in.get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, device=dev); self._dev_in[key] = buf
        buf.copy_(t, non_blocking=True)
        return buf
