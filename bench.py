#!/usr/bin/env python
"""bench.py -- gait windows / second per training step (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # reference arm: CPU port on the host cores

Workload (config.workload): BASELINE.json configs[1] -- WearGait full multimodal training step, sync
loader semantics (one label per window, shared head), GCL (m=0.2, s=25), CAGrad c=0.5, SGD(momentum 0.9,
wd 1e-4); B windows of (64,2)+(64,13)+(64,24) fp32 per GPU (weak scaling), synthetic data of
WearGait shape, random-init weights.  A step = gather-free fused forward + 3 losses +
CAGrad/private backward + (all-reduce) + SGD over one batch.

`value` : inputs resident in HBM, CUDA-event timing on the launching stream, max over ranks.
`e2e`   : the same step through the public API with pinned HOST buffers: H2D copies of the batch and a
          D2H read of (loss, correct) every step inside the timed region.
`roofline`: the dominant kernel (the insole stream kernel) timed live with CUDA events; algorithmic
          bytes = its input bytes read once per window.
`cpu_baseline`: the oracle port (oracle/gait_oracle.py, torch-CPU ops as the reference dispatches)
          on a bounded sample; rank 0, N=1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np
import torch

METRIC = "gait windows/sec per train step"
UNIT = "windows/s"
T, DIMS = 64, (2, 13, 24)
BYTES_PER_WINDOW = T * sum(DIMS) * 4 + 8          # fp32 inputs read once + one int64 label
COUNTS = [[400, 600]] * 3                          # p(PD) = 0.6 -> unequal class counts (GCL needs them)


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread on rank 0 (a
    20-step timed region lasts ~18 ms, too short for nvidia-smi's own loop; a faster poll on every rank competes with the
    launching threads for the host cores), nvidia-smi -lms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index = index; self.rows = []; self.proc = None; self.nv = None; self.h = None
        self.sm = []; self.bits = 0; self.mx = None; self._stop = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        return pynvml, h

    def _poll(self):
        nv, h = self.nv, self.h
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                break
            time.sleep(0.002)       # a 20-step timed region lasts ~18 ms: ~2 ms polling gives it 8-10 samples (rank 0 only)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True); self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self._stop = True; self.th.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(n for b, n in self.REASONS if self.bits & b), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------- synthetic inputs
# Generators of WearGait- / FoG-shaped data.  They live here so that the GPU arm never imports oracle/
# (only the cpu_baseline / reference leg does); same distributions as the generators the tests use.
def synth_weargait_labels(B: int, seed: int = 0, p_pd: float = 0.6):
    rng = np.random.default_rng(seed)
    y = (rng.random(B) < p_pd).astype(np.int64)
    if B > 1:
        y[0], y[1] = 0, 1
    return y


def synth_weargait_batch(B: int, T_: int = 64, seed: int = 0, p_pd: float = 0.6):
    """walkway U[0,1), insole N(0,1), IMU N(0, sigma^2) with sigma = 2 for PD and 1 for HC; p(PD) = 0.6"""
    rng = np.random.default_rng(seed)
    y = (rng.random(B) < p_pd).astype(np.int64)
    if B > 1:
        y[0], y[1] = 0, 1
    xw = rng.random((B, T_, 2), dtype=np.float32)
    xi = rng.standard_normal((B, T_, 13), dtype=np.float32)
    sig = np.where(y == 1, 2.0, 1.0).astype(np.float32)[:, None, None]
    xm = rng.standard_normal((B, T_, 24), dtype=np.float32) * sig
    return [xw, xi, xm], y


def synth_fog_batch(B: int, seed: int = 0, pose_len=101, sens_len=426, joints=7, sens_ch=6):
    """skeleton U[0,1) and sensor N(0,1) clips zero-padded at the end from random true lengths; 3 classes p=(.5,.3,.2)"""
    rng = np.random.default_rng(seed)
    y = rng.choice(3, size=B, p=[0.5, 0.3, 0.2]).astype(np.int64)
    sk = rng.random((B, pose_len, joints * 3), dtype=np.float32)
    se = rng.standard_normal((B, sens_len, sens_ch), dtype=np.float32)
    Ls = rng.integers(min(40, pose_len // 3), pose_len + 1, size=B)
    Lt = rng.integers(min(140, sens_len // 3), sens_len + 1, size=B)
    for b in range(B):
        sk[b, Ls[b]:] = 0.0; se[b, Lt[b]:] = 0.0
    return sk, se, y


def synth_fog_labels(B: int, seed: int = 0):
    return np.random.default_rng(seed).choice(3, size=B, p=[0.5, 0.3, 0.2]).astype(np.int64)


def synth(B, seed, T_=T):
    return synth_weargait_batch(B, T_=T_, seed=seed)


def stream_labels(kind: str, B: int, rank: int, i: int):
    """Label vectors (one per stream) of host batch i on `rank` -- regenerated from the seeds alone, so that every rank can
    form the label vectors of the GLOBAL batch (the weighted-mean denominators) without an exchange."""
    if kind == "fog":
        y = synth_fog_labels(B, seed=1000 * rank + i)
        return [y, np.random.default_rng(7 + i).permutation(y)]
    y = synth_weargait_labels(B, seed=1000 * rank + i)
    if kind == "weargait_async":
        r = np.random.default_rng(50 + i)
        return [y, r.permutation(y), r.permutation(y)]
    return [y, y, y]


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_port_step_time(B: int, steps: int, warmup: int, threads: int):
    """Fallback when no staged reference exists: one reference training step (forward_batch + 3 criteria +
    step_cagrad_three + SGD, weargait_train.py:163-248,305-311) restated by the oracle with the same torch-CPU / SciPy calls."""
    import gait_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    import gaitk                                       # only to obtain reference-identical initial weights
    m = gaitk.WearGaitThreeModal()
    state = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    p = O.canonical_params(state, True); bufs = {}
    xs, y = synth(B, 1)
    xt = [torch.from_numpy(x) for x in xs]; yt = [torch.from_numpy(y)] * 3
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.weargait_train_step(p, bufs, xt, yt, synchronized=True, wm="gcl", counts=COUNTS, alpha=0.5)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def reference_step_times(B: int, steps: int, warmup: int, device: str, threads: int):
    """The UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py): forward_batch + make_criteria's three GCL
    losses + step_cagrad_three + SGD (train/weargait_train.py:163-248,300-311), driven by oracle/ref_harness.py on
    `device` ("cpu": all host threads; "cuda": stock torch-CUDA fp32 on the B200).  -> (seconds per step, kind)"""
    import ref_harness as H
    if H.load_reference() is None:
        if device != "cpu":
            return None, "unavailable"
        return cpu_port_step_time(B, steps, warmup, threads), "port"
    st = H.RefWearGaitStep(device, threads=threads)
    batches = []
    for i in range(2):
        xs, y = synth(B, 1 + i)
        batches.append(([torch.from_numpy(x).to(device) for x in xs], torch.from_numpy(y).to(device)))
    return st.time_steps(batches, steps, warmup), "reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.batch if args.cpu_batch <= 0 else args.cpu_batch
    # bounded: the whole run must end within a few minutes -> probe one step, then shrink the per-step sample if needed
    probe, kind = reference_step_times(min(B, 4096), 1, 1, "cpu", threads)
    est = probe[0] * B / min(B, 4096) * (args.steps + max(args.warmup, 1))
    while est > 240 and B > 4096:
        B //= 2; est /= 2
    times, kind = reference_step_times(B, args.steps, max(args.warmup, 1), "cpu", threads)
    total = sum(times); val = B * len(times) / total
    sample = (f"{len(times)} steps of B={B} windows" + (" (the GPU arm's per-GPU batch)" if B == args.batch else f" (bounded sample of B={args.batch})")
              + f", {'unmodified reference from oracle/_ref' if kind == 'reference' else 'oracle port'}, torch {torch.__version__} CPU, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, getattr(args, "workload", "weargait")), "cpu_sample_batch": B,
                   "same_config": B == args.batch, "timing": "time.perf_counter around each step (incl. the three loss .item() reads)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(B, kind="weargait"):
    if kind == "fog":
        return (f"FoG 2-stream train step (configs[3]): B={B} sequences/GPU of skeleton (101,21) + sensor (426,6) fp32, async heads, "
                "GCL m=0.2 s=25, CAGrad c=0.1, SGD mom 0.9 wd 1e-4")
    if kind == "scaled":
        return (f"WearGait scaled sweep (configs[4]: T=256, enc_out_ch 24 (insole hidden 48), shared_out_ch 32): B={B} windows/GPU of "
                "(256,2)+(256,13)+(256,24) fp32, sync labels, GCL m=0.2 s=25, CAGrad c=0.5, SGD mom 0.9 wd 1e-4; fp32 FFMA stream kernels, "
                "every window split by time over a 2-CTA thread-block cluster")
    tag = {"weargait": "configs[1]: sync labels, shared head",
           "weargait_async": "configs[2] async: independent per-stream labels, three private heads",
           "weargait_relaxed": "configs[2] relaxed input: per-step modality mask cycling the 7 MASK_COMBOS, masked streams zero-filled "
                               "and dropped from the CAGrad task list"}[kind]
    return (f"WearGait 3-stream multimodal train step ({tag}): B={B} windows/GPU of (64,2)+(64,13)+(64,24) fp32, "
            "GCL m=0.2 s=25, CAGrad c=0.5, SGD mom 0.9 wd 1e-4")


MASK_CYCLE = [(True, False, False), (False, True, False), (False, False, True), (True, True, False), (True, False, True),
              (False, True, True), (True, True, True)]


def build_workload(args, gaitk, dev, rank):
    """-> dict(model, crit, step_kw(i), host batches, stream names, dims, units)"""
    kind = args.workload; B = args.batch
    if kind == "fog":
        model = gaitk.MultiModalMultiTaskModel(21, 6, 6, 6, 426, 16, 8, 128, 3, synchronized_loading=False).to(dev)
        counts = [[500, 300, 200], [450, 330, 220]]
        crit = [gaitk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in counts]
        host = []
        for i in range(2):
            sk, se, y = synth_fog_batch(B, seed=1000 * rank + i)
            ys = stream_labels(kind, B, rank, i)
            assert np.array_equal(ys[0], y)
            host.append(([torch.from_numpy(sk).pin_memory(), torch.from_numpy(se).pin_memory()],
                         [torch.from_numpy(v).pin_memory() for v in ys]))
        return dict(model=model, crit=crit, host=host, names=("skeleton", "sensor"), dims=((101, 21), (426, 6)),
                    cagrad_c=0.1, private_mult=1.0, dtype="f32", kw=lambda i: {})
    sync = kind != "weargait_async"
    scaled = kind == "scaled"
    Tw = 256 if scaled else T
    model = gaitk.WearGaitThreeModal(synchronized=sync, **(dict(enc_out_ch=24, shared_out_ch=32) if scaled else {})).to(dev)
    crit = [gaitk.GCLLoss(cls_num_list=c, m=0.2, s=25, noise_mul=0.0) for c in COUNTS]
    host = []
    for i in range(2):
        xs, y = synth(B, 1000 * rank + i, Tw)
        yl = stream_labels(kind, B, rank, i)
        assert np.array_equal(yl[0], y)
        if sync:
            yt = torch.from_numpy(y).pin_memory(); ys = [yt, yt, yt]
        else:
            ys = [torch.from_numpy(v).pin_memory() for v in yl]
        host.append(([torch.from_numpy(x).pin_memory() for x in xs], ys))
    kw = (lambda i: dict(enabled=MASK_CYCLE[i % 7], tasks=MASK_CYCLE[i % 7])) if kind == "weargait_relaxed" else (lambda i: {})
    return dict(model=model, crit=crit, host=host, names=("walkway", "insole", "imu"), dims=((Tw, 2), (Tw, 13), (Tw, 24)),
                cagrad_c=0.5, private_mult=2.0, dtype="f32" if scaled else args.dtype, kw=kw)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    import gaitk
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(0)                               # identical replicas on every rank
    wl = build_workload(args, gaitk, dev, rank)
    model, crit, host, names, dims = wl["model"], wl["crit"], wl["host"], wl["names"], wl["dims"]
    for c in crit:
        c.consume_rng = False
    ns = len(names)
    bytes_per_unit = sum(t * d for t, d in dims) * 4 + 8
    dtype_id = {"tf32": gaitk.DTYPE_TF32, "bf16x3": gaitk.DTYPE_BF16X3}.get(wl["dtype"], gaitk.DTYPE_F32)
    step = gaitk.FusedTrainStep(model, crit, cagrad_c=wl["cagrad_c"], max_norm=1.0, lr=1e-3, momentum=0.9, weight_decay=1e-4,
                                private_mult=wl["private_mult"], process_group=None if world > 1 else False, dtype=dtype_id,
                                use_graph=bool(args.graph), p2p=bool(args.p2p) and world > 1)
    NBUF = len(host)
    devb = [([x.to(dev) for x in xs], [y.to(dev) for y in ys]) for xs, ys in host]
    # global label vectors (all ranks' labels; cheap) fix the weighted-mean denominators
    def global_labels(i):
        if world == 1:
            return None
        per_rank = [stream_labels(args.workload, B, r, i) for r in range(world)]
        out = [torch.from_numpy(np.concatenate([pr[s_] for pr in per_rank])).to(dev) for s_ in range(ns)]
        if args.workload in ("weargait", "weargait_relaxed", "scaled"):
            out = [out[0]] * ns                           # one shared label vector (same device pointer: one histogram)
        return out
    yglob = [global_labels(i) for i in range(NBUF)]
    l2_flush = None
    if B * bytes_per_unit < 2 * 126e6:                 # small batches: flush L2 between steps instead
        l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seq = [0]                                          # running step index: batch rotation, mask cycle and (data-parallel) exchange-buffer
                                                       # parity advance together, so the same few (batch, options, parity) combinations recur
    def one_step(_i=None):
        i = seq[0]; seq[0] += 1
        xs, ys = devb[i % NBUF]
        if l2_flush is not None:
            l2_flush.fill_(i & 0xff)
        step.step(xs, ys, ys_global=yglob[i % NBUF], **wl["kw"](i))

    # ---- device-resident timing
    # CUDA-graph priming (untimed, before the W warm-up steps): every (batch, options, parity) combination is captured on its second
    # sighting (fused_step.py), so two passes over the combinations leave nothing to capture inside the warm-up or the timed region
    n_combo = (14 if args.workload == "weargait_relaxed" else 2)
    for _ in range(2 * n_combo + 2):
        one_step()
    for i in range(args.warmup):
        one_step(i)
    # the sampler starts BEFORE the barrier: NVML initialisation on rank 0 must not delay its first timed step, or the
    # other ranks wait for it inside their timed region (the exchange couples the ranks every step)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # one more untimed step BEHIND the barrier, with the start event recorded right after it on the stream: its exchange aligns
        # the GPUs on the device, so the host-side skew with which the ranks leave the barrier (up to ~1 ms = 5 % of a 20-step
        # region; the ranks that start first would wait for the last one inside their timed region) stays out of the device-timed steps
        one_step()
    ev0.record()
    for i in range(args.steps):
        one_step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if l2_flush is not None:                           # subtract nothing: the flush is part of the timed loop; report it
        pass
    clocks = sampler.stop() if rank == 0 else None
    loss, correct = step.stats(); loss = loss.cpu().tolist()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    rank_ms = [ms]
    if world > 1:
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        rank_ms = [float(x.item()) / args.steps for x in allms]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        rank_ms = [ms / args.steps]
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)
    # replicas must stay bit-identical (every rank ran the same deterministic solve + update on the same reduced buffer)
    flat = model.flat_params()
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum(), (flat.view(torch.int32).to(torch.int64)).sum().double()])
    chks = [chk.cpu().tolist()]
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        chks = [c.cpu().tolist() for c in allc]
    replicas_identical = all(c == chks[0] for c in chks)

    # ---- end-to-end: pinned host buffers -> H2D -> step -> D2H of (loss, correct), every step.  The H2D copy of
    # batch i+1 is issued on a copy stream before step i is launched (two device slots), so it overlaps compute.
    def e2e_loop(n):
        xs, ys = host[0]; step.stage_host(xs, ys, slot=0)
        for i in range(n):
            if i + 1 < n:
                xs, ys = host[(i + 1) % NBUF]; step.stage_host(xs, ys, slot=(i + 1) % 2)
            out = step.step_staged(slot=i % 2, ys_global=yglob[i % NBUF], **wl["kw"](i))
        return out
    e2e_loop(min(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t2.item()) * 1e-3)
    h2d = B * (bytes_per_unit - 8) + sum(int(y.numel()) * 8 for y in dict.fromkeys(host[0][1]))
    d2h = 2 * ns * 4

    # ---- end-to-end with the dataset resident in HBM (the B200-first data path): the frame stores are uploaded once
    # per fold; every step the host sends only the window-start indices and labels (pinned), the kernels gather.
    res_value = None
    if args.workload == "weargait":
        # every rank holds ITS shard of the fold's frame stores (sharded by window); the per-step index /
        # label vectors of all ranks are derived from the seeds, so the global label vector needs no exchange
        stores = [torch.cat([devb[i][0][s].reshape(-1, DIMS[s]) for i in range(NBUF)]) for s in range(3)]
        def res_batch(r_, i):
            gen = torch.Generator().manual_seed(1234 + r_ + 97 * i)
            perm = torch.randperm(NBUF * B, generator=gen)[:B]
            yc = np.concatenate([stream_labels("weargait", B, r_, j)[0] for j in range(NBUF)])
            return perm, torch.from_numpy(yc)[perm].contiguous()
        idx_host, y_host, yg_res = [], [], []
        for i in range(4):
            perm, yy = res_batch(rank, i)
            idx_host.append((perm * T).to(torch.int64).pin_memory()); y_host.append(yy.pin_memory())
            if world > 1:
                g = torch.cat([res_batch(r_, i)[1] for r_ in range(world)]).to(dev); yg_res.append([g, g, g])
            else:
                yg_res.append(None)
        model.set_window(T)
        for i in range(3):
            step.step_indices(stores, [idx_host[i % 4]] * 3, [y_host[i % 4]] * 3, ys_global=yg_res[i % 4])
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:                                   # align the GPUs on the device (see the device-resident loop above)
            step.step_indices(stores, [idx_host[3]] * 3, [y_host[3]] * 3, ys_global=yg_res[3])
        r0.record()
        for i in range(args.steps):
            step.step_indices(stores, [idx_host[i % 4]] * 3, [y_host[i % 4]] * 3, ys_global=yg_res[i % 4])
        r1.record(); barrier()
        t3 = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        res_value = world * B * args.steps / (float(t3.item()) * 1e-3)

    # ---- dominant kernel, timed alone with CUDA events on the launching stream (rank 0)
    roof = None; per_stream = {}
    if rank == 0:
        peak, peak_src = peaks()
        scratch = torch.empty(model.plan().NP, dtype=torch.float32, device=dev)
        step.pg = False                                   # rank-local timing: no collective in this section
        for s_ in range(ns):
            tasks = [k == s_ for k in range(ns)]
            xs, ys = devb[0]
            # the gradient half of the step for this stream alone (label histogram + zero fill + THE stream kernel + its reduce;
            # no solve / update), launched eagerly on the current stream
            for _ in range(2):
                step._step_impl(xs, ys, tasks=tasks, part="grads")
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record()
            for r in range(reps):
                xs, ys = devb[r % NBUF]
                step._step_impl(xs, ys, tasks=tasks, part="grads")
            b.record(); torch.cuda.synchronize()
            per_stream[names[s_]] = a.elapsed_time(b) / reps
        dom = max(per_stream, key=per_stream.get)
        s_ = names.index(dom)
        alg = B * (dims[s_][0] * dims[s_][1] * 4 + 8)
        ach = alg / (per_stream[dom] * 1e-3) / 1e9
        kname = {"tf32": "stream_kernel_tc", "bf16x3": "stream_kernel_ws"}.get(wl["dtype"], "stream_kernel")
        # DRAM traffic / tensor-pipe utilisation of the same kernel from the committed `ncu --set full` capture
        # (profiles/r1_ncu_full_final.json, made by scratch/ncu_summary.py); per launch, scaled to this batch
        traffic = tensor_pct = None; traffic_src = None
        sig = {"walkway": "StreamCfg<0, 2,", "insole": "StreamCfg<1, 13,", "imu": "StreamCfg<0, 24,",
               "skeleton": "StreamCfg<2, 21,", "sensor": "StreamCfg<5, 18,"}.get(dom)
        pdir = ROOT / "profiles"
        cands = {"tf32": ["r1_ncu_full_final.json"], "bf16x3": sorted(q.name for q in pdir.glob("r3_ncu_full_ws*.json"))[::-1] + sorted(q.name for q in pdir.glob("r2_ncu_full_ws_*.json"))[::-1],
                 "f32": sorted(q.name for q in pdir.glob("r3_ncu_full_fog*.json"))[::-1]}.get(wl["dtype"], [])
        for pf in cands:
            prof = pdir / pf
            if not (sig and prof.exists()) or traffic is not None:
                continue
            for nm, rec in json.load(open(prof)).items():
                if kname + "<" in nm and sig in nm and rec.get("dram_bytes_read") is not None:
                    traffic = (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * B / rec["batch"]
                    tensor_pct = rec.get("tensor_pipe_pct_of_peak"); traffic_src = "profiles/%s (ncu --set full, B=%d)" % (pf, rec["batch"])
        roof = {"bound": "hbm", "kernel": f"{kname}<{dom}> (fused fwd+loss+bwd)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "tensor_pipe_pct_of_peak": tensor_pct, "peak_source": peak_src,
                "ms_per_launch": per_stream[dom], "algorithmic_bytes_per_launch": alg,
                "note": ("fp32 FFMA2 kernel, one 101-row clip per 128-thread CTA, 3 CTAs per SM: bound by CTA barriers, shared-memory "
                         "latency and dependent-instruction latency at 12 warps per SM (DESIGN.md 3.1d), not by HBM; DRAM traffic equals "
                         "the algorithmic bytes (the next clip is bulk-prefetched while one is computed); the timed launch also contains "
                         "the label histogram, the zero fill and the stream's reduce kernel" if args.workload == "fog" else
                         "fp32 FFMA2 path (parity 1e-5): issue- and latency-bound, not HBM-bound (DESIGN.md 3.1 / 3.1d); the timed launch "
                         "also contains the label histogram, the zero fill and the stream's reduce kernel" if wl["dtype"] == "f32" else
                         "not HBM-bound (DESIGN.md 3.1c): DRAM traffic equals the algorithmic bytes (inputs are read once); at the "
                         "reference's channel widths (N = 16 outputs) the kernel is bound by the tensor pipe's fixed per-instruction "
                         "cost (~40 clocks per M = 128 tcgen05.mma whatever N <= 32 is; 70 / 134 / 76 MMAs per tile) and the row warps' "
                         "epilogues between them; the timed launch also contains the label histogram, the zero fill and the stream's "
                         "reduce kernel (~15 us together)"),
                "per_stream_ms": per_stream,
                "step_hbm_gbs": B * bytes_per_unit / (ms_max / args.steps * 1e-3) / 1e9}

        if args.workload == "scaled":
            # 22.816 MFLOP per window per step, AI 286 FLOP/B > ridge: the only tensor-bound configuration
            pk = ROOT / "MEASURED_PEAKS.json"
            tpeak = float(json.loads(pk.read_text())["bf16_tflops_sustained"]) if pk.exists() else 1413.9
            tf = 22.816e6 * B / (ms_max / args.steps * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "whole step (3 fp32 FFMA stream kernels on 2-CTA clusters + reduce + update)",
                    "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "traffic": None,
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if pk.exists() else "fallback",
                    "algorithmic_flops_per_window": 22.816e6, "per_stream_ms": per_stream,
                    "note": "this configuration runs on the fp32 FFMA path (parity 1e-5 against the wg_scaled golden); the tensor-core "
                            "kernels are instantiated for the reference's default widths only, so the fraction of the bf16 tensor peak "
                            "is the honest distance to north_star's >= 50 % tensor-pipe target, not a tensor-pipe measurement"}

    # ---- batch sweep of the fused step (device-resident, CUDA-graph replay), rank 0, N = 1: the reference trains at B = 64
    sweep = None
    if rank == 0 and world == 1 and args.workload == "weargait" and not args.no_sweep:
        sweep = {}
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        for Bs in (64, 512, 4096, 32768):
            if Bs > B:
                continue
            xs_, ys_ = devb[0]
            xs_ = [x[:Bs].contiguous() for x in xs_]; y_ = ys_[0][:Bs].contiguous(); ys_ = [y_, y_, y_]
            for _ in range(3):
                step.step(xs_, ys_)
            torch.cuda.synchronize()
            n_it = 20; tot = 0.0
            for _ in range(n_it):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step.step(xs_, ys_); b.record(); torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            sweep[str(Bs)] = {"ms_per_step": tot / n_it, "windows_per_s": Bs / (tot / n_it * 1e-3)}
        del flush

    cpu = None; cuda_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, kind = reference_step_times(args.cpu_batch, 3, 1, "cpu", threads)
        cpu = {"value": args.cpu_batch * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{len(times)} steps of B={args.cpu_batch} windows (bounded sample of B={B}), "
                         f"{'unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port'}, torch {torch.__version__} CPU"}
        # north_star's denominator: the reference's OWN code on this B200 under stock torch-CUDA (fp32), same workload
        if args.workload == "weargait":
            cuda_ref = {}
            for Bs in (64, 4096, 32768):
                if Bs > B:
                    continue
                tms, kind2 = reference_step_times(Bs, 5, 2, "cuda", threads)
                if tms is None:
                    cuda_ref = {"unavailable": "no staged reference (oracle/_ref)"}; break
                mean = sum(tms) / len(tms)
                cuda_ref[str(Bs)] = {"ms_per_step": 1e3 * mean, "windows_per_s": Bs / mean,
                                     "gaitk_speedup": (sweep[str(Bs)]["windows_per_s"] / (Bs / mean)) if sweep and str(Bs) in sweep else None}
            if "unavailable" not in cuda_ref:
                cuda_ref["how"] = ("unmodified reference step (forward_batch + GCL x3 + step_cagrad_three + SGD, oracle/_ref via "
                                   "oracle/ref_harness.py) on this GPU under stock torch-CUDA fp32, host wall clock with "
                                   "torch.cuda.synchronize on both sides, 2 warm-up + 5 timed steps, incl. the .item() loss reads")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "rank_ms_per_step": rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": workload_name(B, args.workload), "parallelism": f"dp{world}", "global_batch": world * B,
                       "l2": "inputs larger than L2 (2 rotating batches)" if l2_flush is None else "256 MiB L2 flush between steps",
                       "timing": "CUDA events on the launching stream, barrier+sync both sides, max over ranks",
                       "cuda_graph": bool(args.graph),
                       "exchange": "none (1 GPU)" if world == 1 else ("gaitk_p2p_allreduce over NVLink peer memory, inside the step's single CUDA graph"
                                                                    if (args.p2p and getattr(step, "_p2p", None) is not None) else
                                                                    "NCCL all_reduce of gbuf between two CUDA graphs" + (" (peer-memory exchange unavailable: fell back)" if args.p2p else ""))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "pinned host batch of " + "+".join("(B,%d,%d)" % (t_, c_) for t_, c_ in dims) + " fp32 + labels copied every "
                            "step (copy of batch i+1 overlaps step i), result (loss[n], correct[n]) read back every step"},
            "e2e_resident": None if res_value is None else {
                "value": res_value, "unit": UNIT, "h2d_bytes_per_step": B * 16, "d2h_bytes_per_step": d2h,
                "note": "frame stores resident in HBM, sharded by window over the ranks (uploaded once per fold); per step the "
                        "host sends int64 window-start indices + labels, the stream kernels gather the windows (win_start "
                        "path); result read back every step; whole-job windows/s, max over ranks"},
            "gpu_launches": (4 + ns) * args.steps,          # denominators, zero-fill, ns stream kernels, one reduce, update
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "cuda_reference": cuda_ref, "batch_sweep": sweep,
            "replicas_identical": replicas_identical, "param_checksum": chks[0],
            "final_losses": loss,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gaitk", choices=["gaitk", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="windows per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=-1, help="windows per step of the CPU arm (default: the GPU arm's batch for "
                    "--impl reference, shrunk only if the run would exceed a few minutes; 4096 for the gaitk arm's cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the B = 64 .. 32768 sweep of the fused step")
    ap.add_argument("--p2p", type=int, default=1, help="1 (default) = data-parallel exchange by the in-tree peer-memory all-reduce kernel "
                    "gaitk_p2p_allreduce inside the step's single CUDA graph (round 2, 8 GPUs: 1.098 vs 1.110 ms per step with NCCL; falls back "
                    "to NCCL when torch symmetric memory is unavailable, and says so in config.exchange); 0 = NCCL all_reduce between two graphs")
    ap.add_argument("--graph", type=int, default=1, help="replay the step (kernels + the NCCL all-reduce when data-parallel) as one CUDA graph")
    ap.add_argument("--workload", default="weargait", choices=["weargait", "weargait_async", "weargait_relaxed", "fog", "scaled"],
                    help="default = BASELINE.json configs[1]; the others are extra report lines")
    ap.add_argument("--dtype", default="bf16x3", choices=["f32", "tf32", "bf16x3"],
                    help="contraction arithmetic of the stream kernels: bf16x3 = split-bf16 operands (hi + lo, three tcgen05 passes, "
                         "fp32 accumulate; the warp-specialised kernel), tf32 = round-1 tcgen05 + mma.sync kernel, f32 = FFMA")
    args = ap.parse_args()
    if args.workload == "scaled" and args.batch == 32768:
        args.batch = 4096                                   # BASELINE configs[4]: batch 4096 per GPU
    args.warmup = max(args.warmup, 3) if args.impl == "gaitk" else args.warmup
    if args.impl == "gaitk" and args.cpu_batch <= 0:
        args.cpu_batch = 4096
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
