"""tcgen05.mma cost by shape / layout (design evidence): cycles per MMA for the shapes the stream kernel could use, and the
TMEM lane layout of an M = 64 accumulator."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gaitk
from test_gpu_umma import idesc_bf16, run_bf16, bf16_bits, bf16_val
L = gaitk.lib()
PL = 132 * 16

def bench(name, ops, reps=256, ncols=256):
    od = torch.from_numpy(np.asarray(ops, dtype=np.uint32).view(np.int32).ravel().copy()).cuda()
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        gaitk._lib.check(L.gaitk_umma_bench(od.data_ptr(), len(ops), reps, ncols, 160 * 1024, cyc.data_ptr(), gaitk._lib.stream_handle()))
        torch.cuda.synchronize()
    c = cyc.cpu().numpy()
    print(f"{name:58s} {c[0] / (reps * len(ops)):7.1f} cyc/MMA total, {c[1] / (reps * len(ops)):6.1f} issue", flush=True)

def kmajor(M, N, n=4):       # conv-like: A K-major (rows at 16 B), B weights K-major
    return [[(2 + i % 3) * 16, PL, 128, 64 * 1024 + (i % 3) * N * 32, N * 16, 128, 1, idesc_bf16(M, N, False, False)] for i in range(n)]
def mnmajor(M, N, n=4, dcols=1):   # wgrad-like: both MN-major, K = 16 rows
    return [[(2 + 16 * (i % 8)) * 16, 128, PL, 40 * 1024 + (2 + 16 * (i % 8)) * 16, 128, PL, 1 | (((i % dcols) * (N // 8)) << 8), idesc_bf16(M, N, True, True)] for i in range(n)]
for M in (128, 64):
    for N in (16, 32, 64, 128, 256):
        bench(f"K-major  M={M} N={N} K=16 (same D)", kmajor(M, N))
for M in (128, 64):
    for N in (16, 32, 64):
        bench(f"MN-major M={M} N={N} K=16 (same D)", mnmajor(M, N))
        bench(f"MN-major M={M} N={N} K=16 (3 D regions round robin)", mnmajor(M, N, 4, 3))

# M = 64 accumulator layout: which TMEM lanes hold row i?
rng = np.random.default_rng(0)
RB = 136
IN = rng.standard_normal((8, RB, 8)).astype(np.float32); DO = rng.standard_normal((2, RB, 8)).astype(np.float32)
ops = [[(4 + 16 * ks) * 16, 128, RB * 16, (4 + 16 * ks) * 16, 128, RB * 16, int(ks > 0), idesc_bf16(64, 16, True, True)] for ks in range(8)]
D = run_bf16(bf16_bits(IN), bf16_bits(DO), ops, 32)
rows = 4 + np.arange(128)
A = bf16_val(bf16_bits(IN)).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, 64)
Bm = bf16_val(bf16_bits(DO)).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, 16)
ref = A.T @ Bm                                   # (64, 16)
lane_of = []
for i in range(64):
    d = np.abs(D[:, :16] - ref[i][None, :]).max(1) / np.abs(ref[i]).max()
    lane_of.append(int(np.argmin(d)) if d.min() < 1e-4 else -1)
print("M=64 accumulator: row -> TMEM lane", lane_of)

# ---- concurrent issuers: is the ~45-clock cost per issuing thread or per tensor pipe?
def bench_multi(name, ops, n, reps=256, ncols=512):
    od = torch.from_numpy(np.asarray(ops, dtype=np.uint32).view(np.int32).ravel().copy()).cuda()
    cyc = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        gaitk._lib.check(L.gaitk_umma_bench_multi(od.data_ptr(), len(ops), reps, ncols, 160 * 1024, n, cyc.data_ptr(), gaitk._lib.stream_handle()))
        torch.cuda.synchronize()
    c = cyc.cpu().numpy().reshape(4, 2)[:n]
    tot = c[:, 0].max()
    print(f"{name:44s} issuers {n}: {tot / (reps * len(ops)):7.1f} cyc per MMA per issuer -> {tot / (reps * len(ops) * n):6.1f} cyc per MMA overall", flush=True)
for n in (1, 2, 3, 4):
    bench_multi("K-major  M=128 N=32", kmajor(128, 32), n)
for n in (1, 2, 4):
    bench_multi("MN-major M=64 N=32", mnmajor(64, 32), n)
for n in (1, 2, 4):
    bench_multi("K-major  M=128 N=64", kmajor(128, 64), n)
