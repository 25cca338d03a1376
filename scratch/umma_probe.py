import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_umma import tf32, idesc, run
rng = np.random.default_rng(1)
halo, RB, N = 4, 137, 16
X = rng.standard_normal((32, RB, 4)).astype(np.float32)
dY = rng.standard_normal((N // 4, RB, 4)).astype(np.float32)
rows = halo + np.arange(128)
Xm = tf32(X).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, 128)   # [r][j]
Ym = tf32(dY).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, N)    # [r][n]
ref = Xm.T @ Ym
def report(tag, D):
    D = D[:, :N]
    print(f"{tag:55s} err={np.abs(D-ref).max()/np.abs(ref).max():.3e}  |D|max={np.abs(D).max():.3e} nonzero_rows={int((np.abs(D).sum(1)>0).sum())}", flush=True)
# K-major copies: A_k[j][r] as [kchunk=r/4][j][4], B_k[n][r] as [r/4][n][4]
Ak = np.zeros((32, 128, 4), np.float32); Bk = np.zeros((32, N, 4), np.float32)
for rc in range(32):
    for e in range(4):
        Ak[rc, :, e] = X[:, halo + rc*4 + e, :].reshape(128)   # j = chunk*4+e' -> X[chunk][row][e']
        Bk[rc, :, e] = dY[:, halo + rc*4 + e, :].reshape(N)
def ops_k(a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo):
    ops = []
    for kg in range(16):
        a_off = (halo + 8*kg)*16 if a_mn else (2*kg*128)*16
        b_off = (halo + 8*kg)*16 if b_mn else (2*kg*N)*16
        ops.append([a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, int(kg > 0), idesc(128, N, a_mn, b_mn)])
    return ops
report("both K-major (control)", run(Ak, Bk, ops_k(False, False, 128*16, 128, N*16, 128), 32))
report("A MN (lbo=128,sbo=chunk), B K", run(X, Bk, ops_k(True, False, 128, RB*16, N*16, 128), 32))
report("A MN (lbo=chunk,sbo=128), B K", run(X, Bk, ops_k(True, False, RB*16, 128, N*16, 128), 32))
report("A K, B MN (lbo=128,sbo=chunk)", run(Ak, dY, ops_k(False, True, 128*16, 128, 128, RB*16), 32))
report("A K, B MN (lbo=chunk,sbo=128)", run(Ak, dY, ops_k(False, True, 128*16, 128, RB*16, 128), 32))
report("both MN (lbo=128,sbo=chunk)", run(X, dY, ops_k(True, True, 128, RB*16, 128, RB*16), 32))
report("both MN (lbo=chunk,sbo=128)", run(X, dY, ops_k(True, True, RB*16, 128, RB*16, 128), 32))
# MN-major with dense 8-chunk blocks? try chunk stride aligned to 128B multiples (RB2 = 136)
RB2 = 136
X2 = np.zeros((32, RB2, 4), np.float32); X2[:, :RB2, :] = X[:, :RB2, :]
Y2 = np.zeros((N//4, RB2, 4), np.float32); Y2[:, :RB2, :] = dY[:, :RB2, :]
def ops2(a_lbo, a_sbo, b_lbo, b_sbo):
    return [[(halo+8*kg)*16, a_lbo, a_sbo, (halo+8*kg)*16, b_lbo, b_sbo, int(kg>0), idesc(128, N, True, True)] for kg in range(16)]
report("both MN RB=136 (lbo=128,sbo=chunk)", run(X2, Y2, ops2(128, RB2*16, 128, RB2*16), 32))
report("both MN RB=136 (lbo=chunk,sbo=128)", run(X2, Y2, ops2(RB2*16, 128, RB2*16, 128), 32))
