"""Hardware check of the tcgen05 layer (csrc/umma.cuh): the two operand conventions the tensor-core stream
kernel relies on, against a CPU product of tf32-rounded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away), keep 10 mantissa bits."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return b.view(np.float32)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(A, B, ops, ncols):
    import gaitk
    L = gaitk.lib()
    Ad = torch.from_numpy(A.ravel()).cuda(); Bd = torch.from_numpy(B.ravel()).cuda()
    od = torch.from_numpy(np.asarray(ops, dtype=np.uint32).view(np.int32).ravel().copy()).cuda()
    D = torch.zeros(128, ncols, device="cuda")
    gaitk._lib.check(L.gaitk_umma_selftest(Ad.data_ptr(), Ad.numel(), Bd.data_ptr(), Bd.numel(), od.data_ptr(), len(ops), ncols,
                                           D.data_ptr(), gaitk._lib.stream_handle()), "umma_selftest")
    torch.cuda.synchronize()
    return D.cpu().numpy()


def test_kmajor_tap_shifted_conv():
    """D[r][n] = sum_tap sum_c X[r + (tap-1) W][c] * Wt[tap][c][n]: A = activation buffer [chunk][row][4] read
    K-major with a 16-byte-per-row start shift, B = weights [tap][chunk][n][4] K-major."""
    rng = np.random.default_rng(0)
    KC, N, W, halo, taps = 4, 16, 2, 4, 3
    RB = 128 + 2 * halo + 1
    X = rng.standard_normal((KC, RB, 4)).astype(np.float32)
    Wt = rng.standard_normal((taps, KC, N, 4)).astype(np.float32)
    ops = []
    for tap in range(taps):
        for kp in range(KC // 2):
            a_off = ((2 * kp) * RB + halo + (tap - 1) * W) * 16
            b_off = ((tap * KC + 2 * kp) * N) * 16
            ops.append([a_off, RB * 16, 128, b_off, N * 16, 128, int(len(ops) > 0), idesc(128, N, False, False)])
    D = run(X, Wt, ops, 32)[:, :N]
    Xr, Wr = tf32(X).astype(np.float64), tf32(Wt).astype(np.float64)
    ref = np.zeros((128, N))
    for tap in range(taps):
        rows = halo + np.arange(128) + (tap - 1) * W
        ref += np.einsum("krc,knc->rn", Xr[:, rows, :], Wr[tap])
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 1e-5, err


def test_mnmajor_tf32_without_swizzle_is_not_usable():
    """Design evidence: a weight gradient needs the TRANSPOSED activations, i.e. MN-major operands.  For tf32
    tcgen05 accepts those only in the SWIZZLE_128B_BASE32B layout (CUTLASS sm100_common.inl:92); with the
    no-swizzle layout of our [chunk][row][4] buffers the MMA does not produce A^T B (observed: accumulator
    left untouched).  Hence the weight gradients run on mma.sync (stream_kernel_tc.cuh)."""
    rng = np.random.default_rng(1)
    halo, RB, N = 4, 137, 16
    X = rng.standard_normal((32, RB, 4)).astype(np.float32)
    dY = rng.standard_normal((N // 4, RB, 4)).astype(np.float32)
    ops = [[(halo + 8 * kg) * 16, 128, RB * 16, (halo + 8 * kg) * 16, 128, RB * 16, int(kg > 0), idesc(128, N, True, True)]
           for kg in range(16)]
    D = run(X, dY, ops, 32)[:, :N]
    rows = halo + np.arange(128)
    Xm = tf32(X).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, 128)
    Ym = tf32(dY).astype(np.float64)[:, rows, :].transpose(1, 0, 2).reshape(128, N)
    ref = Xm.T @ Ym
    assert np.abs(D - ref).max() / np.abs(ref).max() > 0.5
