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


# ------------------------------------------------------------------------------------------------ kind::f16 / bf16
def bf16_bits(x):
    """fp32 -> bf16 bit pattern, round to nearest even (cvt.rn.bf16.f32)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def bf16_val(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)


def idesc_bf16(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run_bf16(A16, B16, ops, ncols):
    import gaitk
    L = gaitk.lib()
    Ad = torch.from_numpy(A16.ravel().view(np.int16).copy()).cuda(); Bd = torch.from_numpy(B16.ravel().view(np.int16).copy()).cuda()
    od = torch.from_numpy(np.asarray(ops, dtype=np.uint32).view(np.int32).ravel().copy()).cuda()
    D = torch.zeros(128, ncols, device="cuda")
    gaitk._lib.check(L.gaitk_umma_selftest_bf16(Ad.data_ptr(), Ad.numel(), Bd.data_ptr(), Bd.numel(), od.data_ptr(), len(ops), ncols,
                                                D.data_ptr(), gaitk._lib.stream_handle()), "umma_selftest_bf16")
    torch.cuda.synchronize()
    return D.cpu().numpy()


@pytest.mark.parametrize("N", [16, 24, 32])
def test_bf16_kmajor_tap_shifted_conv_with_zero_plane(N):
    """Forward / data-gradient convolutions of the split-bf16 kernel: A = activation planes [chunk8][row][8 x bf16] read
    K-major with a 16-byte-per-row start shift; a K step of 16 channels pairs plane 2k with plane 2k+1 through LBO, and
    an odd plane count pairs the last plane with a shared ZERO plane somewhere else in shared memory (LBO = distance).
    B = weights [tap][chunk8][n][8] K-major.  N = 24 checks that steps of 8 are legal at M = 128."""
    rng = np.random.default_rng(0)
    NCH, W, halo, taps = 3, 2, 4, 3
    RB = 128 + 2 * halo
    X = rng.standard_normal((NCH + 2, RB, 8)).astype(np.float32)
    X[NCH] = np.nan                       # a plane that must never be read
    X[NCH + 1] = 0.0                      # the shared zero plane
    Wt = rng.standard_normal((taps, 4, N, 8)).astype(np.float32)
    Wt[:, 3] = rng.standard_normal((taps, N, 8))      # weights against the zero plane: arbitrary, contribute nothing
    ops = []
    for tap in range(taps):
        for ks in range(2):
            a_off = ((2 * ks) * RB + halo + (tap - 1) * W) * 16
            lbo = RB * 16 if ks == 0 else (NCH + 1 - 2) * RB * 16          # plane 2 pairs with the zero plane (index 4)
            b_off = ((tap * 4 + 2 * ks) * N) * 16
            ops.append([a_off, lbo, 128, b_off, N * 16, 128, int(len(ops) > 0), idesc_bf16(128, N, False, False)])
    D = run_bf16(bf16_bits(X), bf16_bits(Wt), ops, 32)[:, :N]
    Xr = bf16_val(bf16_bits(X)).astype(np.float64); Wr = bf16_val(bf16_bits(Wt)).astype(np.float64)
    ref = np.zeros((128, N))
    for tap in range(taps):
        rows = halo + np.arange(128) + (tap - 1) * W
        ref += np.einsum("krc,knc->rn", Xr[:NCH, rows, :], Wr[tap, :NCH])
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 1e-5, err


def test_bf16_mnmajor_weight_gradient_with_tap_shift_and_column_offsets():
    """Weight gradients of the split-bf16 kernel on tcgen05: D_tap[m][n] = sum_r IN[r + (tap-1) W][m] * DOUT[r][n] with
    BOTH operands MN-major straight out of the activation planes (rows = K, 16 rows per instruction, LBO = 128 B between
    the two groups of 8 rows, SBO = plane stride between 8-channel cores).  M = 128 reads 16 planes: the planes past the
    real channels hold garbage (here NaN) and only pollute accumulator rows that are never read.  Every tap accumulates
    over 8 K steps into its own TMEM column block."""
    rng = np.random.default_rng(1)
    W, halo, taps, N = 2, 4, 3, 16
    RB = 128 + 2 * halo
    INr, DOr = 3, 2                                           # real planes: 24 input channels, 16 output channels
    IN = rng.standard_normal((16, RB, 8)).astype(np.float32); IN[INr:] = np.nan
    DO = rng.standard_normal((DOr, RB, 8)).astype(np.float32)
    ops = []
    for tap in range(taps):
        for ks in range(8):
            a_off = (halo + 16 * ks + (tap - 1) * W) * 16
            b_off = (halo + 16 * ks) * 16
            ops.append([a_off, 128, RB * 16, b_off, 128, RB * 16, int(ks > 0) | ((tap * 2) << 8), idesc_bf16(128, N, True, True)])
    D = run_bf16(bf16_bits(IN), bf16_bits(DO), ops, 64)
    Ir = bf16_val(bf16_bits(IN[:INr])).astype(np.float64); Dr = bf16_val(bf16_bits(DO)).astype(np.float64)
    for tap in range(taps):
        rows = halo + np.arange(128)
        A = Ir[:, rows + (tap - 1) * W, :].transpose(1, 0, 2).reshape(128, INr * 8)      # [r][m]
        Bm = Dr[:, rows, :].transpose(1, 0, 2).reshape(128, DOr * 8)                     # [r][n]
        ref = A.T @ Bm
        got = D[:INr * 8, tap * 16:tap * 16 + N]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (tap, err)


def test_bf16_split_three_pass_product_is_fp32_grade():
    """x = hi + lo (both bf16): hi*Whi + lo*Whi + hi*Wlo reproduces the fp32 product to ~2^-16 -- the arithmetic of the
    'bf16x3' stream kernel -- where a single tf32 or bf16 pass is at 5e-4 / 4e-3."""
    rng = np.random.default_rng(2)
    RB, N = 128, 16
    X = rng.standard_normal((2, RB, 8)).astype(np.float32)
    Wt = rng.standard_normal((2, N, 8)).astype(np.float32)
    xh = bf16_bits(X); xl = bf16_bits(X - bf16_val(xh)); wh = bf16_bits(Wt); wl = bf16_bits(Wt - bf16_val(wh))
    A = np.stack([xh, xl]); Bm = np.stack([wh, wl])                  # [part][plane][row][8]
    pa, pb = 2 * RB * 16, 2 * N * 16
    ops = []
    for (ia, ib) in ((0, 0), (1, 0), (0, 1)):
        ops.append([ia * pa, RB * 16, 128, ib * pb, N * 16, 128, int(len(ops) > 0), idesc_bf16(128, N, False, False)])
    D = run_bf16(A, Bm, ops, 32)[:, :N]
    ref = np.einsum("krc,knc->rn", X.astype(np.float64), Wt.astype(np.float64))
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 5e-5, err
