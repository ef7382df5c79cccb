"""CPU checks of the two identities the wav2vec feature-encoder kernels rest on (no GPU needed):

* conv 0's GroupNorm statistics from 10-tap window moments (w2v_conv0_moments_kernel / w2v_conv0_coef_kernel),
* the pitched activation layout whose overlapping strided rows are exactly the im2col rows of a strided
  "valid" Conv1d (w2v_pitches in wav2vec_host.inc), and the padded per-group layout of the positional conv.
"""
import numpy as np
import torch

KERNELS = (10, 3, 3, 3, 3, 2, 2)
STRIDES = (5, 2, 2, 2, 2, 2, 2)


def frames(L):
    T = [L]
    for k, s in zip(KERNELS, STRIDES):
        T.append((T[-1] - k) // s + 1)
    return T


def pitches(T):
    """mirror of w2v_pitches: P[c] = stride[c] * P[c+1] >= T[c] for c = 1..6, smallest P[7] >= T[7]"""
    p7 = T[7]
    while True:
        P = [0] * 8
        P[7] = p7
        for c in range(6, 0, -1):
            P[c] = STRIDES[c] * P[c + 1]
        if all(P[c] >= T[c] for c in range(1, 8)):
            return P
        p7 += 1


def test_pitches_ten_seconds():
    T = frames(160000)
    assert T[1:] == [31999, 15999, 7999, 3999, 1999, 999, 499]
    assert pitches(T)[1:] == [32000, 16000, 8000, 4000, 2000, 1000, 500]
    for L in (400, 16000, 16001, 47999, 240000, 479999):      # odd lengths, MERT's 24 kHz clips, 30 s
        T = frames(L)
        P = pitches(T)
        assert all(P[c] >= T[c] for c in range(1, 8)) and P[7] - T[7] <= 2


def test_groupnorm_statistics_from_window_moments():
    rng = np.random.default_rng(0)
    L = 16000
    x = rng.standard_normal(L)
    w = rng.standard_normal((512, 10)) * 0.3
    T1 = (L - 10) // 5 + 1
    win = np.lib.stride_tricks.as_strided(x, (T1, 10), (5 * x.strides[0], x.strides[0]))
    y = win @ w.T                                              # [T1, 512] conv output (no bias)
    m = win.mean(0)
    R = win.T @ win / T1
    mean = w @ m
    var = np.einsum("cj,jk,ck->c", w, R - np.outer(m, m), w)
    np.testing.assert_allclose(mean, y.mean(0), rtol=0, atol=1e-12)
    np.testing.assert_allclose(var, y.var(0), rtol=1e-10)
    # and against torch's GroupNorm(512, 512) on the conv output
    conv = torch.nn.functional.conv1d(torch.from_numpy(x)[None, None], torch.from_numpy(w)[:, None], stride=5)
    gn = torch.nn.functional.group_norm(conv, 512, eps=1e-5)[0].T.numpy()
    ours = (y - mean) / np.sqrt(var + 1e-5)
    np.testing.assert_allclose(ours, gn, atol=1e-9)


def test_overlapping_rows_equal_im2col_rows():
    rng = np.random.default_rng(1)
    L, C, B = 4000, 8, 3
    T = frames(L)
    P = pitches(T)
    for c in (1, 5):                                           # a k = 3 and a k = 2 layer
        k, s = KERNELS[c], STRIDES[c]
        x = rng.standard_normal((B, T[c], C))
        buf = np.full((B * P[c] + k, C), np.nan)               # pitch rows (and the slack behind the last clip) are garbage
        for b in range(B):
            buf[b * P[c]:b * P[c] + T[c]] = x[b]
        flat = buf.reshape(-1)
        rows = np.lib.stride_tricks.as_strided(flat, (B * P[c + 1], k * C), (s * C * flat.strides[0], flat.strides[0]))
        w = rng.standard_normal((C, C, k))
        ref = torch.nn.functional.conv1d(torch.from_numpy(x).transpose(1, 2), torch.from_numpy(w), stride=s).transpose(1, 2).numpy()
        wk = w.transpose(0, 2, 1).reshape(C, k * C)            # column = tap * Cin + c, as packed for the GEMM
        for b in range(B):
            got = rows[b * P[c + 1]:b * P[c + 1] + T[c + 1]] @ wk.T
            np.testing.assert_allclose(got, ref[b], atol=1e-12)


def test_positional_conv_padded_group_layout():
    rng = np.random.default_rng(2)
    S, d, groups, k, B = 37, 32, 4, 128, 2
    cg = d // groups
    h = rng.standard_normal((B, S, d))
    w = rng.standard_normal((d, cg, k)) * 0.1
    conv = torch.nn.functional.conv1d(torch.from_numpy(h).transpose(1, 2), torch.from_numpy(w), padding=k // 2, groups=groups)
    ref = conv[:, :, :-1].transpose(1, 2).numpy()              # Wav2Vec2SamePadLayer drops the last step of an even kernel
    Pp = S + 128
    slab = B * Pp + 128
    for g in range(groups):
        a = np.zeros((slab, cg))
        for b in range(B):
            a[b * Pp + 64:b * Pp + 64 + S] = h[b, :, g * cg:(g + 1) * cg]
        flat = a.reshape(-1)
        rows = np.lib.stride_tricks.as_strided(flat, (B * Pp, k * cg), (cg * flat.strides[0], flat.strides[0]))
        wg = w[g * cg:(g + 1) * cg].transpose(0, 2, 1).reshape(cg, k * cg)      # column = tap * cg + ci
        for b in range(B):
            got = rows[b * Pp:b * Pp + S] @ wg.T
            np.testing.assert_allclose(got, ref[b, :, g * cg:(g + 1) * cg], atol=1e-10)
