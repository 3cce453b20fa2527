"""The algebra behind the engine's pool-fused convolutions (engine.cu prepare_conv): a k x k SAME convolution
followed by a 2x2 average pool equals ONE 3x3 SAME convolution over the space-to-depth input (2x2 pixel blocks in channels)
with pre-averaged weights, and the per-border-class bias of the tiled action/state channels maps from k x k classes to the
3 x 3 classes of the block conv.  float64 on the CPU: the identity is exact up to rounding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def block_weights(w, k):
    """w: (k, k, cin, cout) -> (3, 3, 4*cin, cout); channel order (sy*2+sx)*cin + c (the layout k_pack_s2d / k_lstm_out write)."""
    pad, cin = k // 2, w.shape[2]
    w2 = torch.zeros(3, 3, 4 * cin, w.shape[3], dtype=w.dtype)
    for by in range(3):
        for bx in range(3):
            for sy in range(2):
                for sx in range(2):
                    for py in range(2):
                        for px in range(2):
                            dy, dx = 2 * by + sy - py + pad - 2, 2 * bx + sx - px + pad - 2
                            if 0 <= dy < k and 0 <= dx < k:
                                w2[by, bx, (sy * 2 + sx) * cin:(sy * 2 + sx + 1) * cin] += 0.25 * w[dy, dx]
    return w2


def space_to_depth(x):
    n, c, h, w = x.shape
    out = torch.zeros(n, 4 * c, h // 2, w // 2, dtype=x.dtype)
    for sy in range(2):
        for sx in range(2):
            out[:, (sy * 2 + sx) * c:(sy * 2 + sx + 1) * c] = x[:, :, sy::2, sx::2]
    return out


@pytest.mark.parametrize("k,H,W", [(5, 16, 24), (3, 12, 8), (5, 6, 8)])
def test_conv_then_pool_is_a_block_conv(k, H, W):
    torch.manual_seed(k * 100 + H)
    cin, cout, A, pad = 8, 16, 3, k // 2
    w = torch.randn(k, k, cin + A, cout, dtype=torch.float64)
    x = torch.randn(2, cin, H, W, dtype=torch.float64)
    sa = torch.randn(A, dtype=torch.float64)
    xin = torch.cat([x, sa.view(1, A, 1, 1).expand(2, A, H, W)], 1)
    ref = F.avg_pool2d(F.conv2d(xin, w.permute(3, 2, 0, 1), padding=pad), 2)
    out = F.conv2d(space_to_depth(x), block_weights(w[:, :, :cin], k).permute(3, 2, 0, 1), padding=1)

    # border-class sums of the constant channels for the k x k conv (engine.cu prepare_conv), then their 2x2 averages
    def rep(c, n):
        return c if c < pad else (n - 1 - (k - 1 - c) if c > pad else pad)
    wc = torch.zeros(k * k, A, cout, dtype=torch.float64)
    for cy in range(k):
        for cx in range(k):
            for dy in range(k):
                for dx in range(k):
                    yy, xx = rep(cy, H) + dy - pad, rep(cx, W) + dx - pad
                    if 0 <= yy < H and 0 <= xx < W:
                        wc[cy * k + cx] += w[dy, dx, cin:]
    cls_of = lambda c3, p: p if c3 == 0 else (k - 2 + p if c3 == 2 else pad)
    wc9 = torch.zeros(9, A, cout, dtype=torch.float64)
    for cy in range(3):
        for cx in range(3):
            for py in range(2):
                for px in range(2):
                    wc9[cy * 3 + cx] += 0.25 * wc[cls_of(cy, py) * k + cls_of(cx, px)]
    bc = lambda o, n: 0 if o < 1 else (2 if o >= n - 1 else 1)
    for Y in range(H // 2):
        for X in range(W // 2):
            out[:, :, Y, X] += sa @ wc9[bc(Y, H // 2) * 3 + bc(X, W // 2)]
    assert (out - ref).abs().max().item() < 1e-12
