"""Host-side weight preparation of the decoder fast paths (CPU only, no kernels): the sub-pixel phase kernels, the
mma.m16n8k8 B-fragment orders and the tcgen05 K-major SWIZZLE_128B slab image are pure index algebra in torch."""
import numpy as np
import torch
import torch.nn.functional as F

from riders_b200 import ops
from riders_b200.rcnet import networks


def _tf32_rna(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _block(cin=8, cskip=4, cout=4):
    torch.manual_seed(3)
    blk = networks.DecoderBlock(cin, cskip, cout, 'kaiming_uniform', 'leaky_relu', False).eval()
    for m in blk.modules():
        if isinstance(m, networks.Conv2d):
            m.fold_batch_norm_()
    return blk.fold_upsample_()


def test_subpixel_phase_kernels_equal_resize_then_conv():
    """DecoderBlock.fold_upsample_: both forms of the phase kernels reproduce conv3x3(nearest 2x resize(x)) (the kernels are float32 pre-sums of the taps, evaluated in fp64)."""
    blk = _block()
    w = blk.deconv.conv.conv.weight.detach().double()
    x = torch.randn(2, 8, 5, 7, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2), w, None, 1, 1)
    co = w.shape[0]
    y3 = F.conv2d(x, blk._w4.double(), None, 1, 1)                        # (N, 4 co, h, w), phase-major channels
    y2 = F.conv2d(x, blk._w2.double(), None, 1, 1)                        # (N, 4 co, h + 1, w + 1): phase (a, b) at (i + a, j + b)
    assert tuple(blk._w4.shape) == (4 * co, 8, 3, 3) and tuple(blk._w2.shape) == (4 * co, 8, 2, 2)
    out3, out2 = torch.empty_like(ref), torch.empty_like(ref)
    for a in (0, 1):
        for b in (0, 1):
            ph = slice((2 * a + b) * co, (2 * a + b + 1) * co)
            out3[:, :, a::2, b::2] = y3[:, ph]
            out2[:, :, a::2, b::2] = y2[:, ph, a:a + 5, b:b + 7]
    np.testing.assert_allclose(out3.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out2.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_odd_target_first_column_and_row_kernels():
    """2n+1 targets: torch's nearest resize picks source max(j - 1, 0) >> 1; the first output column (row) comes from the
    3x1 (1x3) phase kernels, everything else from the phase kernels on the replicate-padded map read one pixel further."""
    blk = _block()
    w = blk.deconv.conv.conv.weight.detach().double()
    co = w.shape[0]
    x = torch.randn(2, 8, 5, 4, dtype=torch.float64)
    for sy, sx in ((0, 1), (1, 0)):
        H, W = 2 * 5 + sy, 2 * 4 + sx
        ref = F.conv2d(F.interpolate(x, size=(H, W)), w, None, 1, 1)
        xp = F.pad(x, (sx, 0, sy, 0), mode='replicate')
        y2 = F.conv2d(xp, blk._w2.double(), None, 1, 1)
        out = torch.empty_like(ref)
        for Y in range(H):
            for X in range(W):
                if (sx and X == 0) or (sy and Y == 0):
                    continue
                ky, kx = Y + sy, X + sx
                ph = 2 * (ky & 1) + (kx & 1)
                out[:, :, Y, X] = y2[:, ph * co:(ph + 1) * co, (ky >> 1) + (ky & 1), (kx >> 1) + (kx & 1)]
        if sx:
            yf = F.conv2d(x[:, :, :, 0:1], blk._wcol.double(), None, 1, (1, 0))          # (N, 2 co, h, 1): channel a * co + c
            for Y in range(H):
                out[:, :, Y, 0] = yf[:, (Y & 1) * co:((Y & 1) + 1) * co, Y >> 1, 0]
        else:
            yf = F.conv2d(x[:, :, 0:1, :], blk._wrow.double(), None, 1, (0, 1))          # (N, 2 co, 1, w): channel b * co + c
            for X in range(W):
                out[:, :, 0, X] = yf[:, (X & 1) * co:((X & 1) + 1) * co, 0, X >> 1]
        np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_tail_weight_fragments_follow_the_mma_layout():
    """ops._tail_weight_fragments: [K step][lane][N half][2] with b0 = W[8 nt + g][8 j + t][ky][kx], b1 at ci + 4,
    lane = 4 g + t, K step = (ky * 3 + kx) * 2 + j; TF32-rounded."""
    torch.manual_seed(0)
    w = torch.randn(16, 16, 3, 3)
    f = ops._tail_weight_fragments(w)
    assert tuple(f.shape) == (18, 32, 2, 2)
    wr = _tf32_rna(w)
    for s, lane, nt, e in ((0, 0, 0, 0), (5, 13, 1, 1), (17, 31, 1, 0), (8, 6, 0, 1)):
        tap, j, g, t = s // 2, s % 2, lane // 4, lane % 4
        assert f[s, lane, nt, e] == wr[nt * 8 + g, 8 * j + t + 4 * e, tap // 3, tap % 3]
    assert ops._tail_weight_fragments(w) is f                             # cached per tensor object and version
    w.mul_(2.0)
    assert ops._tail_weight_fragments(w) is not f                         # in-place update -> rebuilt
    assert ops._tail_weight_fragments(w.clone()) is not ops._tail_weight_fragments(w)


def test_subpixel_fragments_and_tcgen05_slabs():
    torch.manual_seed(1)
    w2 = torch.randn(64, 32, 2, 2)
    wr = _tf32_rna(w2)
    f = ops._subpixel_weight_fragments(w2)                                # [K step][N tile pair][lane][4]
    assert tuple(f.shape) == (16, 4, 32, 4)
    for ks, npair, lane, e in ((0, 0, 0, 0), (7, 3, 21, 3), (15, 1, 31, 2), (9, 2, 5, 1)):
        tap, g, t = ks // 4, lane // 4, lane % 4
        assert f[ks, npair, lane, e] == wr[(2 * npair + e // 2) * 8 + g, (ks % 4) * 8 + t + 4 * (e % 2), tap // 2, tap % 2]
    p = ops._subpixel_tc_weights(w2)                                      # four K-major SWIZZLE_128B slabs of 64 rows x 128 B
    assert p.numel() == 4 * 64 * 32
    seen = torch.zeros(p.numel(), dtype=torch.bool)
    for tap in range(4):
        for n in (0, 1, 7, 8, 37, 63):
            for k in (0, 3, 4, 17, 31):
                off = tap * 2048 + (n >> 3) * 256 + (n & 7) * 32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3)
                assert p[off] == wr[n, k, tap // 2, tap % 2]
                seen[off] = True
    assert seen.sum() == 4 * 6 * 5                                        # the sampled offsets are all distinct (a permutation)
    assert torch.equal(torch.sort(p).values, torch.sort(wr.permute(2, 3, 0, 1).reshape(-1)).values)


def test_tcgen05_tail_weight_pack_is_the_unswizzled_k_major_image():
    """ops._tail_tc_weights: the 6144-byte fp16 B-operand image of riders_decoder_tail_tc_nhwc.  Emulate what the tensor core
    computes from it - D[m][n] = sum_k A[m][k] B[n][k] with B read through the unswizzled K-major descriptor (core matrix =
    8 rows x 16 bytes, 256 B between 8-row groups, 128 B between the two K chunks) and A = the rows of an operand row slot
    shifted by dx pixels - and check that accumulating the N blocks into z rows r+1, r, r-1 IS the 3x3 convolution (and the
    head's N = 16 image its three kernel rows)."""
    torch.manual_seed(7)
    wb, wh = torch.randn(16, 16, 3, 3) * 0.2, torch.randn(1, 16, 3, 3) * 0.2
    pack = ops._tail_tc_weights(wb, wh)
    assert pack.dtype == torch.float16 and pack.numel() * 2 == 6144
    raw = pack.numpy()

    def b_matrix(byte0, rows):                       # rows x 16 matrix as the descriptor (LBO 128, SBO 256) addresses it
        m = np.zeros((rows, 16), np.float32)
        for n in range(rows):
            for k in range(16):
                off = byte0 + (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2
                m[n, k] = raw[off // 2]
        return m
    H, W = 6, 9
    y = torch.randn(1, 16, H, W).half().float()
    wbh, whh = wb.half().float(), wh.half().float()
    ref_z = F.conv2d(y, wbh, None, 1, 1)[0]                                    # (16, H, W)
    ypad = F.pad(y, (1, 1, 0, 0))[0].permute(1, 2, 0).numpy()                 # (H, W + 2, 16): column x + dx of the padded row
    z = np.zeros((H, W, 16), np.float32)
    for r in range(H):                                                         # y row r feeds z rows r + 1 - dy'
        for dx in range(3):
            B = b_matrix(dx * 1536, 48)                                        # n = 16 dy' + co
            D = ypad[r, dx:dx + W, :] @ B.T                                    # (W, 48)
            for dyp in range(3):
                q = r + 1 - dyp
                if 0 <= q < H:
                    z[q] += D[:, 16 * dyp:16 * dyp + 16]
    np.testing.assert_allclose(z.transpose(2, 0, 1), ref_z.numpy(), rtol=1e-5, atol=1e-5)
    zt = torch.randn(1, 16, H, W).half().float()
    ref_o = F.conv2d(zt, whh, None, 1, 1)[0, 0]
    zpad = F.pad(zt, (1, 1, 0, 0))[0].permute(1, 2, 0).numpy()
    Dh = np.zeros((H, W, 16), np.float32)
    for q in range(H):
        for dx in range(3):
            Dh[q] += zpad[q, dx:dx + W, :] @ b_matrix(4608 + dx * 512, 16).T
    assert not Dh[:, :, 3:].any()                                              # rows 3..15 of the head image are zero
    out = np.zeros((H, W), np.float32)
    for o in range(H):                                                         # logits[o] = Dh[o-1][0] + Dh[o][1] + Dh[o+1][2]
        for dyp in range(3):
            q = o - 1 + dyp
            if 0 <= q < H:
                out[o] += Dh[q, :, dyp]
    np.testing.assert_allclose(out, ref_o.numpy(), rtol=1e-5, atol=1e-5)
    assert ops._tail_tc_weights(wb, wh) is pack                                # cached per parameter version
    wh.data.mul_(2.0)
    wh._version                                                                 # (in-place edits bump the version)
    wh.mul_(1.0)
    assert ops._tail_tc_weights(wb, wh) is not pack


def test_upconv_weight_pack_and_its_cache():
    """ops._upconv_tc_weights: K-major unswizzled fp16 image [tap 2u+v][K half][n // 8][k // 8][n % 8][k % 8] of W2[n][16 kh + k][u][v];
    the phase convolution it encodes (every tap a shifted view of the padded row) equals conv2d(x, W2, padding 1); and the pack
    does not share a cache slot with the mma.sync fragments of the same parameter."""
    torch.manual_seed(11)
    w2 = torch.randn(64, 32, 2, 2) * 0.2
    pack = ops._upconv_tc_weights(w2)
    assert pack.dtype == torch.float16 and pack.numel() * 2 == 16384
    raw = pack.numpy()

    def b_matrix(t, kh):
        m = np.zeros((64, 16), np.float32)
        for n in range(64):
            for k in range(16):
                off = (t * 2 + kh) * 2048 + (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2
                m[n, k] = raw[off // 2]
        return m
    h, w = 3, 5
    x = torch.randn(1, 32, h, w).half().float()
    ref = F.conv2d(x, w2.half().float(), None, 1, 1)[0]                        # (64, h + 1, w + 1)
    xp = F.pad(x, (1, 1, 1, 1))[0].permute(1, 2, 0).numpy()                    # staged rows: row i + 1, position j + 1
    out = np.zeros((h + 1, w + 1, 64), np.float32)
    for I in range(h + 1):
        for t in range(4):
            u, v = t // 2, t % 2
            for kh in range(2):
                out[I] += xp[I + u, v:v + w + 1, 16 * kh:16 * kh + 16] @ b_matrix(t, kh).T
    np.testing.assert_allclose(out.transpose(2, 0, 1), ref.numpy(), rtol=1e-5, atol=1e-5)
    frag = ops._subpixel_weight_fragments(w2)
    assert frag.dtype == torch.float32 and ops._upconv_tc_weights(w2) is pack and ops._subpixel_weight_fragments(w2) is frag


def test_conv3x3_weight_pack_emulates_the_nine_tap_views():
    """ops._conv3x3_tc_weights: [tap 3 ky + kx][K step][n // 8][k // 8][n % 8][k % 8] fp16; summing, over the nine taps and four K
    steps, the products of the (row + ky, pixel + kx) views of the zero-padded map with these matrices is conv2d(x, W, padding 1)."""
    torch.manual_seed(5)
    wt = torch.randn(32, 64, 3, 3) * 0.1
    pack = ops._conv3x3_tc_weights(wt)
    assert pack.dtype == torch.float16 and pack.numel() * 2 == 9 * 4 * 32 * 32
    raw = pack.numpy()

    def b_matrix(t, kh):
        m = np.zeros((32, 16), np.float32)
        for n in range(32):
            for k in range(16):
                off = (t * 4 + kh) * 1024 + (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2
                m[n, k] = raw[off // 2]
        return m
    h, w = 4, 6
    x = torch.randn(1, 64, h, w).half().float()
    ref = F.conv2d(x, wt.half().float(), None, 1, 1)[0]
    xp = F.pad(x, (1, 1, 1, 1))[0].permute(1, 2, 0).numpy()                    # staged rows: row i + 1, position j + 1
    out = np.zeros((h, w, 32), np.float32)
    for I in range(h):
        for t in range(9):
            ky, kx = t // 3, t % 3
            for kh in range(4):
                out[I] += xp[I + ky, kx:kx + w, 16 * kh:16 * kh + 16] @ b_matrix(t, kh).T
    np.testing.assert_allclose(out.transpose(2, 0, 1), ref.numpy(), rtol=1e-5, atol=1e-5)


def test_decoder_block_takes_a_raw_input_with_its_activation():
    """DecoderBlock.forward(x_act=(bias, slope)): ``x`` is the RAW output of the previous block's convolution (the hand-over
    ops.upconv_tc_f32 uses on the GPU); when that kernel does not take the call the block applies bias + LeakyReLU itself -
    on the CPU that is the only path: same result as feeding the activated map."""
    torch.manual_seed(9)
    blk = networks.DecoderBlock(8, 4, 4, 'kaiming_uniform', 'leaky_relu', False).eval()
    x = torch.randn(2, 8, 5, 7)
    skip = torch.randn(2, 4, 10, 14)
    b = torch.randn(8) * 0.3
    with torch.no_grad():
        ref = blk(F.leaky_relu(x + b.view(1, -1, 1, 1), 0.2), skip)
        got = blk(x, skip, x_act=(b, 0.2))
        same = blk(x, skip, x_act=(None, 1.0))
        plain = blk(x, skip)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(same, plain) and not blk.upconv_tc and not blk.last_used_upconv_tc
    assert not blk.upconv_tc_ok(x.shape, (10, 14))
