"""First-light check of the tcgen05 decoder tail (riders_decoder_tail_tc_nhwc) against torch convolutions of the fp16-rounded
operands, with an error map summary (which rows / columns / points are off) - run under a hard timeout on a GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import ops  # noqa: E402

F = torch.nn.functional
DEV = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def h16(x):
    return x.half().float()


def case(N, h, w, taps, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    t = 1 if taps == 2 else 0
    y4 = torch.randn((N, 64, h + t, w + t), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=taps).contiguous()
    z = F.leaky_relu(F.conv2d(h16(y), h16(wb), None, 1, 1) + bb.view(1, -1, 1, 1), 0.2)
    ref = F.conv2d(h16(z), h16(wh), None, 1, 1)
    ref32 = F.conv2d(F.leaky_relu(F.conv2d(y, wb, bb, 1, 1), 0.2), wh, None, 1, 1)
    got = ops.decoder_tail_tc(y4, ba, wb, bb, wh, 0.2, taps=taps)
    torch.cuda.synchronize()
    scale = float(ref32.abs().max())
    err = (got - ref).abs()
    bad = err > 1e-3 * scale
    print("N=%d h=%d w=%d taps=%d: scale %.3f  max err %.3e (vs fp16-operand ref)  %.3e (vs fp32)  bad fraction %.4f" %
          (N, h, w, taps, scale, float(err.max()), float((got - ref32).abs().max()), float(bad.float().mean())), flush=True)
    if bool(bad.any()):
        rows = bad.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        cols = bad.any(dim=2).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        pts = bad.flatten(1).any(dim=1).nonzero().flatten().tolist()
        print("   bad rows", rows[:40], "... cols", cols[:40], "... points", pts[:10], flush=True)
        print("   got[0,0,:4,:6]\n", got[0, 0, :4, :6].cpu().numpy(), "\n   ref\n", ref[0, 0, :4, :6].cpu().numpy(), flush=True)
    return not bool(bad.any())


def case_half(N, h, w, seed=0):
    """fp16-map path: tcgen05 phase convolution finishing the up-convolution (yh) + tcgen05 tail reading it."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    ba, bb = torch.randn(16, device=DEV, generator=g) * 0.3, torch.randn(16, device=DEV, generator=g) * 0.3
    wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.15
    wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.2
    xt = (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)                     # the tensor core truncates to TF32
    w2t = ((w2.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    y4 = F.conv2d(xt, w2t, None, 1, 1).contiguous(memory_format=torch.channels_last)
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=2).contiguous()                       # (N,16,2h,2w)
    yh, flag = ops.subpixel_conv_tc_half(x, w2, ba, 0.2)
    torch.cuda.synchronize()
    yref = y.view(N, 2, 8, 2 * h, 2 * w).permute(0, 3, 1, 4, 2)                               # [n][Y][plane][X][k]
    ey = (yh.float() - yref).abs()
    print("half N=%d h=%d w=%d: yh max err %.3e (scale %.2f), flag %d" % (N, h, w, float(ey.max()), float(yref.abs().max()), int(flag.item())), flush=True)
    z = F.leaky_relu(F.conv2d(h16(y), h16(wb), None, 1, 1) + bb.view(1, -1, 1, 1), 0.2)
    ref = F.conv2d(h16(z), h16(wh), None, 1, 1)
    got = ops.decoder_upconv_tail_tc(x, w2, ba, wb, bb, wh, 0.2)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err = (got - ref).abs()
    bad = err > 2e-3 * scale
    print("   logits scale %.3f  max err %.3e  bad fraction %.4f" % (scale, float(err.max()), float(bad.float().mean())), flush=True)
    if bool(bad.any()):
        rows = bad.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        cols = bad.any(dim=2).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        print("   bad rows", rows[:40], "... cols", cols[:40], flush=True)
    return (not bool(bad.any())) and float(ey.max()) < 4e-3 * float(yref.abs().max())


def case_upconv(N, h, w, seed=0, raw=True):
    """upconv_tc_half: raw 32-channel map (+ previous bias / LeakyReLU) -> fp16 operand map, against torch on fp16-rounded operands."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((N, 32, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
    bp = torch.randn(32, device=DEV, generator=g) * 0.3 if raw else None
    sp = 0.2 if raw else 1.0
    ba = torch.randn(16, device=DEV, generator=g) * 0.3
    xa = F.leaky_relu(x + bp.view(1, -1, 1, 1), sp) if raw else x
    y4 = F.conv2d(h16(xa), h16(w2), None, 1, 1).contiguous(memory_format=torch.channels_last)
    y = ops.d2s_bias_act_concat(y4, ba, None, 0.2, taps=2).contiguous()
    yref = y.view(N, 2, 8, 2 * h, 2 * w).permute(0, 3, 1, 4, 2)
    yh, flag = ops.upconv_tc_half(x, bp, sp, w2, ba, 0.2)
    torch.cuda.synchronize()
    ey = (yh.float() - yref).abs()
    sc = float(yref.abs().max())
    bad = ey > 2.5e-3 * sc
    print("upconv N=%d h=%d w=%d raw=%d: yh max err %.3e (scale %.2f) bad fraction %.5f flag %d" %
          (N, h, w, raw, float(ey.max()), sc, float(bad.float().mean()), int(flag.item())), flush=True)
    if bool(bad.any()):
        idx = bad.nonzero()[:8].tolist()
        print("   first bad [n, Y, plane, X, k]:", idx, flush=True)
        rows = bad.any(dim=4).any(dim=3).any(dim=2).any(dim=0).nonzero().flatten().tolist()
        cols = bad.any(dim=4).any(dim=2).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        print("   bad rows", rows[:40], "cols", cols[:40], flush=True)
    return not bool(bad.any())


def case_upconv_f32(N, h, w, Cs=32, raw=False, seed=0):
    """upconv_tc_f32 (64 -> 32 channels, float32 into the concat buffer) against torch on fp16-rounded operands."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((N, 64, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w2 = torch.randn((128, 64, 2, 2), device=DEV, generator=g) * 0.08
    bp = torch.randn(64, device=DEV, generator=g) * 0.3 if raw else None
    sp = 0.2 if raw else 1.0
    ba = torch.randn(32, device=DEV, generator=g) * 0.3
    skip = torch.randn((N, Cs, 2 * h, 2 * w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last) if Cs else None
    xa = F.leaky_relu(x + bp.view(1, -1, 1, 1), sp) if raw else x
    y4 = F.conv2d(h16(xa), h16(w2), None, 1, 1).contiguous(memory_format=torch.channels_last)
    ref = ops.d2s_bias_act_concat(y4, ba, skip, 0.2, taps=2)
    got, flag = ops.upconv_tc_f32(x, bp, sp, w2, ba, 0.2, skip)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    sc = float(ref[:, :32].abs().max())
    bad = err > 1e-4 * sc
    print("upconv_f32 N=%d h=%d w=%d Cs=%d raw=%d: max err %.3e (scale %.2f) bad fraction %.5f flag %d shape %s" %
          (N, h, w, Cs, raw, float(err.max()), sc, float(bad.float().mean()), int(flag.item()), tuple(got.shape)), flush=True)
    if bool(bad.any()):
        print("   first bad [n, c, Y, X]:", bad.nonzero()[:8].tolist(), flush=True)
        rows = bad.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        cols = bad.any(dim=2).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        chans = bad.any(dim=3).any(dim=2).any(dim=0).nonzero().flatten().tolist()
        print("   bad rows", rows[:40], "cols", cols[:40], "channels", chans[:40], flush=True)
    return not bool(bad.any())


def case_conv3(N, h, w, seed=0):
    """conv3x3_tc_f32 (64 -> 32, raw output) against torch on fp16-rounded operands."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((N, 64, h, w), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    wt = torch.randn((32, 64, 3, 3), device=DEV, generator=g) * 0.05
    ref = F.conv2d(h16(x), h16(wt), None, 1, 1)
    got, flag = ops.conv3x3_tc_f32(x, wt)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    sc = float(ref.abs().max())
    bad = err > 1e-4 * sc
    print("conv3x3 N=%d h=%d w=%d: max err %.3e (scale %.2f) bad fraction %.5f flag %d" % (N, h, w, float(err.max()), sc, float(bad.float().mean()), int(flag.item())), flush=True)
    if bool(bad.any()):
        print("   first bad [n, c, Y, X]:", bad.nonzero()[:8].tolist(), flush=True)
        rows = bad.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        cols = bad.any(dim=2).any(dim=1).any(dim=0).nonzero().flatten().tolist()
        print("   bad rows", rows[:40], "cols", cols[:40], flush=True)
    return not bool(bad.any())


if __name__ == "__main__":
    ok = True
    for N, h, w, taps in ((1, 8, 5, 3), (3, 7, 9, 2), (2, 13, 30, 3), (3, 120, 50, 2), (2, 40, 70, 3), (40, 120, 50, 2)):
        ok = case(N, h, w, taps) and ok
    for N, h, w in ((1, 8, 5), (3, 7, 9), (2, 13, 30), (3, 120, 50), (40, 120, 50)):
        ok = case_half(N, h, w) and ok
    for N, h, w, raw in ((1, 8, 5, True), (3, 7, 9, False), (2, 13, 30, True), (3, 120, 50, True), (200, 120, 50, True), (2, 31, 54, True)):
        ok = case_upconv(N, h, w, raw=raw) and ok
    for N, h, w, Cs, raw in ((1, 8, 5, 32, False), (3, 7, 9, 0, True), (2, 13, 26, 8, True), (3, 60, 25, 32, False), (300, 60, 25, 32, False)):
        ok = case_upconv_f32(N, h, w, Cs, raw) and ok
    for N, h, w in ((1, 8, 5), (3, 7, 9), (2, 13, 54), (3, 120, 50), (300, 120, 50)):
        ok = case_conv3(N, h, w) and ok
    print("ALL OK" if ok else "MISMATCH", flush=True)
