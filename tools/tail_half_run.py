"""Runs the fp16-map pair (tcgen05 phase convolution + tcgen05 tail) on 2048 points of 240 x 100 for profiling / timing:
   ncu --set full -k regex:decoder_tail_tc -s 3 -c 1 python tools/tail_half_run.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
g = torch.Generator(device=DEV).manual_seed(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x32 = torch.randn((R, 32, 120, 50), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
w2 = torch.randn((64, 32, 2, 2), device=DEV, generator=g) * 0.1
ba, bb = torch.randn(16, device=DEV, generator=g) * 0.1, torch.randn(16, device=DEV, generator=g) * 0.1
wb = torch.randn((16, 16, 3, 3), device=DEV, generator=g) * 0.1
wh = torch.randn((1, 16, 3, 3), device=DEV, generator=g) * 0.1
bp = torch.randn(32, device=DEV, generator=g) * 0.1
x64 = torch.randn((R, 64, 60, 25), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
w64 = torch.randn((128, 64, 2, 2), device=DEV, generator=g) * 0.1
b32 = torch.randn(32, device=DEV, generator=g) * 0.1
cat64 = torch.randn((R, 64, 120, 50), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
skip32 = cat64[:, 32:]
w3 = (torch.randn((32, 64, 3, 3), device=DEV, generator=g) * 0.05).contiguous(memory_format=torch.channels_last)
torch.backends.cudnn.allow_tf32 = True


def lib_l1():
    y4 = torch.nn.functional.conv2d(x64, w64, None, 1, 1)
    return ops.d2s_bias_act_concat(y4, b32, skip32, 0.2, taps=2)
for name, fn in (("phase convolution -> fp16 map", lambda: ops.subpixel_conv_tc_half(x32, w2, ba, 0.2)),
                 ("raw map -> upconv -> fp16 map", lambda: ops.upconv_tc_half(x32, bp, 0.2, w2, ba, 0.2)),
                 ("bias_act of the 32-channel map", lambda: ops.bias_act_(x32, bp, 0.2)),
                 ("deconv1: cuDNN phase conv + d2s", lib_l1),
                 ("deconv1: upconv_tc_f32", lambda: ops.upconv_tc_f32(x64, None, 1.0, w64, b32, 0.2, skip32)),
                 ("deconv1.conv 3x3 64->32: cuDNN TF32", lambda: torch.nn.functional.conv2d(cat64, w3, None, 1, 1)),
                 ("deconv1.conv 3x3 64->32: tcgen05", lambda: ops.conv3x3_tc_f32(cat64, w3)),
                 ("tail on the fp16 map", None)):
    if fn is None:
        yh, _ = ops.subpixel_conv_tc_half(x32, w2, ba, 0.2)
        fn = lambda: ops.decoder_tail_tc_half(yh, wb, bb, wh, 0.2, sigmoid=True)   # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-32s %.4f ms per %d points" % (name, e0.elapsed_time(e1) / 10, R), flush=True)
