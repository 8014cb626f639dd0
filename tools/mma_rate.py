"""tcgen05.mma issue-rate probe: cycles per MMA (M=128, K=32 B) for N, accumulator rotation and kind.
Usage: python tools/mma_rate.py   (one B200)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
out = torch.zeros(148, dtype=torch.int64, device=dev)
COUNT = 960
for ctas in (1, 148):
    for tf32 in (0, 1):
        for ncode, n in ((0, 64), (1, 128), (2, 256)):
            for nacc in (1, 2, 4):
                if nacc * n > 512:
                    continue
                mode = ncode | (nacc << 4) | (tf32 << 8)
                for _ in range(2):
                    rc = _lib.lib().riders_debug_tc_mma_rate(mode, COUNT, ctas, C.c_void_p(out.data_ptr()),
                                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
                    _lib.check(rc, "riders_debug_tc_mma_rate")
                torch.cuda.synchronize()
                cyc = out[:ctas].float().mean().item() / COUNT
                print("ctas %3d  kind %s  N %3d  accumulators %d : %6.1f cycles / MMA  (%.0f MAC/cycle/SM)" %
                      (ctas, "tf32" if tf32 else "f16 ", n, nacc, cyc, 128 * n * (8 if tf32 else 16) / cyc))
