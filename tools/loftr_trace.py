"""Phase timeline of the fused tcgen05 LoFTR layer (first tile of CTA 0 and of the middle CTA).
Usage: python tools/loftr_trace.py [N points]   -> per-phase microseconds (SM clock assumed from nvidia-smi)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riders_b200 import _lib, ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
params = {}
for name, shape in (("q_proj.weight", (128, 128)), ("k_proj.weight", (128, 128)), ("v_proj.weight", (128, 128)),
                    ("merge.weight", (128, 128)), ("mlp.0.weight", (256, 256)), ("mlp.2.weight", (128, 256))):
    w = torch.empty(shape)
    torch.nn.init.xavier_uniform_(w, generator=g)
    params[name] = w.to(dev)
for name in ("norm1", "norm2"):
    params[name + ".weight"] = torch.ones(128, device=dev)
    params[name + ".bias"] = torch.zeros(128, device=dev)
x = torch.randn((N, 21, 128), device=dev)
s = torch.randn((N, 21, 128), device=dev)
try:
    mhz = float(subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"]).split()[0])
except Exception:
    mhz = 1965.0
# round-2 kernel: seg 0 = q/k/v, seg 1 = merge, seg 2 = x half of mlp.0 (after the x reload), seg 3 = m half, seg 4 = mlp.2
E_NAMES = ["E0 start", "x (src) split done", "q/k/v ready", "attention quarter 0 done", "attention done (msg in A)", "merge ready",
           "LN1 split done", "h ready", "relu(h) split done", "mlp2 ready", "LN2+store done"]
for label, src in (("self", x), ("cross", s)):
    for _ in range(2):
        ops.loftr_layer(x, src, params, impl=ops.LOFTR_TCGEN05)
    trace = torch.zeros(256, dtype=torch.int64, device=dev)
    _lib.lib().riders_debug_loftr_trace(C.c_void_p(trace.data_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.loftr_layer(x, src, params, impl=ops.LOFTR_TCGEN05)
    b.record()
    torch.cuda.synchronize()
    _lib.lib().riders_debug_loftr_trace(C.c_void_p(0))
    t = trace.cpu().numpy()
    print("== %s layer, N=%d (%d tiles): launch %.1f us (pack + layer), SM clock %.0f MHz" % (label, N, (N + 5) // 6, 1e3 * a.elapsed_time(b), mhz))
    for cta, base in (("CTA 0", 0), ("CTA mid", 128)):
        e, m = t[base:base + 64], t[base + 64:base + 128]
        t0 = e[0]
        print(" %s epilogue thread (us since E0 start):" % cta)
        prev = t0
        for k, name in enumerate(E_NAMES):
            if e[k] == 0:
                continue
            print("   %-28s t=%8.2f  (+%6.2f)" % (name, (e[k] - t0) / mhz, (e[k] - prev) / mhz))
            prev = e[k]
        q = e[20:26]
        if q[0]:
            names = ["TMEM loads of q, k, v landed", "elu + hi/lo split + staging stores", "barrier (all 16 warps staged)", "KV^T + Ksum (mma.sync)", "Q.KV + normaliser + msg stores (mma.sync)", "barrier (staging free)"]
            print(" %s attention quarter 1 (us):  " % cta + "; ".join("%s %.2f" % (nm, (q[i] - (e[3] if i == 0 else q[i - 1])) / mhz) for i, nm in enumerate(names)))
        print(" %s MMA thread: segment [wait A, issue, of which waiting for weights]" % cta)
        for seg in range(5):
            w0, w1, w2, ws = m[seg * 4:seg * 4 + 4]
            if w1 == 0:
                continue
            print("   seg %d: start t=%8.2f  waitA %6.2f  issue %6.2f  (weights wait %6.2f)" %
                  (seg, (w0 - t0) / mhz, (w1 - w0) / mhz, (w2 - w1) / mhz, ws / mhz))
