#!/usr/bin/env python
"""Gradient-kernel variants at a small shape (default BASELINE.json configs[1]: B=4096, D=768): split-K of the ragged
last wave x column-tile width, CUDA-event time per launch (interleaved repetitions, burst clocks — these kernels are too
short to be power-capped)."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributed_sigmoid_loss_b200 import SigmoidLossEngine, _capi

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--D", type=int, default=768)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
txt = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
tp = torch.tensor([math.log(10.0)], device=dev)
b = torch.tensor([-10.0], device=dev)
eng = SigmoidLossEngine(a.B, a.D, dev)
variants = [(sk, tn) for tn in (256, 128) for sk in (0, -1, 2, 3, 4)]
res = {v: [] for v in variants}
for rep in range(3):
    for (sk, tn) in variants:
        eng.set_option(_capi.SIGLIP_OPT_SPLIT_K, sk)
        eng.set_option(_capi.SIGLIP_OPT_GRAD_TILE_N, tn)
        for _ in range(10):
            eng.fwd_bwd(img, txt, tp, b, torch.bfloat16)
        torch.cuda.synchronize()
        eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            eng.fwd_bwd(img, txt, tp, b, torch.bfloat16)
        e1.record()
        torch.cuda.synchronize()
        lm, ln, gm, gn = eng.kernel_times()
        eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 0)
        # back-to-back without the timing events between the launches
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.iters):
            eng.fwd_bwd(img, txt, tp, b, torch.bfloat16)
        f1.record()
        torch.cuda.synchronize()
        res[(sk, tn)].append((lm / ln * 1e3, gm / gn * 1e3, e0.elapsed_time(e1) / a.iters * 1e3,
                              f0.elapsed_time(f1) / a.iters * 1e3))
flops = 4.0 * a.B * a.B * a.D
print(f"B={a.B} D={a.D}: per launch, best of 3 (us): loss kernel | gradient kernel (TFLOP/s) | step with events | step without")
for (sk, tn) in variants:
    r = res[(sk, tn)]
    lo = min(x[0] for x in r)
    gr = min(x[1] for x in r)
    print(f"  split_k={sk:2d} tile_n={tn}: {lo:6.1f} | {gr:6.1f} ({flops / gr / 1e6:5.0f}) | {min(x[2] for x in r):6.1f} | "
          f"{min(x[3] for x in r):6.1f}")
eng.close()
