#!/usr/bin/env python
"""Kernel times at the headline shape for different logit regimes: the loss epilogue has a fast path (whole 32x32 slab
z < -4.2, the regime of a SigLIP batch: bias ~ -10) and a general path (any z)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sigmoid_loss_b200 import SigmoidLossEngine, _capi

dev = torch.device("cuda", 0)
B, D = 16384, 1024
g = torch.Generator().manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
eng = SigmoidLossEngine(B, D, dev)
for (name, t, b) in [("init  t=10  b=-10 (all fast path)", 10.0, -10.0), ("trained-like t=100 b=-12", 100.0, -12.0),
                     ("warm  t=5   b=0   (all general path)", 5.0, 0.0), ("init again", 10.0, -10.0)]:
    tp, bs = torch.tensor([math.log(t)], device=dev), torch.tensor([b], device=dev)
    for _ in range(5):
        eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    torch.cuda.synchronize()
    eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
    for _ in range(20):
        out = eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    lm, ln, gm, gn = eng.kernel_times()
    eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 0)
    print(f"{name}: loss kernel {lm / ln:.4f} ms, gradient kernel {gm / gn:.4f} ms, loss {float(out[0]):.4f}", flush=True)
