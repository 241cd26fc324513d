#!/usr/bin/env python
"""Minimal ncu target: a few fwd_bwd steps of the single-chunk headline shape, nothing else on the GPU."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributed_sigmoid_loss_b200 import SigmoidLossEngine

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16384)
ap.add_argument("--D", type=int, default=1024)
ap.add_argument("--cg", type=int, default=2)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--loopback", type=int, default=0, help="W > 1: rank 0 of a W-rank job against local 'peers'")
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
txt = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
if a.loopback > 1:
    eng = SigmoidLossEngine(a.B, a.D, dev, cta_group=a.cg, rank_world=(0, a.loopback), loopback=True)
    for k in range(a.loopback):
        eng.debug_set_text_chunk(k, txt)
else:
    eng = SigmoidLossEngine(a.B, a.D, dev, cta_group=a.cg)
tp = torch.tensor([math.log(10.0)], device=dev)
b = torch.tensor([-10.0], device=dev)
for _ in range(a.iters):
    if a.fwd_only:
        eng.fwd(img, txt, tp, b)
    else:
        eng.fwd_bwd(img, txt, tp, b)
torch.cuda.synchronize()
print("done", eng.launch_count)
