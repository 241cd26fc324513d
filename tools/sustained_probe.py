#!/usr/bin/env python
"""How the step time drifts from the burst to the sustained power state (single GPU, headline shape)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sigmoid_loss_b200 import SigmoidLossEngine

dev = torch.device("cuda", 0)
B, D = 16384, 1024
g = torch.Generator().manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
eng = SigmoidLossEngine(B, D, dev)
tp, bs = torch.tensor([math.log(10.0)], device=dev), torch.tensor([-10.0], device=dev)
for _ in range(3):
    eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
torch.cuda.synchronize()
time.sleep(2.0)   # let the GPU cool / clocks settle to idle
edges = [0, 10, 20, 40, 80, 160, 320, 640, 1280]
evs = [torch.cuda.Event(enable_timing=True) for _ in edges]
k = 0
for i in range(edges[-1] + 1):
    if i == edges[k]:
        evs[k].record()
        k += 1
        if k == len(edges):
            break
    eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
torch.cuda.synchronize()
t = 0.0
for j in range(len(edges) - 1):
    ms = evs[j].elapsed_time(evs[j + 1])
    n = edges[j + 1] - edges[j]
    t += ms
    print(f"steps {edges[j]:5d}-{edges[j+1]:5d}: {ms / n:.4f} ms/step ({6.0*B*B*D/(ms/n)/1e9:.0f} TFLOP/s)  cumulative {t:.0f} ms", flush=True)
