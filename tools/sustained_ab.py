#!/usr/bin/env python
"""A/B of tuning options in the SUSTAINED power state: alternate configurations in blocks of steps, long enough to be
power-capped, and compare per-kernel times (same process, same GPU, interleaved to cancel drift)."""
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
tp, bs = torch.tensor([math.log(10.0)], device=dev), torch.tensor([-10.0], device=dev)
configs = [("base", {}),
           ("grad sleep 500", {_capi.SIGLIP_OPT_EPI_SLEEP_GRAD_NS: 500}),
           ("grad sleep 2000", {_capi.SIGLIP_OPT_EPI_SLEEP_GRAD_NS: 2000}),
           ("grad sleep 8000", {_capi.SIGLIP_OPT_EPI_SLEEP_GRAD_NS: 8000}),
           ("loss sleep 200", {_capi.SIGLIP_OPT_EPI_SLEEP_LOSS_NS: 200}),
           ("grad stages 4", {_capi.SIGLIP_OPT_STAGES_GRAD: 4}),
           ("loss stages 4", {_capi.SIGLIP_OPT_STAGES_LOSS: 4}),
           ]
defaults = {_capi.SIGLIP_OPT_EPI_SLEEP_GRAD_NS: 0, _capi.SIGLIP_OPT_EPI_SLEEP_LOSS_NS: 0, _capi.SIGLIP_OPT_STAGES_GRAD: 0,
            _capi.SIGLIP_OPT_STAGES_LOSS: 0}
for _ in range(500):   # reach the sustained state
    eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
torch.cuda.synchronize()
acc = {name: [0.0, 0, 0.0, 0] for name, _ in configs}
for rnd in range(4):
    for name, opts in configs:
        for k, v in defaults.items():
            eng.set_option(k, v)
        for k, v in opts.items():
            eng.set_option(k, v)
        for _ in range(5):
            eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
        eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
        for _ in range(40):
            eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
        lm, ln, gm, gn = eng.kernel_times()
        eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 0)
        a = acc[name]
        a[0] += lm; a[1] += ln; a[2] += gm; a[3] += gn
for name, _ in configs:
    a = acc[name]
    print(f"{name:18s}: loss kernel {a[0] / a[1]:.4f} ms, gradient kernel {a[2] / a[3]:.4f} ms, sum {a[0]/a[1] + a[2]/a[3]:.4f}", flush=True)
