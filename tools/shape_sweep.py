#!/usr/bin/env python
"""Single-GPU timing + spot parity of the BASELINE.json per-rank shapes (configs[1..4] at W=1)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sigmoid_loss_b200 import SigmoidLossEngine, _capi

dev = torch.device("cuda", 0)
for (B, D) in [(4096, 768), (8192, 768), (16384, 1024), (32768, 1152)]:
    g = torch.Generator().manual_seed(1234)
    img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    eng = SigmoidLossEngine(B, D, dev)
    tp, bs = torch.tensor([math.log(10.0)], device=dev), torch.tensor([-10.0], device=dev)
    for _ in range(3):
        out = eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    torch.cuda.synchronize()
    eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    lm, ln, gm, gn = eng.kernel_times()
    ms = e0.elapsed_time(e1) / n
    loss, dimg, dtxt, dtp, db = out
    # spot check: Euler identities <dimg,img> = <dtxt,txt> = dt'
    s1 = float((dimg.double() * img.double()).sum()); s2 = float((dtxt.double() * txt.double()).sum())
    print(f"B={B} D={D}: {ms:.3f} ms/step, {B / ms * 1e3 / 1e6:.2f} Mpairs/s, {6.0 * B * B * D / ms / 1e9:.0f} TFLOP/s; "
          f"loss kernel {lm / ln:.3f} ms ({2.0*B*B*D/(lm/ln)/1e9:.0f} TF/s), gradient kernel {gm / gn:.3f} ms "
          f"({4.0*B*B*D/(gm/gn)/1e9:.0f} TF/s); loss={float(loss):.5f} dt'={float(dtp):.5f} <dimg,img>={s1:.5f} <dtxt,txt>={s2:.5f} "
          f"workspace {eng.workspace_bytes / 2**30:.2f} GiB", flush=True)
    eng.close()
    del img, txt, out, dimg, dtxt
    torch.cuda.empty_cache()
