#!/usr/bin/env python
"""A/B timing of the loss kernel's training-only extras (sigma store, fp16 conversions) — timing experiment only."""
import math, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from distributed_sigmoid_loss_b200 import SigmoidLossEngine, _capi
    dev = torch.device("cuda", 0)
    B, D = 16384, 1024
    g = torch.Generator().manual_seed(1234)
    img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    eng = SigmoidLossEngine(B, D, dev, cta_group=2)
    tp, bs = torch.tensor([math.log(10.0)], device=dev), torch.tensor([-10.0], device=dev)
    for _ in range(5):
        eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    torch.cuda.synchronize()
    eng.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
    for _ in range(20):
        eng.fwd_bwd(img, txt, tp, bs, torch.bfloat16)
    lm, ln, gm, gn = eng.kernel_times()
    print(f"{os.environ.get('LABEL')}: loss kernel {lm/ln:.4f} ms, gradient kernel {gm/gn:.4f} ms", flush=True)
else:
    for label, env in (("full", {}), ("no cvt", {"SIGLIP_DEBUG_NO_CVT": "1"}), ("no gstore", {"SIGLIP_DEBUG_NO_GSTORE": "1"}),
                       ("no cvt, no gstore", {"SIGLIP_DEBUG_NO_CVT": "1", "SIGLIP_DEBUG_NO_GSTORE": "1"}), ("full again", {})):
        e = dict(os.environ); e.update(env); e["LABEL"] = label
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e)
