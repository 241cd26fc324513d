#!/usr/bin/env python
"""Power experiment: the same tcgen05 mainloop on the same randn data held as bf16 vs as fp16 operands, burst and
sustained (power-capped) state. Answers whether kind::f16 with fp16 inputs costs more energy per flop than bf16."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sigmoid_loss_b200 import _capi

L = _capi.lib()
dev = torch.device("cuda", 0)
M, N, K = 18944, 1024, 16384
A32 = torch.randn(M, K, device=dev)
B32 = torch.randn(K, N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.float32)
flops = 2.0 * M * N * K
ops = {
    "bf16": (A32.to(torch.bfloat16), B32.to(torch.bfloat16)),
    "fp16": (A32.to(torch.float16), B32.to(torch.float16)),
    # what the gradient kernel actually multiplies: sigma-like A (tiny positive, x 2^14), embedding-like B (x 16)
    "fp16 sigma-like": ((torch.rand(M, K, device=dev) * 3e-4 * 16384).to(torch.float16),
                        (torch.nn.functional.normalize(torch.randn(K, N, device=dev)) * 16).to(torch.float16)),
    "bf16 sigma-like": ((torch.rand(M, K, device=dev) * 3e-4).to(torch.bfloat16),
                        torch.nn.functional.normalize(torch.randn(K, N, device=dev)).to(torch.bfloat16)),
}
del A32, B32


def run(kind, iters):
    a, b = ops[kind]
    if kind.startswith("fp16"):
        os.environ["SIGLIP_DEBUG_AB_F16"] = "1"
    else:
        os.environ.pop("SIGLIP_DEBUG_AB_F16", None)
    ms = ctypes.c_float(0)
    rc = L.siglip_debug_gemm_timed(0, 2, M, N, K, a.data_ptr(), K, 0, b.data_ptr(), N, 1, C.data_ptr(), N, iters,
                                   ctypes.byref(ms), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _capi.last_error()
    return ms.value


for rnd in range(2):
    for kind in ops:
        time.sleep(1.0)
        burst = run(kind, 20)
        sustained = run(kind, 1500)
        print(f"{kind:16s}: burst {burst:.4f} ms {flops / burst / 1e9:.0f} TFLOP/s | sustained {sustained:.4f} ms "
              f"{flops / sustained / 1e9:.0f} TFLOP/s", flush=True)
