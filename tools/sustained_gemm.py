#!/usr/bin/env python
"""Sustained-state comparison of this repo's tcgen05 mainloop (trivial epilogue) with cuBLAS on a shape that fills the
74 SM pairs exactly (no wave quantisation): M=18944 (74*256), N=1024, K=16384, random bf16 operands."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sigmoid_loss_b200 import _capi

L = _capi.lib()
dev = torch.device("cuda", 0)
M, N, K = 18944, 1024, 16384
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
Bm = torch.randn(N, K, device=dev).to(torch.bfloat16)
Bt = Bm.T.contiguous()            # [K, N] for the N-major variant and for cuBLAS
C = torch.empty(M, N, device=dev, dtype=torch.float32)
flops = 2.0 * M * N * K


def ours(iters, bmn):
    ms = ctypes.c_float(0)
    b = Bt if bmn else Bm
    rc = L.siglip_debug_gemm_timed(0, 2, M, N, K, A.data_ptr(), K, 0, b.data_ptr(), b.shape[1], bmn, C.data_ptr(), N,
                                   iters, ctypes.byref(ms), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _capi.last_error()
    return ms.value


def cublas(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    torch.matmul(A, Bt, out=out)
    e0.record()
    for _ in range(iters):
        torch.matmul(A, Bt, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ours_mc(n, bmn, mc):
    os.environ["SIGLIP_DEBUG_MCAST"] = str(mc)
    return ours(n, bmn)


for rnd in range(2):
    for name, fn in (("ours K-major B      ", lambda n: ours_mc(n, 0, 1)), ("ours N-major B      ", lambda n: ours_mc(n, 1, 1)),
                     ("ours K-major B 2x2mc", lambda n: ours_mc(n, 0, 2)), ("ours N-major B 2x2mc", lambda n: ours_mc(n, 1, 2)),
                     ("cuBLAS bf16         ", cublas)):
        time.sleep(1.0)
        burst = fn(20)
        sustained = fn(1500)
        print(f"{name}: burst(20 iters) {burst:.4f} ms {flops / burst / 1e9:.0f} TFLOP/s | sustained(1500 iters) "
              f"{sustained:.4f} ms {flops / sustained / 1e9:.0f} TFLOP/s", flush=True)
