#!/usr/bin/env python
"""How long does ONE wave of gradient-kernel tiles take? siglip_debug_gemm_timed (the same mainloop, trivial epilogue)
on M x 768 x K problems that are <= 1 tile per SM pair, both operand layouts of the two gradient contractions, with the
in-kernel cycle accounting of the MMA-issuing thread (SIGLIP_DEBUG_WAITSTATS)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SIGLIP_DEBUG_WAITSTATS"] = "1"
os.environ["SIGLIP_DEBUG_AB_F16"] = "1"
import torch

from distributed_sigmoid_loss_b200 import _capi

L = _capi.lib()
dev = torch.device("cuda", 0)
for (M, N, K) in [(4096, 768, 4096), (4096, 768, 1024), (4096, 768, 256), (8192, 768, 8192)]:
    for (amn, name) in [(0, "dimg layout: A K-major"), (1, "dtxt layout: A M-major")]:
        A = torch.randn(M, K, device=dev).to(torch.float16)
        Bm = torch.randn(K, N, device=dev).to(torch.float16)          # N-major: stored [K][N]
        Ab = A if not amn else A.T.contiguous()                         # M-major: stored [K][M]
        C = torch.empty(M, N, device=dev)
        ms = ctypes.c_float(0)
        rc = L.siglip_debug_gemm_timed(0, 2, M, N, K, Ab.data_ptr(), Ab.shape[1], amn, Bm.data_ptr(), N, 1,
                                       C.data_ptr(), N, 50, ctypes.byref(ms), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _capi.last_error()
        tiles = ((M + 255) // 256) * ((N + 255) // 256)
        print(f"M={M} N={N} K={K} {name}: {ms.value * 1e3:.1f} us per launch, {tiles} tiles, "
              f"{2.0 * M * N * K / ms.value / 1e9:.0f} TFLOP/s", flush=True)
