#!/usr/bin/env python
"""On-GPU bring-up diagnostics. Each case runs in its own subprocess (a device-side trap kills the CUDA
context), with a timeout, and everything is logged under gpurun_out/.

    python tools/gpu_diag.py --all            # everything
    python tools/gpu_diag.py --case gemm --cg 1 --amn 0 --bmn 0
    python tools/gpu_diag.py --case loss --cg 2
"""
from __future__ import annotations

import argparse
import ctypes
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gemm_case(cg: int, amn: int, bmn: int) -> int:
    import torch

    from distributed_sigmoid_loss_b200 import _capi

    L = _capi.lib()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    shapes = [(256, 256, 64), (256, 256, 128), (256, 256, 512), (512, 768, 1024), (300, 264, 200), (1024, 1024, 4096),
              (128, 256, 64), (2000, 520, 328)]
    bad = 0
    for mc in (1, 2):
      os.environ["SIGLIP_DEBUG_MCAST"] = str(mc)
      for (M, N, K) in shapes:
          A = torch.randn(M, K, device=dev).to(torch.bfloat16)
          B = torch.randn(N, K, device=dev).to(torch.bfloat16)
          ref = A.float() @ B.float().T
          # storage: K-major [rows][K] or MN-major [K][rows_padded8]
          def store(X, mn):
              if not mn:
                  ld = (X.shape[1] + 7) // 8 * 8
                  buf = torch.zeros(X.shape[0], ld, device=dev, dtype=torch.bfloat16)
                  buf[:, : X.shape[1]] = X
                  return buf, ld
              ld = (X.shape[0] + 7) // 8 * 8
              buf = torch.zeros(X.shape[1], ld, device=dev, dtype=torch.bfloat16)
              buf[:, : X.shape[0]] = X.T
              return buf, ld

          Ab, lda = store(A, amn)
          Bb, ldb = store(B, bmn)
          C = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
          torch.cuda.synchronize()
          rc = L.siglip_debug_gemm(0, cg, M, N, K, Ab.data_ptr(), lda, amn, Bb.data_ptr(), ldb, bmn, C.data_ptr(), N,
                                   torch.cuda.current_stream().cuda_stream)
          if rc != 0:
              print(f"[gemm cg={cg} amn={amn} bmn={bmn}] M={M} N={N} K={K}: rc={rc} {_capi.last_error()}", flush=True)
              return 2
          torch.cuda.synchronize()
          err = (C - ref).abs()
          nan = int(torch.isnan(C).sum())
          scale = float(ref.abs().max())
          mx = float(torch.nan_to_num(err, nan=1e30).max())
          ok = nan == 0 and mx <= 2e-3 * scale + 1e-3
          print(f"[gemm cg={cg} mc={mc} amn={amn} bmn={bmn}] M={M} N={N} K={K}: max_err={mx:.3e} ref_max={scale:.3e} nan={nan} "
                f"{'OK' if ok else 'FAIL'}", flush=True)
          if not ok:
              bad += 1
              # block-level error map (64x64 blocks) to localise layout mistakes
              e = torch.nan_to_num(err, nan=1e3)
              mb, nb = (M + 63) // 64, (N + 63) // 64
              rows = []
              for i in range(min(mb, 8)):
                  rows.append(" ".join(f"{float(e[i*64:(i+1)*64, j*64:(j+1)*64].max()):8.2e}" for j in range(min(nb, 8))))
              print("   block max-err map (64x64):\n   " + "\n   ".join(rows), flush=True)
              print("   C[0,:8]  ", C[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist(), flush=True)
              print("   C[:8,0]  ", C[:8, 0].tolist(), "\n   ref[:8,0]", ref[:8, 0].tolist(), flush=True)
    return 1 if bad else 0


def _rel(a, b):
    import torch

    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def _loss_case(cg: int) -> int:
    import torch

    from distributed_sigmoid_loss_b200 import SigmoidLossEngine
    from oracle.siglip_oracle import torch_reference_fp32

    dev = torch.device("cuda", 0)
    bad = 0
    for (B, D) in [(256, 64), (512, 256), (1000, 136), (4096, 768)]:
        g = torch.Generator().manual_seed(1234)
        img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
        txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
        for (tp, bs) in [(math.log(10.0), -10.0), (math.log(30.0), -3.0)]:
            ref = torch_reference_fp32(img, [txt], tp, bs, 0)
            eng = SigmoidLossEngine(B, D, dev, cta_group=cg)
            tpt = torch.tensor([tp], device=dev, dtype=torch.float32)
            bt = torch.tensor([bs], device=dev, dtype=torch.float32)
            loss, dimg, dtxt, dtp, db = eng.fwd_bwd(img, txt, tpt, bt)
            loss_f = eng.fwd(img, txt, tpt, bt)
            torch.cuda.synchronize()
            r = dict(
                loss=abs(float(loss) - ref["loss"]) / abs(ref["loss"]),
                loss_fwd=abs(float(loss_f) - ref["loss"]) / abs(ref["loss"]),
                dimg=_rel(dimg, ref["dimg"]),
                dtxt=_rel(dtxt, ref["dtxt_chunks"][0]),
                dtp=abs(float(dtp) - ref["dt_prime"]) / (abs(ref["dt_prime"]) + 1e-30),
                db=abs(float(db) - ref["dbias"]) / (abs(ref["dbias"]) + 1e-30),
            )
            ok = all(v < 1e-3 for v in r.values())
            print(f"[loss cg={cg}] B={B} D={D} t'={tp:.3f} b={bs}: " + " ".join(f"{k}={v:.2e}" for k, v in r.items())
                  + f" loss={float(loss):.6f} ref={ref['loss']:.6f} {'OK' if ok else 'FAIL'}", flush=True)
            if not ok:
                bad += 1
            eng.close()
    return 1 if bad else 0


def _time_case(cg: int) -> int:
    """Quick device timing of one step at the headline single-chunk shape."""
    import torch

    from distributed_sigmoid_loss_b200 import SigmoidLossEngine

    dev = torch.device("cuda", 0)
    for (B, D) in [(4096, 768), (16384, 1024)]:
        g = torch.Generator().manual_seed(1234)
        img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
        txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
        eng = SigmoidLossEngine(B, D, dev, cta_group=cg)
        tpt = torch.tensor([math.log(10.0)], device=dev, dtype=torch.float32)
        bt = torch.tensor([-10.0], device=dev, dtype=torch.float32)
        from distributed_sigmoid_loss_b200 import _capi
        for mc in ((2, 1) if cg == 1 else (1,)):
          if cg == 1:
              eng.set_option(_capi.SIGLIP_OPT_MCAST, mc)
          for fn, name, flops in ((lambda: eng.fwd_bwd(img, txt, tpt, bt), f"mc={mc} fwd_bwd", 6.0 * B * B * D),
                                (lambda: eng.fwd(img, txt, tpt, bt), f"mc={mc} fwd", 2.0 * B * B * D)):
              for _ in range(3):
                  fn()
              torch.cuda.synchronize()
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              n = 10
              e0.record()
              for _ in range(n):
                  fn()
              e1.record()
              torch.cuda.synchronize()
              ms = e0.elapsed_time(e1) / n
              print(f"[time cg={cg}] B={B} D={D} {name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s "
                    f"{B / ms * 1e3 / 1e6:.2f} Mpairs/s", flush=True)
        eng.close()
    return 0


def _gemmtime_case(cg: int) -> int:
    """Mainloop-only timing (trivial epilogue) at the three contraction shapes of the headline step."""
    import torch

    from distributed_sigmoid_loss_b200 import _capi

    L = _capi.lib()
    dev = torch.device("cuda", 0)
    B, D = 16384, 1024
    cases = [("S=img@txt^T  (K-major,K-major)", B, B, D, 0, 0),
             ("dimg=G@txt   (K-major,N-major)", B, D, B, 0, 1),
             ("dtxt=G^T@img (M-major,N-major)", B, D, B, 1, 1)]
    stage_opts = (3, 4) if cg == 1 else (4, 5, 6, 7)
    for name, M, N, K, amn, bmn in cases:
        A = torch.randn((K, M) if amn else (M, K), device=dev).to(torch.bfloat16)
        Bm = torch.randn((K, N) if bmn else (N, K), device=dev).to(torch.bfloat16)
        C = torch.empty(M, N, device=dev, dtype=torch.float32)
        for st in stage_opts:
            os.environ["SIGLIP_DEBUG_STAGES"] = str(st)
            ms = ctypes.c_float(0)
            rc = L.siglip_debug_gemm_timed(0, cg, M, N, K, A.data_ptr(), A.shape[1], amn, Bm.data_ptr(), Bm.shape[1],
                                           bmn, C.data_ptr(), N, 10, ctypes.byref(ms),
                                           torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                print(f"[gemmtime cg={cg}] {name}: rc={rc} {_capi.last_error()}", flush=True)
                return 2
            print(f"[gemmtime cg={cg} stages={st}] {name} M={M} N={N} K={K}: {ms.value:.3f} ms "
                  f"{2.0 * M * N * K / ms.value / 1e9:.1f} TFLOP/s", flush=True)
        del A, Bm, C
    # experiments: both operands K-major at the gradient shape; epilogue back-off
    M, N, K = B, D, B
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    Bm = torch.randn(N, K, device=dev).to(torch.bfloat16)
    A2 = torch.randn(K, M, device=dev).to(torch.bfloat16)
    B2 = torch.randn(K, N, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev, dtype=torch.float32)
    os.environ.pop("SIGLIP_DEBUG_STAGES", None)
    os.environ["SIGLIP_DEBUG_WAITSTATS"] = "1"
    for sleep in ((1, 2) if cg == 1 else (1,)):
        os.environ["SIGLIP_DEBUG_MCAST"] = str(sleep)
        for (nm, a, amn, b, bmn) in (("KK", A, 0, Bm, 0), ("KN", A, 0, B2, 1), ("MN", A2, 1, B2, 1)):
            ms = ctypes.c_float(0)
            rc = L.siglip_debug_gemm_timed(0, cg, M, N, K, a.data_ptr(), a.shape[1], amn, b.data_ptr(), b.shape[1],
                                           bmn, C.data_ptr(), N, 20, ctypes.byref(ms),
                                           torch.cuda.current_stream().cuda_stream)
            print(f"[gemmtime cg={cg} mcast={sleep}] {nm} grad shape: {ms.value:.3f} ms "
                  f"{2.0 * M * N * K / ms.value / 1e9:.1f} TFLOP/s rc={rc}", flush=True)
    os.environ.pop("SIGLIP_DEBUG_EPI_SLEEP", None)
    os.environ.pop("SIGLIP_DEBUG_WAITSTATS", None)
    del A, Bm, A2, B2, C
    # which cuBLAS kernel runs on the gradient shape (tile / cluster shape is in the name)
    try:
        from torch.profiler import ProfilerActivity, profile
        a = torch.randn(B, B, device=dev).to(torch.bfloat16)
        b = torch.randn(B, D, device=dev).to(torch.bfloat16)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                a @ b
            torch.cuda.synchronize()
        for ev in prof.key_averages():
            print("[cublas kernel]", ev.key[:150], f"{ev.device_time_total / max(ev.count, 1):.1f} us", flush=True)
        del a, b
    except Exception as ex:  # noqa: BLE001
        print("[cublas kernel] profiler failed:", ex, flush=True)
    # cuBLAS reference points on the same shapes
    for name, M, N, K, _, _ in cases:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(K, N, device=dev).to(torch.bfloat16)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[gemmtime cublas] M={M} N={N} K={K}: {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return 0


def _run_all(args) -> int:
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = open(os.path.join(out_dir, "diag.log"), "a")
    cases = []
    for cg in (1, 2):
        for (amn, bmn) in ((0, 0), (0, 1), (1, 1), (1, 0)):
            cases.append(["--case", "gemm", "--cg", str(cg), "--amn", str(amn), "--bmn", str(bmn)])
    for cg in (1, 2):
        cases.append(["--case", "loss", "--cg", str(cg)])
    for cg in (1, 2):
        cases.append(["--case", "time", "--cg", str(cg)])
    for cg in (1, 2):
        cases.append(["--case", "gemmtime", "--cg", str(cg)])
    if args.only:
        cases = [c for c in cases if c[1] in args.only.split(",")]
    worst = 0
    for c in cases:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__)] + c, capture_output=True, text=True,
                               timeout=args.timeout)
            rc, so, se = p.returncode, p.stdout, p.stderr
        except subprocess.TimeoutExpired as ex:
            rc, so, se = 124, (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else (ex.stdout or ""), "TIMEOUT"
        msg = f"=== {' '.join(c)} rc={rc} ({time.time() - t0:.1f}s)\n{so}"
        if rc != 0:
            msg += f"--- stderr tail:\n{se[-2000:]}\n"
        print(msg, flush=True)
        log.write(msg + "\n")
        log.flush()
        worst = max(worst, rc)
    return worst


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--case", default="")
    ap.add_argument("--cg", type=int, default=1)
    ap.add_argument("--amn", type=int, default=0)
    ap.add_argument("--bmn", type=int, default=0)
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.all:
        return _run_all(args)
    if args.case == "gemm":
        return _gemm_case(args.cg, args.amn, args.bmn)
    if args.case == "loss":
        return _loss_case(args.cg)
    if args.case == "time":
        return _time_case(args.cg)
    if args.case == "gemmtime":
        return _gemmtime_case(args.cg)
    ap.error("need --all or --case")
    return 2


if __name__ == "__main__":
    sys.exit(main())
