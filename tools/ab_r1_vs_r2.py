#!/usr/bin/env python
"""Interleaved A/B of the round-1 library against the current one. Build the old library first (it is not kept in the
tree):   for f in ptx.cuh siglip_kernels.cu siglip_kernels.cuh siglip_capi.cu; do git show b304770:distributed_sigmoid_loss_b200/csrc/$f > /tmp/r1/distributed_sigmoid_loss_b200/csrc/$f; done;
         git show b304770:include/siglip_b200.h > /tmp/r1/include/siglip_b200.h;
         (cd /tmp/r1/distributed_sigmoid_loss_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared
          -o <repo>/tools/_r1_libsiglip_b200.so siglip_kernels.cu siglip_capi.cu)
Interleaved A/B of that library (tools/_r1_libsiglip_b200.so, built from git b304770) against the current one
on the same GPU in the same process: siglip_fwd_bwd at one shape, alternating ~0.25 s blocks in the sustained
(power-capped) state, kernel times from the libraries' own CUDA-event brackets (SIGLIP_OPT_KERNEL_TIMING)."""
import argparse
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16384)
ap.add_argument("--D", type=int, default=1024)
ap.add_argument("--rounds", type=int, default=10)
ap.add_argument("--block-ms", type=float, default=250.0)
ap.add_argument("--r2-lib", default=os.path.join(ROOT, "distributed_sigmoid_loss_b200", "libsiglip_b200.so"))
ap.add_argument("--r1-lib", default=os.path.join(ROOT, "tools", "_r1_libsiglip_b200.so"),
                help="the library in the 'r1' slot (any earlier build of the same C ABI, e.g. one made from `git show HEAD:...`)")
ap.add_argument("--fwd-only", action="store_true", help="loss kernel only (siglip_fwd): no gradient kernel in between")
ap.add_argument("--only", default="", help="r1 or r2: run a few steps of one library only (ncu target)")
ap.add_argument("--r2-opts", default="", help="comma-separated option=value pairs set on the current library only")
ap.add_argument("--r1-env", default="", help="comma-separated NAME=VALUE pairs in the environment while the 'r1' context is created")
ap.add_argument("--r2-env", default="", help="... while the 'r2' context is created (the library reads its SIGLIP_* switches then)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = torch.Generator().manual_seed(1234)
img = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
txt = torch.nn.functional.normalize(torch.randn(a.B, a.D, generator=g)).to(torch.bfloat16).to(dev)
sc = torch.tensor([math.log(10.0), -10.0, 0, 0, 0], device=dev)
dimg = torch.empty(a.B, a.D, device=dev, dtype=torch.bfloat16)
dtxt = torch.empty(a.B, a.D, device=dev, dtype=torch.bfloat16)
vp = ctypes.c_void_p


def load(path, env=""):
    pairs = [kv.split("=") for kv in filter(None, env.split(","))]
    for k, v in pairs:
        os.environ[k] = v
    try:
        return _load(path)
    finally:
        for k, _ in pairs:
            os.environ.pop(k, None)


def _load(path):
    L = ctypes.CDLL(path)
    L.siglip_ctx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.siglip_ctx_set_option.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.siglip_fwd_bwd.argtypes = [vp] * 11
    L.siglip_fwd.argtypes = [vp] * 7
    L.siglip_ctx_kernel_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    L.siglip_last_error.restype = ctypes.c_char_p
    h = vp()
    assert L.siglip_ctx_create(ctypes.byref(h), 0, 0, 1, a.B, a.D) == 0, L.siglip_last_error()
    assert L.siglip_ctx_set_option(h, 7, 1) == 0          # SIGLIP_OPT_GRAD_BF16
    return L, h


def apply_opts(L, h):
    for kv in filter(None, a.r2_opts.split(",")):
        k, v = kv.split("=")
        assert L.siglip_ctx_set_option(h, int(k), int(v)) == 0, L.siglip_last_error()


libs = {"r1": load(a.r1_lib, a.r1_env),
        "r2": load(a.r2_lib, a.r2_env)}
apply_opts(*libs["r2"])
st = torch.cuda.current_stream().cuda_stream
p = sc.data_ptr()


def step(L, h):
    if a.fwd_only:
        rc = L.siglip_fwd(h, img.data_ptr(), txt.data_ptr(), p, p + 4, p + 8, st)
        assert rc == 0, L.siglip_last_error()
        return
    rc = L.siglip_fwd_bwd(h, img.data_ptr(), txt.data_ptr(), p, p + 4, p + 8, dimg.data_ptr(), dtxt.data_ptr(), p + 12,
                          p + 16, st)
    assert rc == 0, L.siglip_last_error()


def block(name, timing):
    L, h = libs[name]
    L.siglip_ctx_set_option(h, 3, 1 if timing else 0)      # SIGLIP_OPT_KERNEL_TIMING
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(4, int(a.block_ms / est[name]))
    e0.record()
    for _ in range(n):
        step(L, h)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lk = gk = None
    if timing:
        lm, gm = ctypes.c_double(), ctypes.c_double()
        ln, gn = ctypes.c_int(), ctypes.c_int()
        L.siglip_ctx_kernel_times(h, ctypes.byref(lm), ctypes.byref(ln), ctypes.byref(gm), ctypes.byref(gn))
        lk, gk = lm.value / max(ln.value, 1), gm.value / max(gn.value, 1)
    return ms, lk, gk


if a.only:
    L, h = libs[a.only]
    for _ in range(6):
        step(L, h)
    torch.cuda.synchronize()
    print("ran 6 steps of", a.only)
    sys.exit(0)

est = {"r1": 1.3 * (a.B / 16384) ** 2 * a.D / 1024, "r2": 1.3 * (a.B / 16384) ** 2 * a.D / 1024}
for name in libs:          # warm-up into the sustained state
    for _ in range(3):
        est[name] = block(name, False)[0]
res = {n: {"step": [], "step_ev": [], "loss": [], "grad": []} for n in libs}
for r in range(a.rounds):
    for name in (("r1", "r2") if r % 2 == 0 else ("r2", "r1")):
        res[name]["step"].append(block(name, False)[0])
        ms, lk, gk = block(name, True)
        res[name]["step_ev"].append(ms)
        res[name]["loss"].append(lk)
        res[name]["grad"].append(gk)


def med(x):
    x = sorted(x)
    return x[len(x) // 2]


print(f"B={a.B} D={a.D}, {a.rounds} interleaved rounds of ~{a.block_ms:.0f} ms blocks, medians (ms):")
for name in libs:
    r = res[name]
    print(f"  {name}: step {med(r['step']):.4f} (with per-kernel events {med(r['step_ev']):.4f})  loss kernel {med(r['loss']):.4f}  "
          f"gradient kernel {med(r['grad']):.4f}")
print(f"  r2 / r1 step time: {med(res['r2']['step']) / med(res['r1']['step']):.4f}")
