#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it spawns W gloo/CPU processes (exactly how the reference's own tests run it,
test_distributed_sigmoid_loss.py:35-50,122-130), feeds each rank its slice of seeded, L2-normalised global
embeddings (seeds 42 / 40 as in test_distributed_sigmoid_loss.py:57-68) as leaf tensors, calls

  * ``DDPSigmoidLoss(gpu_batch_size)(img, txt)``                 (distributed_sigmoid_loss.py:8-48)
  * ``SigLipLoss(rank, world_size, bidir)(img, txt, scale, bias)`` (rwightman_sigmoid_loss.py:12-124)

then ``.backward()``, and stores loss / dimg / dtxt / dt_prime / dbias of every rank in ``<case>.npz``.
The reference has no golden vectors of its own; these files are the pin for oracle/ and for the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# (name, world, B per rank, D, t_prime, bias)
CASES = [
    ("w1_b16_d32", 1, 16, 32, float(np.log(10)), -10.0),
    ("w2_b32_d512", 2, 32, 512, float(np.log(10)), -10.0),   # BASELINE.json configs[0]
    ("w3_b5_d16", 3, 5, 16, float(np.log(10)), -10.0),
    ("w4_b8_d64", 4, 8, 64, float(np.log(10)), -10.0),
    ("w5_b4_d32", 5, 4, 32, float(np.log(10)), -10.0),
    ("w2_b24_d40_warm", 2, 24, 40, float(np.log(25.0)), -4.5),  # logits near 0: exercises both sigmoid branches
    ("w1_b300_d136", 1, 300, 136, float(np.log(10)), -10.0),   # ragged vs the 128/256 tiles of the CUDA path
    # raw fp32 inputs (NOT bf16-representable), the way the reference's own test feeds them
    # (test_distributed_sigmoid_loss.py:57-68, 99-101): the fp32-input path of the CUDA module (fp16 x 16 operands)
    ("w2_b32_d512_f32", 2, 32, 512, float(np.log(10)), -10.0),  # BASELINE.json configs[0], raw fp32
    ("w1_b300_d136_f32", 1, 300, 136, float(np.log(10)), -10.0),
    ("w3_b40_d64_f32_warm", 3, 40, 64, float(np.log(25.0)), -4.5),
    # BASELINE.json configs[4] embed dim (D = 1152 = 4.5 column tiles of 256): the "parity sweep vs
    # rwightman_sigmoid_loss.py" — DDPSigmoidLoss and SigLipLoss (uni- and bidirectional ring) of the reference, W = 4, 8
    ("w4_b40_d1152", 4, 40, 1152, float(np.log(10)), -10.0),
    ("w8_b12_d1152_warm", 8, 12, 1152, float(np.log(25.0)), -4.5),
]


def is_raw_f32(name: str) -> bool:
    return "_f32" in name


def global_inputs(world: int, b: int, d: int, raw_f32: bool = False):
    torch.manual_seed(42)
    img = torch.randn(world * b, d)
    torch.manual_seed(40)
    txt = torch.randn(world * b, d)
    # L2-normalise (test_distributed_sigmoid_loss.py:99-101), then round to bf16-representable values and hand them to
    # the reference as fp32: "the reference in fp32 on the same bf16-rounded inputs" is the parity yardstick
    #  — the CUDA path consumes exactly these values as bf16.
    if raw_f32:
        return F.normalize(img), F.normalize(txt)
    return (F.normalize(img).to(torch.bfloat16).float(), F.normalize(txt).to(torch.bfloat16).float())


def worker(rank: int, world: int, b: int, d: int, t_prime: float, bias: float, port: int, raw_f32: bool, ret):
    sys.path.insert(0, REF)
    from distributed_sigmoid_loss import DDPSigmoidLoss
    from rwightman_sigmoid_loss import SigLipLoss

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    img_all, txt_all = global_inputs(world, b, d, raw_f32)
    sl = slice(rank * b, (rank + 1) * b)
    out = {}

    # --- product class ---
    img = img_all[sl].clone().requires_grad_(True)
    txt = txt_all[sl].clone().requires_grad_(True)
    mod = DDPSigmoidLoss(b)
    with torch.no_grad():
        mod.t_prime.fill_(t_prime)
        mod.bias.fill_(bias)
    loss = mod(img, txt)
    loss.backward()
    out["ddp"] = dict(loss=float(loss.detach()), dimg=img.grad.numpy().copy(), dtxt=txt.grad.numpy().copy(),
                      dt_prime=float(mod.t_prime.grad), dbias=float(mod.bias.grad))

    # --- vendored open_clip variant, both ring flavours ---
    for bidir in (True, False):
        img = img_all[sl].clone().requires_grad_(True)
        txt = txt_all[sl].clone().requires_grad_(True)
        scale = torch.nn.Parameter(torch.ones([]) * t_prime)
        lbias = torch.nn.Parameter(torch.ones([]) * bias)
        loss = SigLipLoss(rank=rank, world_size=world, bidir=bidir)(img, txt, scale, lbias)
        loss.backward()
        out["rw_bidir" if bidir else "rw_uni"] = dict(
            loss=float(loss.detach()), dimg=img.grad.numpy().copy(), dtxt=txt.grad.numpy().copy(),
            dt_prime=float(scale.grad), dbias=float(lbias.grad))
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def main():
    port = 29610
    only = set(sys.argv[1:])      # optional: names of the cases to (re)generate
    for (name, world, b, d, t_prime, bias) in CASES:
        if only and name not in only:
            continue
        mgr = mp.Manager()
        ret = mgr.dict()
        raw = is_raw_f32(name)
        mp.spawn(worker, args=(world, b, d, t_prime, bias, port, raw, ret), nprocs=world, join=True)
        port += 1
        img_all, txt_all = global_inputs(world, b, d, raw)
        arrays = dict(img_all=img_all.numpy(), txt_all=txt_all.numpy(), world=np.int64(world), batch=np.int64(b),
                      dim=np.int64(d), t_prime=np.float64(t_prime), bias=np.float64(bias))
        for r in range(world):
            for variant, res in ret[r].items():
                for k, v in res.items():
                    arrays[f"{variant}.r{r}.{k}"] = np.asarray(v)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, "->", path, os.path.getsize(path), "bytes",
              " ".join(f"r{r}:{ret[r]['ddp']['loss']:.6f}" for r in range(world)))


if __name__ == "__main__":
    main()
