"""CPU oracle of the distributed sigmoid (SigLIP) loss hot path — TEST INFRASTRUCTURE, never the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the shipped package (``distributed_sigmoid_loss_b200``) never does.

Parity status: PINNED. The reference has no golden vectors of its own, so the pin is the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified ``/root/reference`` modules
(``DDPSigmoidLoss`` and ``SigLipLoss`` under gloo) in the build container and commits their outputs as
fixtures; ``tests/test_oracle.py`` checks both functions below against every fixture.

Restatements, following the reference line by line (file:line cited per function):
  * ``closed_form``  — float64 numpy, the analytic loss and all four gradients for every rank at once;
  * ``closed_form_uneven`` — the same for ranks with different batch sizes (an extension the reference cannot run;
                       pinned on ``closed_form`` for equal batches and on torch autograd otherwise);
  * ``port_step``    — the same sequence of materialised torch ops the reference executes (GEMM, scale, bias,
                       labels, logsigmoid, sum, autograd), one rank, all W chunks; used as the timed CPU baseline.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np


def _softplus(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x: np.ndarray) -> np.ndarray:
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def closed_form(img_all: np.ndarray, txt_all: np.ndarray, t_prime: float, bias: float, world: int) -> List[Dict]:
    """Per-rank loss and gradients of ``DDPSigmoidLoss.forward`` + ``.backward()`` in float64.

    img_all, txt_all: [world * B, D] global embeddings, rank r owns rows r*B:(r+1)*B
    (test_distributed_sigmoid_loss.py:57-68 slices the same way).

    Follows distributed_sigmoid_loss.py:
      :23      t = exp(t')
      :24      logits = img_r @ txt_c.T * t + bias                  for every chunk c (:41-45)
      :26-30   labels = 2*eye - 1 on the rank's own chunk, -1 elsewhere
      :32-33   loss_c = sum(-logsigmoid(labels * logits))
      :47      total / gpu_batch_size  (LOCAL batch)
    Gradients: autograd of the above; the text gradient on rank c is the SUM over all ranks' losses, which is
    what the backward of dist_nn.all_gather delivers (torch distributed/nn/functional.py:343-354).
    Returns one dict per rank: loss, dimg [B,D], dtxt [B,D], dt_prime, dbias.
    """
    img_all = np.asarray(img_all, dtype=np.float64)
    txt_all = np.asarray(txt_all, dtype=np.float64)
    n, _ = img_all.shape
    assert n % world == 0
    b = n // world
    t = math.exp(t_prime)
    s = img_all @ txt_all.T                      # s[i, j], i global image row, j global text row
    z = t * s + bias
    y = -np.ones_like(z)
    y[np.arange(n), np.arange(n)] = 1.0          # positives: same global index (own chunk diagonal)
    g = -y * _sigmoid(-y * z) / b                # dL_r/dz_ij for i in rank r
    lossmat = _softplus(-y * z) / b
    dimg_all = t * (g @ txt_all)
    dtxt_all = t * (g.T @ img_all)               # column sums over ALL image ranks
    out = []
    for r in range(world):
        rows = slice(r * b, (r + 1) * b)
        out.append(dict(
            loss=float(lossmat[rows].sum()),
            dimg=dimg_all[rows].copy(),
            dtxt=dtxt_all[rows].copy(),
            dt_prime=float(t * (g[rows] * s[rows]).sum()),
            dbias=float(g[rows].sum()),
        ))
    return out


def closed_form_uneven(img_blocks, txt_blocks, t_prime: float, bias: float) -> List[Dict]:
    """`closed_form` for ranks with DIFFERENT batch sizes. The reference cannot express this (its
    labels are gpu_batch_size x gpu_batch_size, distributed_sigmoid_loss.py:26-30, and all_gather needs equal shapes);
    the semantics extend it where it is defined: rank r scores its B_r images against every rank's texts (:41-45), the
    positives are the diagonal of its own chunk (:28), and its loss is divided by ITS batch (:47).
    img_blocks[r]: [B_r, D], txt_blocks[c]: [B_c, D]. Returns per rank: loss, dimg [B_r, D], dt_prime, dbias and
    contrib[c] = that rank's contribution to the text gradient of chunk c ([B_c, D]); the text gradient rank c ends up
    with is the sum over ranks of contrib[c] (what the backward of all_gather delivers)."""
    t = math.exp(t_prime)
    out = []
    for r, img in enumerate(img_blocks):
        img = np.asarray(img, dtype=np.float64)
        br = img.shape[0]
        res = dict(loss=0.0, dimg=np.zeros_like(img), dt_prime=0.0, dbias=0.0, contrib=[])
        for c, txt in enumerate(txt_blocks):
            txt = np.asarray(txt, dtype=np.float64)
            s = img @ txt.T
            z = t * s + bias
            y = -np.ones_like(z)
            if c == r:
                y[np.arange(br), np.arange(br)] = 1.0
            g = -y * _sigmoid(-y * z) / br
            res["loss"] += float(_softplus(-y * z).sum() / br)
            res["dimg"] += t * (g @ txt)
            res["contrib"].append(t * (g.T @ img))
            res["dt_prime"] += float(t * (g * s).sum())
            res["dbias"] += float(g.sum())
        out.append(res)
    return out


def port_step(img, txt_chunks, t_prime, bias, rank: int, backward: bool = True):
    """One rank's forward(+backward) with the reference's own op sequence on torch CPU tensors.

    img: [B, D] leaf tensor; txt_chunks: list of W [B, D] leaf tensors (what all_gather returns,
    distributed_sigmoid_loss.py:35); t_prime, bias: 0-dim leaf tensors. Every B x B intermediate is
    materialised exactly as in distributed_sigmoid_loss.py:22-33 (that is the point of the baseline).
    Returns the loss tensor; gradients land in .grad of the leaves.
    """
    import torch

    bsz = img.shape[0]
    logsig = torch.nn.LogSigmoid()
    total = 0
    for c, txt in enumerate(txt_chunks):                      # :41
        t = t_prime.exp()                                     # :23
        logits = img @ txt.T * t + bias                       # :24
        if c == rank:                                         # :26-30
            labels = 2 * torch.eye(bsz) - torch.ones(bsz)
        else:
            labels = -1 * torch.ones(bsz)
        total = total + (-logsig(labels * logits)).sum()      # :32-33, :45
    total = total / bsz                                       # :47
    if backward:
        total.backward()
    return total


def torch_reference_fp32(img, txt_chunks, t_prime: float, bias: float, rank: int):
    """fp32 autograd evaluation on ANY torch device of the same math (used on the GPU box, where
    /root/reference does not exist, to check the CUDA path at sizes numpy fp64 cannot reach in seconds).
    Returns dict(loss, dimg, dtxt_chunks[list], dt_prime, dbias) — dtxt_chunks[c] is THIS rank's contribution
    to chunk c's text gradient."""
    import torch

    dev = img.device
    img32 = img.detach().float().requires_grad_(True)
    chunks = [x.detach().float().requires_grad_(True) for x in txt_chunks]
    tp = torch.tensor(float(t_prime), device=dev, dtype=torch.float32, requires_grad=True)
    bb = torch.tensor(float(bias), device=dev, dtype=torch.float32, requires_grad=True)
    bsz = img32.shape[0]
    total = torch.zeros((), device=dev, dtype=torch.float32)
    for c, txt in enumerate(chunks):
        logits = img32 @ txt.T * tp.exp() + bb
        if c == rank:
            labels = 2 * torch.eye(bsz, device=dev) - 1
        else:
            labels = -torch.ones(bsz, device=dev)
        total = total + (-torch.nn.functional.logsigmoid(labels * logits)).sum()
    total = total / bsz
    total.backward()
    return dict(loss=float(total.detach()), dimg=img32.grad, dtxt_chunks=[x.grad for x in chunks],
                dt_prime=float(tp.grad), dbias=float(bb.grad))
