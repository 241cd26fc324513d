"""world_size-2 (and 3) gloo tests on CPU of the N>1 path's host-side logic:
  * the oracle port under a REAL process group (dist_nn.all_gather forward, reduce-scatter-equivalent backward)
    reproduces the reference fixtures, rank by rank;
  * the slot protocol the CUDA path uses for the text gradient — every rank writes one fp32 contribution per owner,
    owners sum the W contributions addressed to them — is emulated with all_gather over gloo and gives the same dtxt;
  * the handle bootstrap used by SigmoidLossEngine (all_gather_object of fixed-size blobs, rank order).
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)


def _worker(rank, world, port, case_name, ret):
    import torch.distributed.nn.functional as dist_nn

    from distributed_sigmoid_loss_b200 import chunk_schedule
    from oracle.siglip_oracle import port_step

    _init(rank, world, port)
    c = load_golden(case_name)
    B = c["batch"]
    img = torch.from_numpy(c["img_all"][rank * B:(rank + 1) * B]).clone().requires_grad_(True)
    txt = torch.from_numpy(c["txt_all"][rank * B:(rank + 1) * B]).clone().requires_grad_(True)
    tp = torch.tensor(c["t_prime"], dtype=torch.float64, requires_grad=True)
    bb = torch.tensor(c["bias"], dtype=torch.float32, requires_grad=True)

    # (1) oracle port over a real differentiable all_gather (what distributed_sigmoid_loss.py:35 does)
    chunks = list(dist_nn.all_gather(txt))
    loss = port_step(img, chunks, tp, bb, rank)
    out = dict(loss=float(loss.detach()), dimg=img.grad.numpy().copy(), dtxt=txt.grad.numpy().copy(),
               dt_prime=float(tp.grad), dbias=float(bb.grad))

    # (2) slot protocol: per-owner contributions computed chunk by chunk in schedule order, exchanged, summed
    with torch.no_grad():
        gathered = [torch.empty_like(txt) for _ in range(world)]
        dist.all_gather(gathered, txt.detach())
    slots = [None] * world
    for owner in chunk_schedule(rank, world, bidir=(world > 2)):   # either visiting order covers the same pairs
        a = img.detach().clone().requires_grad_(True)
        t = gathered[owner].clone().requires_grad_(True)
        tp2 = torch.tensor(c["t_prime"], dtype=torch.float64)
        z = a @ t.T * tp2.exp() + c["bias"]
        lab = 2 * torch.eye(B) - 1 if owner == rank else -torch.ones(B)
        (-(torch.nn.functional.logsigmoid(lab * z)).sum() / B).backward()
        slots[owner] = t.grad.float().contiguous()
    mine = torch.stack(slots)                       # [W, B, D]: my contribution to every owner
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)                 # (gloo has no all_to_all; owners pick their slot of each rank)
    out["dtxt_slots"] = torch.stack([e[rank] for e in everyone]).sum(0).numpy()

    # (2b) SIGLIP_OPT_SYNC_SCALAR_GRADS semantics: mean over ranks of (dt', dbias), rank-ordered sum (bit-identical
    # on every rank) — what average_gradients does for the two parameters (test_distributed_sigmoid_loss.py:79-83)
    pair = torch.tensor([out["dt_prime"], out["dbias"]], dtype=torch.float32)
    pairs = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(pairs, pair)
    acc = torch.zeros(2)
    for p_ in pairs:
        acc = acc + p_
    out["scalar_mean"] = (acc / world).numpy()

    # (3) handle bootstrap
    blob = bytes([rank]) * 208
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    out["blobs_ok"] = all(b == bytes([r]) * 208 for r, b in enumerate(blobs))
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name,port", [("w2_b32_d512", 29711), ("w3_b5_d16", 29712)])
def test_multi_rank_host_logic_over_gloo(case_name, port):
    c = load_golden(case_name)
    world = c["world"]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, case_name, ret), nprocs=world, join=True)
    for r in range(world):
        ref = c["variants"]["ddp"][r]
        out = ret[r]
        assert out["blobs_ok"]
        assert abs(out["loss"] - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"])) + 1e-6
        np.testing.assert_allclose(out["dimg"], ref["dimg"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(out["dtxt"], ref["dtxt"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(out["dtxt_slots"], ref["dtxt"], rtol=1e-4, atol=1e-6)
        assert abs(out["dt_prime"] - float(ref["dt_prime"])) <= 1e-5 * abs(float(ref["dt_prime"])) + 1e-7
        assert abs(out["dbias"] - float(ref["dbias"])) <= 1e-5 * abs(float(ref["dbias"])) + 1e-7
        want = np.mean([[float(c["variants"]["ddp"][q]["dt_prime"]), float(c["variants"]["ddp"][q]["dbias"])]
                        for q in range(world)], axis=0)
        np.testing.assert_allclose(out["scalar_mean"], want, rtol=1e-5)
        np.testing.assert_array_equal(out["scalar_mean"], ret[0]["scalar_mean"])
