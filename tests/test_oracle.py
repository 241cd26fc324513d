"""The oracle (oracle/siglip_oracle.py) against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py). CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle.siglip_oracle import closed_form, port_step, torch_reference_fp32

# fp32 reference vs fp64 closed form: fp32 summation noise only (SURVEY.md §8c measured <= 2.6e-7 abs on grads)
REL = 2e-5
ABS = 2e-6


def _close(a, b, rel=REL, abs_=ABS):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= abs_ + rel * np.abs(b))


@pytest.mark.parametrize("name", golden_cases("all"))
@pytest.mark.parametrize("variant", ["ddp", "rw_bidir", "rw_uni"])
def test_closed_form_matches_reference(name, variant):
    c = load_golden(name)
    out = closed_form(c["img_all"], c["txt_all"], c["t_prime"], c["bias"], c["world"])
    for r in range(c["world"]):
        ref = c["variants"][variant][r]
        assert abs(out[r]["loss"] - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"])) + 1e-6
        assert _close(out[r]["dimg"], ref["dimg"]), f"dimg rank {r}"
        assert _close(out[r]["dtxt"], ref["dtxt"]), f"dtxt rank {r}"
        assert abs(out[r]["dt_prime"] - float(ref["dt_prime"])) <= 2e-5 * abs(float(ref["dt_prime"])) + 1e-6
        assert abs(out[r]["dbias"] - float(ref["dbias"])) <= 2e-5 * abs(float(ref["dbias"])) + 1e-6


@pytest.mark.parametrize("name", golden_cases("all"))
def test_port_step_matches_reference(name):
    """The op-for-op torch port (the timed CPU baseline) equals the reference module rank by rank. The text
    gradient of the port is per chunk; summing the chunk gradients over ranks reproduces all_gather's backward."""
    c = load_golden(name)
    W, B = c["world"], c["batch"]
    img_all = torch.from_numpy(c["img_all"])
    txt_all = torch.from_numpy(c["txt_all"])
    dtxt_sum = torch.zeros_like(txt_all)
    for r in range(W):
        img = img_all[r * B:(r + 1) * B].clone().requires_grad_(True)
        chunks = [txt_all[k * B:(k + 1) * B].clone().requires_grad_(True) for k in range(W)]
        tp = torch.tensor(c["t_prime"], dtype=torch.float64, requires_grad=True)
        bb = torch.tensor(c["bias"], dtype=torch.float32, requires_grad=True)
        loss = port_step(img, chunks, tp, bb, r)
        ref = c["variants"]["ddp"][r]
        assert abs(float(loss) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"])) + 1e-6
        assert _close(img.grad.numpy(), ref["dimg"], rel=1e-5, abs_=1e-7)
        assert abs(float(tp.grad) - float(ref["dt_prime"])) <= 1e-5 * abs(float(ref["dt_prime"])) + 1e-7
        assert abs(float(bb.grad) - float(ref["dbias"])) <= 1e-5 * abs(float(ref["dbias"])) + 1e-7
        for k in range(W):
            dtxt_sum[k * B:(k + 1) * B] += chunks[k].grad
    for r in range(W):
        assert _close(dtxt_sum[r * B:(r + 1) * B].numpy(), c["variants"]["ddp"][r]["dtxt"], rel=1e-5, abs_=1e-7)


@pytest.mark.parametrize("name", ["w3_b5_d16", "w2_b24_d40_warm"])
def test_torch_reference_fp32_matches_reference(name):
    """The helper the GPU tests use at large sizes is the same math (checked here on CPU)."""
    c = load_golden(name)
    W, B = c["world"], c["batch"]
    img_all = torch.from_numpy(c["img_all"])
    txt_all = torch.from_numpy(c["txt_all"])
    for r in range(W):
        chunks = [txt_all[k * B:(k + 1) * B] for k in range(W)]
        out = torch_reference_fp32(img_all[r * B:(r + 1) * B], chunks, c["t_prime"], c["bias"], r)
        ref = c["variants"]["ddp"][r]
        assert abs(out["loss"] - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
        assert _close(out["dimg"].numpy(), ref["dimg"], rel=1e-4, abs_=1e-6)


def test_variants_agree_with_each_other():
    """The reference's own invariant (test_sigmoid_loss_variants.py:112-113): all-gather variant == ring variant."""
    for name in golden_cases("all"):
        c = load_golden(name)
        for r in range(c["world"]):
            a, b = c["variants"]["ddp"][r], c["variants"]["rw_bidir"][r]
            assert _close(a["dimg"], b["dimg"], rel=1e-3, abs_=1e-6)
            assert _close(a["dtxt"], b["dtxt"], rel=1e-3, abs_=1e-6)


def test_world_invariance():
    """The reference's other invariant (test_distributed_sigmoid_loss.py:140-141): the W-rank mean of the per-rank
    objectives has the same image/text gradients as one rank holding the whole batch, once scaled by 1/W
    (DDP's gradient averaging)."""
    c = load_golden("w4_b8_d64")
    W = c["world"]
    multi = closed_form(c["img_all"], c["txt_all"], c["t_prime"], c["bias"], W)
    single = closed_form(c["img_all"], c["txt_all"], c["t_prime"], c["bias"], 1)[0]
    dimg = np.concatenate([m["dimg"] for m in multi]) / W
    dtxt = np.concatenate([m["dtxt"] for m in multi]) / W
    assert np.allclose(dimg, single["dimg"], rtol=1e-9, atol=1e-12)
    assert np.allclose(dtxt, single["dtxt"], rtol=1e-9, atol=1e-12)


def test_closed_form_equals_port_on_random_problems():
    """Property check beyond the fixtures: for random (W, B, D, t', bias) — cold, warm and saturated logits — the fp64
    closed form and the reference's op sequence (fp64 inputs through torch autograd) agree to 1e-9."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(W=st.integers(1, 4), B=st.integers(1, 9), D=st.integers(1, 12), tp=st.floats(-1.0, 4.5),
           bias=st.floats(-15.0, 3.0), seed=st.integers(0, 10_000))
    def check(W, B, D, tp, bias, seed):
        g = torch.Generator().manual_seed(seed)
        img_all = torch.nn.functional.normalize(torch.randn(W * B, D, generator=g, dtype=torch.float64), dim=-1)
        txt_all = torch.nn.functional.normalize(torch.randn(W * B, D, generator=g, dtype=torch.float64), dim=-1)
        out = closed_form(img_all.numpy(), txt_all.numpy(), tp, bias, W)
        dtxt_sum = torch.zeros_like(txt_all)
        for r in range(W):
            img = img_all[r * B:(r + 1) * B].clone().requires_grad_(True)
            chunks = [txt_all[k * B:(k + 1) * B].clone().requires_grad_(True) for k in range(W)]
            t = torch.tensor(tp, dtype=torch.float64, requires_grad=True)
            b = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
            loss = port_step(img, chunks, t, b, r)
            assert abs(float(loss) - out[r]["loss"]) <= 1e-9 * max(1.0, abs(out[r]["loss"]))
            assert _close(img.grad.numpy(), out[r]["dimg"], rel=1e-9, abs_=1e-12)
            assert abs(float(t.grad) - out[r]["dt_prime"]) <= 1e-9 * max(1.0, abs(out[r]["dt_prime"]))
            assert abs(float(b.grad) - out[r]["dbias"]) <= 1e-9 * max(1.0, abs(out[r]["dbias"]))
            for k in range(W):
                dtxt_sum[k * B:(k + 1) * B] += chunks[k].grad
        for r in range(W):
            assert _close(dtxt_sum[r * B:(r + 1) * B].numpy(), out[r]["dtxt"], rel=1e-9, abs_=1e-12)

    check()


def test_uneven_closed_form_reduces_to_the_pinned_one_and_matches_autograd():
    """closed_form_uneven (used by the GPU tests of siglip_ctx_create_uneven) is pinned twice: for equal batches it must
    reproduce closed_form — itself pinned on the reference's fixtures — and for unequal ones torch autograd of the
    reference's op sequence with per-rank labels."""
    import torch

    from oracle.siglip_oracle import closed_form, closed_form_uneven

    rng = np.random.default_rng(5)
    W, B, D = 3, 7, 12
    img = rng.standard_normal((W * B, D))
    txt = rng.standard_normal((W * B, D))
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    tp, bias = math.log(9.0), -6.0
    eq = closed_form(img, txt, tp, bias, W)
    un = closed_form_uneven([img[r * B:(r + 1) * B] for r in range(W)], [txt[c * B:(c + 1) * B] for c in range(W)], tp, bias)
    for r in range(W):
        assert abs(un[r]["loss"] - eq[r]["loss"]) < 1e-12 and np.allclose(un[r]["dimg"], eq[r]["dimg"], atol=1e-14)
        assert abs(un[r]["dt_prime"] - eq[r]["dt_prime"]) < 1e-12 and abs(un[r]["dbias"] - eq[r]["dbias"]) < 1e-12
        assert np.allclose(sum(un[q]["contrib"][r] for q in range(W)), eq[r]["dtxt"], atol=1e-14)
    Bs = (5, 9, 3)
    imgs = [rng.standard_normal((b, D)) for b in Bs]
    txts = [rng.standard_normal((b, D)) for b in Bs]
    un = closed_form_uneven(imgs, txts, tp, bias)
    for r, b in enumerate(Bs):
        a = torch.tensor(imgs[r], requires_grad=True)
        ts = [torch.tensor(x, requires_grad=True) for x in txts]
        t_ = torch.tensor(tp, dtype=torch.float64, requires_grad=True)
        b_ = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
        total = 0
        for c, tx in enumerate(ts):
            logits = a @ tx.T * t_.exp() + b_
            labels = 2 * torch.eye(b, dtype=torch.float64) - 1 if c == r else -torch.ones(Bs[c], dtype=torch.float64)
            total = total + (-torch.nn.functional.logsigmoid(labels * logits)).sum()
        (total / b).backward()
        assert abs(float(total / b) - un[r]["loss"]) < 1e-10
        assert np.allclose(a.grad.numpy(), un[r]["dimg"], atol=1e-12)
        assert abs(float(t_.grad) - un[r]["dt_prime"]) < 1e-10 and abs(float(b_.grad) - un[r]["dbias"]) < 1e-10
        for c in range(W):
            assert np.allclose(ts[c].grad.numpy(), un[r]["contrib"][c], atol=1e-12)
