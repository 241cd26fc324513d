#!/usr/bin/env python
"""Where does the weak-scaling loss come from? One rank per GPU, all phases in the power-capped sustained state:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        tools/scaling_diag.py [--batch 16384 --dim 1024]

  A  uncoupled, W=1   : every GPU runs the single-chunk step at the same time, no cross-rank dependency
                        -> the spread between the GPUs of the box under identical load (silicon / power cap)
  B  uncoupled, W=N   : every GPU runs one rank of an N-rank job against LOCAL "peers" (loopback context)
                        -> cost of the W>1 data flow (fp32 slots, dimg accumulation, fold) without any waiting
  C  coupled, W=N     : the real job (in-kernel NVSwitch pulls / folds, peer flags)
                        -> C - B on a rank = time spent waiting for peers inside its kernels
Every phase prints, per rank: ms/step, loss-kernel and gradient-kernel ms per launch (CUDA events on the stream).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--sustain-ms", type=float, default=700.0)
    ap.add_argument("--phases", default="ABC")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from distributed_sigmoid_loss_b200 import SigmoidLossEngine, _capi

    B, D = args.batch, args.dim
    g = torch.Generator().manual_seed(1234 + rank)
    img = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(B, D, generator=g)).to(torch.bfloat16).to(dev)
    tp, bt = torch.tensor([math.log(10.0)], device=dev), torch.tensor([-10.0], device=dev)

    def nvml_clock():
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), \
                float(pynvml.nvmlDeviceGetPowerUsage(h)) / 1000.0
        except Exception:  # noqa: BLE001
            return None, None

    results = {}

    def run(engine, label, coupled):
        def step():
            engine.fwd_bwd(img, txt, tp, bt, torch.bfloat16)
        for _ in range(3):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        # sustained state: keep stepping until sustain_ms of GPU time have passed (collective loop when coupled)
        spent = 0.0
        while spent < args.sustain_ms:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(8):
                step()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)     # same iteration count on every rank
            spent += float(t)
        dist.barrier()
        torch.cuda.synchronize()
        engine.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        mhz, watts = nvml_clock()
        lm, ln, gm, gn = engine.kernel_times()
        engine.set_option(_capi.SIGLIP_OPT_KERNEL_TIMING, 0)
        mine = torch.tensor([e0.elapsed_time(e1) / args.steps, lm / max(ln, 1), gm / max(gn, 1),
                             (lm + gm) / args.steps], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rows = [[float(x) for x in r] for r in allr]
        results[label] = {"per_rank": rows, "coupled": coupled, "chunks": engine.world}
        if rank == 0:
            w = engine.world
            print(f"== {label} (W={w} chunk(s) per rank) ==", flush=True)
            for r, (ms, lk, gk, ksum) in enumerate(rows):
                print(f"  rank {r}: {ms:8.3f} ms/step ({ms / w:.3f} per chunk)  loss kernel {lk:.3f}  gradient kernel {gk:.3f}"
                      f"  kernels {ksum:.3f} ({100 * ksum / ms:.1f}% of step)", flush=True)
            steps = [r[0] for r in rows]
            print(f"  step ms: min {min(steps):.3f} max {max(steps):.3f} spread {100 * (max(steps) / min(steps) - 1):.1f}%",
                  flush=True)
        print(f"  [rank {rank}] {label}: SM clock after the timed steps {mhz} MHz, {watts} W", flush=True)
        dist.barrier()

    if "A" in args.phases:
        single = SigmoidLossEngine(B, D, dev, rank_world=(0, 1))
        run(single, "A uncoupled W=1", False)
        single.close()
    if "B" in args.phases and world > 1:
        lb = SigmoidLossEngine(B, D, dev, rank_world=(rank, world), loopback=True)
        for k in range(world):
            lb.debug_set_text_chunk(k, txt)
        run(lb, "B uncoupled loopback W=N", False)
        lb.close()
    if "E" in args.phases and world > 1:
        # per-step traces without any coupling: how much does a GPU's step time fluctuate from step to step (power-cap
        # control loop)? A coupled job pays max over ranks EVERY step: mean_s max_r t_r[s] >= max_r mean_s t_r[s].
        lb = SigmoidLossEngine(B, D, dev, rank_world=(rank, world), loopback=True)
        for k in range(world):
            lb.debug_set_text_chunk(k, txt)
        for _ in range(3):
            lb.fwd_bwd(img, txt, tp, bt, torch.bfloat16)
        dist.barrier()
        torch.cuda.synchronize()
        nsteps = 260
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
        evs[0].record()
        for s in range(nsteps):
            lb.fwd_bwd(img, txt, tp, bt, torch.bfloat16)
            evs[s + 1].record()
        torch.cuda.synchronize()
        ts = torch.tensor([evs[s].elapsed_time(evs[s + 1]) for s in range(nsteps)], device=dev, dtype=torch.float64)
        allt = [torch.empty_like(ts) for _ in range(world)]
        dist.all_gather(allt, ts)
        lb.close()
        if rank == 0:
            m = torch.stack(allt)[:, 120:]          # drop the burst -> sustained transition
            per_rank_mean = m.mean(1)
            per_rank_std = m.std(1)
            coupled_est = m.max(0).values.mean()
            # block maxima: if ranks can drift apart by up to k steps, the max is over k-step block means
            est = {}
            for kblk in (1, 2, 4, 10):
                nb = m.shape[1] // kblk
                mb = m[:, :nb * kblk].reshape(world, nb, kblk).mean(2)
                est[kblk] = float(mb.max(0).values.mean())
            print("== E per-step traces, uncoupled loopback W=N ==", flush=True)
            for r in range(world):
                print(f"  rank {r}: mean {float(per_rank_mean[r]):.3f} ms  std {float(per_rank_std[r]):.3f} ms "
                      f"({100 * float(per_rank_std[r] / per_rank_mean[r]):.2f}%)  min {float(m[r].min()):.3f} max {float(m[r].max()):.3f}",
                      flush=True)
            print(f"  max_r mean_s = {float(per_rank_mean.max()):.3f} ms;  mean_s max_r (per-step lock-step estimate) = "
                  f"{float(coupled_est):.3f} ms (+{100 * (float(coupled_est) / float(per_rank_mean.max()) - 1):.1f}%); "
                  f"block maxima {est}", flush=True)
            results["E traces"] = {"per_rank_mean": per_rank_mean.tolist(), "per_rank_std": per_rank_std.tolist(),
                                   "lockstep_estimate": float(coupled_est), "block_estimates": est,
                                   "trace_rank0": m[0, :64].tolist(), "trace_rank1": m[1, :64].tolist()}
        dist.barrier()
    if "C" in args.phases and world > 1:
        eng = SigmoidLossEngine(B, D, dev)
        run(eng, "C coupled W=N", True)
        if "T" in args.phases:
            # in-kernel timeline of the auxiliary warps (CTA 0) for three coupled steps: how long they wait for peer
            # flags, how long the pulls / folds take, and how much of the kernel is left after them
            eng.set_option(_capi.SIGLIP_OPT_AUX_TRACE, 1)
            for _ in range(3):
                eng.fwd_bwd(img, txt, tp, bt, torch.bfloat16)
            tr = eng.aux_trace()
            eng.set_option(_capi.SIGLIP_OPT_AUX_TRACE, 0)
            rows = []
            for (t0, tflag, tdone, tend, *_rest) in tr[-2 * world:]:          # the last step: L0 L1 G1 ... G0
                wait_us = (tflag - t0) / 1e3 if tflag else 0.0
                jobs_us = (tdone - (tflag if tflag else t0)) / 1e3
                kern_us = (tend - t0) / 1e3 if tend else float("nan")
                rows.append((wait_us, jobs_us, kern_us))
            tt = torch.tensor(rows, device=dev, dtype=torch.float64)
            allr = [torch.empty_like(tt) for _ in range(world)]
            dist.all_gather(allr, tt)
            if rank == 0:
                names = ["L0"] + [x for k in range(1, world) for x in (f"L{k}", f"G{k}")] + ["G0"]
                print("== T aux-warp timeline of the last coupled step (us): wait for the last peer flag | time from there "
                      "to jobs done | kernel (aux start -> last CTA done) ==", flush=True)
                for r in range(world):
                    print(f"  rank {r}: " + "  ".join(f"{n}:{w:.0f}|{j:.0f}|{k:.0f}" for n, (w, j, k) in
                                                    zip(names, allr[r].tolist())), flush=True)
                results["T trace"] = {"names": names, "per_rank": [a.tolist() for a in allr]}
        if "D" in args.phases:
            eng.set_option(_capi.SIGLIP_OPT_OVERLAP_REDUCE, 0)
            run(eng, "D coupled W=N, reduction at the end", True)
            eng.set_option(_capi.SIGLIP_OPT_OVERLAP_REDUCE, 1)
        eng.close()
    if rank == 0:
        a = results.get("A uncoupled W=1")
        c = results.get("C coupled W=N")
        b = results.get("B uncoupled loopback W=N")
        if a and c:
            t1_0 = a["per_rank"][0][0]
            t1_max = max(r[0] for r in a["per_rank"])
            tw = max(r[0] for r in c["per_rank"])
            print(f"FLOP-normalised efficiency W*t(1)/t(W): vs rank 0's single-chunk step {world * t1_0 / tw:.3f}; "
                  f"vs the slowest GPU's single-chunk step {world * t1_max / tw:.3f}", flush=True)
            if b:
                tb = max(r[0] for r in b["per_rank"])
                print(f"uncoupled W=N (slowest GPU) {tb:.3f} ms vs coupled {tw:.3f} ms: coupling costs "
                      f"{100 * (tw / tb - 1):.1f}%; W>1 data flow costs {100 * (tb / (world * t1_max) - 1):.1f}% over W x t(1)",
                      flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump({"world": world, "batch": B, "dim": D, "results": results}, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
