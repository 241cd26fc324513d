"""A CPU model of the cross-rank protocol of the fused step (csrc/siglip_capi.cu `fused_impl`, kernels' auxiliary warps in
csrc/siglip_kernels.cu): W ranks, each running the launch sequence  L0  L1 G1 ... L(W-1) G(W-1)  G0  per step, with every
wait / signal / buffer access of the real schedule as an event, executed under RANDOM interleavings (any rank may be
arbitrarily late at any point). The model checks what the GPU runs can only sample:

  * no deadlock: some action is always enabled until every rank has finished every step;
  * every text pull reads the owner's exported slot while it holds the text of THIS step (not the previous one, not
    half-overwritten by the next);
  * every fold reads a contribution slot that its producer has finished writing for THIS backward and has not started
    overwriting for the next;
  * the last gradient launch adds the folded sum only after the last fold has completed.

Flags are the monotonic counters of the implementation: kind 0 text ready (value s), kind 1 contribution of gradient
slot j ready (value (n-1) W + j), kind 2 "I pulled everyone's text" (s), kind 3 "I read everyone's contributions" (n).
This is a model of the protocol, not of the C code; the mapping is one action per auxiliary job / end-of-launch signal.
"""
import random

import pytest

from distributed_sigmoid_loss_b200 import chunk_schedule


class Rank:
    def __init__(self, r, world, bidir):
        self.r, self.W = r, world
        self.sched = chunk_schedule(r, world, bidir)          # owner of the chunk of step k
        self.offset = [(self.sched[k] - r) % world for k in range(world)]
        # flags[kind][peer]
        self.flags = [[0] * world for _ in range(4)]
        # exported buffers with (version, state): state "ok" or "writing"
        self.txt_slot = (0, "ok")
        self.slots = {c: (0, "ok") for c in range(world)}     # my contribution to owner c
        self.gathered = {}                                    # chunk -> step whose text I hold locally
        self.fold_done = 0
        self.folded = set()
        self.step = 1
        self.pc = 0                                           # index into the action list of the current step
        self.actions = []
        self.done_steps = 0


def producer_of(world, bidir, r, j):
    """Rank whose gradient slot j produced the contribution for owner r (mirrors `pr` in fused_impl)."""
    for p in range(world):
        if chunk_schedule(p, world, bidir)[j] == r:
            return p
    raise AssertionError


def build_actions(rk, bidir):
    """The event list of one step of rank rk, in stream / job order. Each action: (name, guard(ranks) -> bool, effect)."""
    W, r = rk.W, rk.r
    s = n = rk.step
    base = (n - 1) * W
    acts = []

    def add(name, guard, effect):
        acts.append((name, guard, effect))

    def sig(kind, value):
        def eff(ranks):
            for p in ranks:
                p.flags[kind][r] = max(p.flags[kind][r], value)
        return eff

    if W > 1:
        # ---- L0: copy own text into the exported slot (wait flag 2 of everyone >= s-1), signal flag 0 = s ----
        add("L0.copy.begin", lambda ranks: all(rk.flags[2][p] >= s - 1 for p in range(W)),
            lambda ranks: setattr(rk, "txt_slot", (s, "writing")))
        add("L0.copy.end", lambda ranks: True, lambda ranks: (setattr(rk, "txt_slot", (s, "ok")), sig(0, s)(ranks)))
        # wait-only job: everyone has read my contribution slots of backward n-1 (flag 3)
        add("L0.wait3", lambda ranks: all(rk.flags[3][p] >= n - 1 for p in range(W)), lambda ranks: None)
        for k in range(0, W):
            # pull of the chunk of step k+1 inside L(k)
            if k + 1 < W:
                o = rk.sched[k + 1]

                def pull_guard(ranks, o=o):
                    return rk.flags[0][o] >= s

                def pull_eff(ranks, o=o):
                    ver, state = ranks[o].txt_slot
                    assert state == "ok" and ver == s, f"rank {r} step {s}: pulled text of rank {o} in state {ver, state}"
                    rk.gathered[o] = s
                add(f"L{k}.pull({o})", pull_guard, pull_eff)
            if k >= 1:
                o = rk.sched[k]
                add(f"L{k}.mma", lambda ranks, o=o: True,
                    lambda ranks, o=o: _check(rk.gathered.get(o) == s, f"rank {r}: L{k} scores stale text of {o}"))
            if k == W - 1:
                add("L(W-1).end", lambda ranks: True, sig(2, s))
            if k >= 1:
                j = k
                c = rk.sched[k]
                # G(j): epilogue writes my contribution to owner c; aux fold of the contribution produced one slot earlier
                add(f"G{j}.write.begin", lambda ranks: True,
                    lambda ranks, c=c: rk.slots.__setitem__(c, (n, "writing")))
                if j >= 2:
                    pr = producer_of(W, bidir, r, j - 1)

                    def fold_guard(ranks, pr=pr, j=j):
                        return rk.flags[1][pr] >= base + j - 1

                    def fold_eff(ranks, pr=pr, j=j):
                        ver, state = ranks[pr].slots[r]
                        assert state == "ok" and ver == n, f"rank {r} bwd {n}: folded slot of rank {pr} in state {ver, state}"
                        rk.folded.add((n, pr))
                    add(f"G{j}.fold({pr})", fold_guard, fold_eff)
                add(f"G{j}.end", lambda ranks: True,
                    lambda ranks, c=c, j=j: (rk.slots.__setitem__(c, (n, "ok")), sig(1, base + j)(ranks)))
        # ---- G0: last fold, then the dtxt tiles add the folded sum, then "I read everything" ----
        pr = producer_of(W, bidir, r, W - 1)

        def last_guard(ranks, pr=pr):
            return rk.flags[1][pr] >= base + W - 1

        def last_eff(ranks, pr=pr):
            ver, state = ranks[pr].slots[r]
            assert state == "ok" and ver == n, f"rank {r} bwd {n}: last fold read slot of rank {pr} in state {ver, state}"
            rk.folded.add((n, pr))
            rk.fold_done = n
        add(f"G0.fold({pr})", last_guard, last_eff)
        add("G0.p1", lambda ranks: rk.fold_done >= n,
            lambda ranks: _check(len([1 for (m, _) in rk.folded if m == n]) == W - 1,
                                 f"rank {r} bwd {n}: dtxt written with {len(rk.folded)} folds"))
        add("G0.end", lambda ranks: True, sig(3, n))
    else:
        add("L0", lambda ranks: True, lambda ranks: None)
        add("G0", lambda ranks: True, lambda ranks: None)
    return acts


def _check(cond, msg):
    assert cond, msg


def simulate(world, steps, bidir, seed, lag_rank=None):
    rng = random.Random(seed)
    ranks = [Rank(r, world, bidir) for r in range(world)]
    for rk in ranks:
        rk.actions = build_actions(rk, bidir)
    executed = 0
    while any(rk.done_steps < steps for rk in ranks):
        enabled = []
        for rk in ranks:
            if rk.done_steps >= steps:
                continue
            name, guard, _ = rk.actions[rk.pc]
            # the aux jobs of one launch run in order and the launches of a rank run in order: one enabled action per rank
            if guard(ranks):
                enabled.append(rk)
        assert enabled, "deadlock: " + "; ".join(
            f"rank {rk.r} step {rk.step} at {rk.actions[rk.pc][0]}" for rk in ranks if rk.done_steps < steps)
        # a lagging rank is picked rarely: the others run ahead as far as the protocol lets them
        weights = [0.05 if (lag_rank is not None and rk.r == lag_rank) else 1.0 for rk in enabled]
        rk = rng.choices(enabled, weights)[0]
        rk.actions[rk.pc][2](ranks)
        executed += 1
        rk.pc += 1
        if rk.pc == len(rk.actions):
            rk.done_steps += 1
            rk.step += 1
            rk.pc = 0
            rk.actions = build_actions(rk, bidir)
    return executed


@pytest.mark.parametrize("bidir", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
def test_fused_step_protocol_under_random_interleavings(world, bidir):
    for seed in range(40):
        simulate(world, steps=4, bidir=bidir, seed=seed)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_step_protocol_with_a_rank_that_is_always_late(world):
    for lag in range(world):
        for seed in range(8):
            simulate(world, steps=3, bidir=False, seed=100 * lag + seed, lag_rank=lag)


def test_the_model_catches_a_missing_wait():
    """Sanity of the model itself: without the wait on the owner's text-ready flag a fast rank pulls a slot that still
    holds the previous step's text (or is being overwritten), and without the wait on the contribution-ready flag a fold
    reads a slot its producer has not finished; the model must notice both. (The waits on flags 2 and 3 — "peers are done
    with my buffers" — cannot be violated in this schedule even when removed: a rank can only reach the overwrite after
    pulling text that its peer publishes after it is done. They are kept in the kernels as a cheap second line.)"""
    for victim in ("pull(", "fold("):
        hits = 0
        orig = build_actions

        def broken(rk, bidir, victim=victim, orig=orig):
            acts = orig(rk, bidir)
            return [(n, (lambda ranks: True) if victim in n else g, e) for (n, g, e) in acts]
        globals()["build_actions"] = broken
        try:
            for seed in range(60):
                try:
                    simulate(3, steps=3, bidir=False, seed=seed, lag_rank=seed % 3)
                except AssertionError as ex:
                    assert "deadlock" not in str(ex)
                    hits += 1
        finally:
            globals()["build_actions"] = orig
        assert hits > 0, victim


def test_buffer_reuse_waits_are_implied_by_the_data_dependencies():
    """Documented property: dropping the flag-2 / flag-3 waits of L0 changes nothing the model can observe."""
    orig = build_actions

    def relaxed(rk, bidir):
        acts = orig(rk, bidir)
        return [(n, (lambda ranks: True) if n in ("L0.copy.begin", "L0.wait3") else g, e) for (n, g, e) in acts]
    globals()["build_actions"] = relaxed
    try:
        for world in (2, 3, 4, 5):
            for seed in range(30):
                simulate(world, steps=4, bidir=(seed % 2 == 1), seed=seed, lag_rank=seed % world)
    finally:
        globals()["build_actions"] = orig


def test_loss_epilogue_slab_classes_cover_exactly_the_special_elements():
    """CPU model of the per-slab path choice of the loss epilogue (csrc/siglip_kernels.cu, `slab_diag` / `slab_edge`):
    a 32x32 slab whose first row and first column are multiples of 32 holds positive pairs (row == col, own chunk) iff
    row0 == col0, and holds masked elements (row >= M or col >= N) iff it crosses the border — for every tile shape of
    cta_group 1 (128x256) and 2 (256x256), aligned and ragged batches, including the fully masked extra tile of an odd
    tile row under multicast. Everything else may take the fast path, which applies neither labels nor masks."""
    for cg in (1, 2):
        tile_m = 128 * cg
        for (M, N) in [(256, 256), (512, 512), (300, 136), (1000, 1000), (40, 24), (1024, 520)]:
            tiles_m, tiles_n = -(-M // tile_m), -(-N // 256)
            for m_blk in range(tiles_m + 1):          # + 1: the fully masked tile a 2-tile cluster may be handed
                for n_blk in range(tiles_n):
                    tile_m0, tile_n0 = m_blk * tile_m, n_blk * 256
                    edge = (tile_m0 + tile_m > M) or (tile_n0 + 256 > N)
                    diag = (tile_m0 < tile_n0 + 256) and (tile_n0 < tile_m0 + tile_m)
                    for cta_rank in range(cg):
                        for q in range(4):
                            row0 = tile_m0 + cta_rank * 128 + q * 32
                            for cgrp in range(4):
                                for c in range(2):
                                    col0 = tile_n0 + cgrp * 64 + c * 32
                                    has_pos = any(row0 + l == col0 + j for l in range(32) for j in range(32)
                                                  if row0 + l < M and col0 + j < N)
                                    has_masked = (row0 + 32 > M) or (col0 + 32 > N)
                                    slab_diag = diag and row0 == col0
                                    slab_edge = edge and ((row0 + 32 > M) or (col0 + 32 > N))
                                    assert slab_edge == has_masked
                                    # a slab with positives is always routed to the general path
                                    assert (not has_pos) or slab_diag or slab_edge
                                    # and a slab routed to the fast path has neither positives nor masked elements
                                    if not (slab_diag or slab_edge):
                                        assert not has_pos and not has_masked
