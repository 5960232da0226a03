"""Completion flags of the fused tile gather (sharding.FusedTileGather, sync="flags"), checked without a GPU: the protocol's
bookkeeping runs against a fake library that records every stream operation it would enqueue (flag writes / waits,
cross-context orderings); a discrete-event simulation then executes the recorded streams of all ranks under random
interleavings and checks that
  * nothing deadlocks (every stream drains),
  * rank 0 never reads an assembled frame before every rank has stored its rows of it, and
  * no rank overwrites a frame slot before rank 0 has finished reading the slot's previous frame.
Frame sequences cover the round-robin deal over several contexts, the single-context pass bench.py's roofline arm runs in
between, and arbitrary slot / context mixes."""
import random

import pytest

import raymarching_engine_b200.sharding as sh
from raymarching_engine_b200 import _lib


class FakeCtx:
    def __init__(self, rank, index):
        self.handle = (rank, index)
        self.ops = []                  # the stream: executed in order

    def last_error(self):
        return "fake"


class FakeLib:
    """stands in for _lib.lib: the stream-ordered calls append to the context's op list"""

    def __init__(self, ctx_by_handle):
        self.ctx = ctx_by_handle

    def rmb_stream_write_u32(self, handle, addr, value):
        self.ctx[handle].ops.append(("write", addr, value))
        return _lib.RMB_OK

    def rmb_stream_wait_geq_u32(self, handle, addr, value):
        self.ctx[handle].ops.append(("wait", addr, value))
        return _lib.RMB_OK

    def rmb_ctx_wait_ctx(self, handle, other):
        o = self.ctx[other]
        self.ctx[handle].ops.append(("wait_ctx", other, len(o.ops)))      # everything enqueued on `other` so far
        return _lib.RMB_OK

    def rmb_ctx_set_gather_target(self, handle, ptr, nbytes):
        return _lib.RMB_OK


def make_gather(rank, world, nctx, slots, lib_ctx):
    g = object.__new__(sh.FusedTileGather)
    g.rank, g.world, g.slots = rank, world, slots
    g.contexts = [FakeCtx(rank, k) for k in range(nctx)]
    for c in g.contexts:
        lib_ctx[c.handle] = c
    g._index = {id(c): k for k, c in enumerate(g.contexts)}
    g._gen = [0] * slots
    g._cur, g._pending, g._last_ctx = {}, [[] for _ in g.contexts], {}
    g.sync, g._flags_addr = "flags", 1 << 20
    g.ptrs, g.nbytes = [0] * slots, 0
    return g


def run_protocol(frames, world, nctx, slots, seed):
    """frames: list of (context index, slot).  Returns the global execution order of the marker ops."""
    lib_ctx = {}
    gathers = [make_gather(r, world, nctx, slots, lib_ctx) for r in range(world)]
    fake = FakeLib(lib_ctx)
    real = _lib.lib
    _lib.lib = fake
    try:
        for c, slot in frames:
            for g in gathers:                                   # every rank issues the same calls in the same order
                ctx = g.contexts[c]
                g.aim(ctx, slot)
                s, gen = g._cur[c]
                ctx.ops.append(("store", g.rank, s, gen))        # the display kernel's stores into the slot
                g.complete(ctx)
                if g.rank == 0:
                    ctx.ops.append(("read", s, gen))             # the caller's read of the assembled frame
    finally:
        _lib.lib = real
    # ---- simulate
    rng = random.Random(seed)
    flags, pos, order = {}, {h: 0 for h in lib_ctx}, []
    streams = list(lib_ctx)
    while any(pos[h] < len(lib_ctx[h].ops) for h in streams):
        ready = []
        for h in streams:
            if pos[h] >= len(lib_ctx[h].ops):
                continue
            op = lib_ctx[h].ops[pos[h]]
            if op[0] == "wait" and flags.get(op[1], 0) < op[2]:
                continue
            if op[0] == "wait_ctx" and pos[op[1]] < op[2]:
                continue
            ready.append(h)
        assert ready, "deadlock: " + str({h: lib_ctx[h].ops[pos[h]] for h in streams if pos[h] < len(lib_ctx[h].ops)})
        h = rng.choice(ready)
        op = lib_ctx[h].ops[pos[h]]
        pos[h] += 1
        if op[0] == "write":
            flags[op[1]] = op[2]
        elif op[0] in ("store", "read"):
            order.append(op)
    return order


def check_order(order, world):
    stored, read = {}, {}
    for t, op in enumerate(order):
        if op[0] == "store":
            _, rank, s, gen = op
            if gen > 1:
                assert (s, gen - 1) in read, f"rank {rank} overwrites slot {s} (generation {gen}) before rank 0 has read generation {gen - 1}"
            stored.setdefault((s, gen), set()).add(rank)
        else:
            _, s, gen = op
            assert stored.get((s, gen), set()) == set(range(world)), f"rank 0 reads slot {s} generation {gen} before every rank has stored it"
            read[(s, gen)] = t


@pytest.mark.parametrize("world", [2, 3, 8])
def test_round_robin_frames_then_single_context_pass_then_round_robin(world):
    nctx, slots = 3, 6
    frames = [(f % nctx, f % slots) for f in range(20)]
    frames += [(0, f % slots) for f in range(9, 22)]             # bench.py's roofline arm: every frame on context 0
    frames += [(f % nctx, f % slots) for f in range(40, 61)]
    for seed in range(12):
        check_order(run_protocol(frames, world, nctx, slots, seed), world)


def test_arbitrary_slot_and_context_sequences():
    rng = random.Random(7)
    for trial in range(40):
        nctx, slots, world = rng.randint(1, 4), rng.randint(1, 8), rng.randint(2, 5)
        frames = [(rng.randrange(nctx), rng.randrange(slots)) for _ in range(rng.randint(1, 60))]
        check_order(run_protocol(frames, world, nctx, slots, trial), world)


def test_the_simulation_catches_a_broken_protocol(monkeypatch):
    """the checker is not vacuous: without the consumed[] wait a rank overwrites a slot rank 0 has not read yet"""
    orig = sh.FusedTileGather._begin

    def no_flow_control(self, context, slot):
        rank, self.rank = self.rank, 0                    # everybody behaves like rank 0: writes, never waits
        try:
            orig(self, context, slot)
        finally:
            self.rank = rank
    monkeypatch.setattr(sh.FusedTileGather, "_begin", no_flow_control)
    frames = [(f % 2, f % 2) for f in range(12)]
    with pytest.raises(AssertionError, match="overwrites slot"):
        for seed in range(50):
            check_order(run_protocol(frames, 2, 2, 2, seed), 2)
