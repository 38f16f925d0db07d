"""Random-interleaving simulation of k_wgrad's mbarrier protocol (nerf_rs_b200/csrc/mlp_tc.cu) on the REAL work split
(nerf_debug_wgrad_partition): producer, MMA issuer (asynchronous tensor pipe), d(sigma) loader warp, four epilogue warps and
the all-warp accumulator flush, across a CTA's segments with ring stage/phase counters running on. Checks, for every CTA:
no deadlock; a ring stage is refilled only after the MMAs and all epilogue warps are done with it; MMAs and epilogue passes
start only after the stage's bytes AND its d(sigma) strip have landed; the accumulator is flushed only after the segment's last
MMA has executed, and the next segment's first MMA executes only after every warp has flushed. Three deliberately broken
variants of the protocol must be caught."""
import ctypes
import random

import numpy as np
import pytest

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib

STAGES = 3


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        assert self.pending > 0, "mbarrier arrival overflow"
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):          # try_wait.parity: the phase with this parity has completed
        return (self.phase & 1) != parity


def simulate(segments, rng, full_count=3, empty_count=5, flush_waits_done=True, sync_between_segments=True, max_steps=400000):
    """segments: list of iteration counts (half tiles) per segment of ONE CTA. full barrier = producer arrival + loader arrival
    + the bulk copies' byte count (modelled as a third arrival by the copy engine)."""
    full = [Bar(full_count) for _ in range(STAGES)]
    empty = [Bar(empty_count) for _ in range(STAGES)]
    done = Bar(1)
    st = dict(data=[-1] * STAGES, strip=[-1] * STAGES, mma_exec=-1, seg_mma_last=[None] * len(segments), flushed=[0] * len(segments),
              epi_done=[[-1] * STAGES for _ in range(4)], mma_stage_done=[-1] * STAGES, pipe=[], copies=[], sync=[0] * len(segments))
    base = np.concatenate([[0], np.cumsum(segments)]).astype(int)       # global iteration index of each segment's first iteration

    def producer():
        stage, phase = 0, 0
        for seg, n in enumerate(segments):
            if sync_between_segments and seg:
                yield lambda seg=seg: st["sync"][seg - 1] == 8
            for it in range(n):
                g = base[seg] + it
                yield lambda s=stage, p=phase: empty[s].done(p ^ 1)
                # refill: the previous occupant (g - STAGES) must be fully consumed
                if g >= STAGES:
                    assert st["mma_stage_done"][stage] >= g - STAGES, f"stage {stage} refilled under an in-flight MMA (iteration {g})"
                    assert all(e[stage] >= g - STAGES for e in st["epi_done"]), f"stage {stage} refilled under an epilogue pass (iteration {g})"
                full[stage].arrive()                       # arrive.expect_tx
                st["copies"].append((stage, g))            # the copy engine lands the bytes later
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1

    def loader():
        stage, phase = 0, 0
        for seg, n in enumerate(segments):
            if sync_between_segments and seg:
                yield lambda seg=seg: st["sync"][seg - 1] == 8
            for it in range(n):
                g = base[seg] + it
                yield lambda s=stage, p=phase: empty[s].done(p ^ 1)
                if g >= STAGES:
                    assert all(e[stage] >= g - STAGES for e in st["epi_done"]), f"d(sigma) strip {stage} rewritten under an epilogue pass"
                st["strip"][stage] = g
                full[stage].arrive()
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, phase = 0, 0
        for seg, n in enumerate(segments):
            if sync_between_segments and seg:
                yield lambda seg=seg: st["sync"][seg - 1] == 8
            for it in range(n):
                g = base[seg] + it
                yield lambda s=stage, p=phase: full[s].done(p)
                assert st["data"][stage] == g, f"MMA of iteration {g} issued before its bytes landed"
                st["pipe"].append(("mma", g, stage, seg))
                st["pipe"].append(("commit_empty", g, stage, seg))
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1
            st["pipe"].append(("commit_done", base[seg] + n - 1, None, seg))

    def epilogue(w):
        stage, phase = 0, 0
        for seg, n in enumerate(segments):
            if sync_between_segments and seg:
                yield lambda seg=seg: st["sync"][seg - 1] == 8
            for it in range(n):
                g = base[seg] + it
                yield lambda s=stage, p=phase: full[s].done(p)
                assert st["data"][stage] == g and st["strip"][stage] == g, f"epilogue pass of iteration {g} read a stale stage or d(sigma) strip"
                yield lambda: True                                   # (the pass takes a while)
                assert st["data"][stage] == g and st["strip"][stage] == g, f"stage {stage} changed under the epilogue pass of iteration {g}"
                st["epi_done"][w][stage] = g
                empty[stage].arrive()
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1
            yield from flush(seg)

    def service(_):
        for seg in range(len(segments)):
            yield from flush(seg)

    def flush(seg):
        if flush_waits_done:
            yield lambda seg=seg: done.done(seg & 1)
        assert st["mma_exec"] >= base[seg + 1] - 1, f"accumulator of segment {seg} flushed before its last MMA executed"
        assert st["mma_exec"] < base[seg + 1] or seg + 1 == len(segments), f"segment {seg + 1}'s MMAs overwrote the accumulator before the flush of segment {seg}"
        yield lambda: True
        assert st["mma_exec"] < base[seg + 1] or seg + 1 == len(segments), f"segment {seg + 1}'s MMAs overwrote the accumulator during the flush of segment {seg}"
        st["flushed"][seg] += 1
        st["sync"][seg] += 1                                       # __syncthreads at the end of the segment
        yield lambda seg=seg: st["sync"][seg] == 8

    roles = {"producer": producer(), "loader": loader(), "mma": mma()}
    roles.update({f"epi{w}": epilogue(w) for w in range(4)})
    roles["flush_warp2"] = service(2)                     # warp 2 (TMEM allocator) only flushes
    waiting = {}
    for name, g in list(roles.items()):
        try:
            waiting[name] = next(g)
        except StopIteration:
            roles.pop(name)
    # warps 0, 1, 3 (producer, MMA issuer, loader: single-thread loops above) take part in every segment's flush and
    # segment-end barrier as whole warps: modelled as three more flush participants
    for w in range(3):
        g = service(w)
        roles[f"flush{w}"] = g
        waiting[f"flush{w}"] = next(g)
    steps = 0
    while roles:
        steps += 1
        assert steps < max_steps, "simulation did not finish"
        choices = [n for n in roles if waiting[n]()]
        async_ok = bool(st["pipe"]) or bool(st["copies"])
        if not choices and not async_ok:
            raise AssertionError("deadlock: " + ", ".join(sorted(roles)))
        pick = rng.random()
        if async_ok and (not choices or pick < 0.45):
            if st["copies"] and (not st["pipe"] or rng.random() < 0.5):
                s, g = st["copies"].pop(0)
                st["data"][s] = g
                full[s].arrive()                       # complete_tx
            else:
                kind, g, s, seg = st["pipe"].pop(0)   # the tensor pipe executes in issue order
                if kind == "mma":
                    assert st["data"][s] == g, f"stage {s} was refilled before the MMA of iteration {g} executed"
                    if g == base[seg] and seg:
                        assert st["flushed"][seg - 1] == 8, f"first MMA of segment {seg} executed before every warp flushed segment {seg - 1}"
                    st["mma_exec"] = g
                elif kind == "commit_empty":
                    st["mma_stage_done"][s] = g
                    empty[s].arrive()
                else:
                    done.arrive()
            continue
        name = rng.choice(choices)
        try:
            waiting[name] = next(roles[name])
        except StopIteration:
            roles.pop(name)
    assert all(f == 8 for f in st["flushed"])
    return steps


def _partition(hidden, n_tiles, n_ctas=148):
    cfg = nb.default_config(hidden=hidden)
    out = np.zeros((n_ctas, 10), np.int32)
    nu = ctypes.c_int32(0)
    assert _lib.load().nerf_debug_wgrad_partition(ctypes.byref(cfg), n_ctas, n_tiles, out.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(nu)) == 0
    return [[2 * int(out[c, 3 + 3 * k] - out[c, 2 + 3 * k]) for k in range(out[c, 0])] for c in range(n_ctas)]


@pytest.mark.parametrize("hidden,n_tiles", [(256, 2048), (512, 512), (256, 7)])
def test_protocol_holds_on_the_real_work_split(hidden, n_tiles):
    rng = random.Random(hidden + n_tiles)
    ctas = [s for s in _partition(hidden, n_tiles) if s]
    multi = [s for s in ctas if len(s) > 1]
    assert multi or n_tiles < 16                     # at bench size some CTAs do work through two segments
    picked = rng.sample(multi, min(4, len(multi))) + rng.sample(ctas, min(4, len(ctas)))
    for segs in picked:
        simulate(segs, rng, max_steps=4_000_000)


@pytest.mark.parametrize("kw,needle", [
    (dict(full_count=2), "before its bytes landed|stale stage|changed under"),             # loader (or producer) arrival not counted
    (dict(empty_count=4), "refilled|rewritten|changed under|stale"),                      # one consumer missing from the empty barrier
    (dict(flush_waits_done=False), "flushed before its last MMA"),                        # flush without the done barrier
    (dict(sync_between_segments=False), "overwrote the accumulator|before every warp flushed"),   # no barrier between segments
])
def test_broken_protocols_are_caught(kw, needle):
    import re
    rng = random.Random(7)
    caught = 0
    for trial in range(40):
        try:
            simulate([10, 8, 6], rng, **kw)
        except AssertionError as e:
            assert re.search(needle + "|deadlock|overflow|did not finish", str(e)), str(e)
            caught += 1
    assert caught > 0
