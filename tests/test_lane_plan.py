"""CPU protocol check of the CTA-pair chain kernel (nerf_rs_b200/csrc/mlp_tc2.cu).

The kernel's five roles (weight producer, peer relay, MMA issuer, 16 epilogue warps, store warp) are restated here as
Python generators over a model of its mbarriers -- per-CTA barriers, multicast commits that land asynchronously and in
order, remote arrives -- and run under random interleavings on the REAL lane programs (host-only debug entry point).
The simulation asserts what the hardware would silently get wrong:
  * every MMA reads A panels / slot E written by the right job of the right tile, in BOTH CTAs, and the weight chunk
    it expects in the ring stage;
  * no panel, slot E or ring stage is overwritten while an MMA that reads it is still in flight, or (training) while a
    bulk store is still reading it;
  * an accumulator is drained before the lane's next GEMM overwrites it and complete before the epilogue reads it;
  * the run terminates (no deadlock) for odd/even tile counts, one or two live lanes, every supported geometry.
"""
import ctypes
import random

import numpy as np
import pytest

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib

LANE_OP = np.dtype([("w_off", "<u4"), ("n", "<u2"), ("a_slot", "u1"), ("kflags", "u1")])   # kflags: K16 steps | 0x40 half 1 | 0x80 first
LANE_GEMM = np.dtype([("op_begin", "<u2"), ("op_end", "<u2")])
LANE_JOB = np.dtype([("kind", "u1"), ("enc", "u1"), ("ncols", "<u2"), ("bias_off", "<u2"), ("save_slot", "<i2"),
                     ("enc_save_slot", "<i2"), ("mask_slot", "<i2"), ("out_slot", "u1"), ("pad0", "u1"), ("pad1", "<u2")])
K_STAGES, SLOT_E, EPI_WARPS = 4, 4, 16
HIDDEN_KINDS = (1, 2, 6, 7)   # RELU, LINEAR, DMASK, DCOPY


def lane_plan(cfg, program):
    lib = _lib.load()
    ops, gemms, jobs = np.zeros(256, LANE_OP), np.zeros(32, LANE_GEMM), np.zeros(32, LANE_JOB)
    n = [ctypes.c_int32(256), ctypes.c_int32(32), ctypes.c_int32(32)]
    rc = lib.nerf_debug_lane_plan(ctypes.byref(cfg), program, ops.ctypes.data, ctypes.byref(n[0]), gemms.ctypes.data,
                                  ctypes.byref(n[1]), jobs.ctypes.data, ctypes.byref(n[2]))
    assert rc == 0, rc
    return ops[:n[0].value], gemms[:n[1].value], jobs[:n[2].value]


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier's count in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):   # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Sched:   # LaneSched of mlp_tc2.cu (wide: one lane walks every pair tile of the cluster)
    def __init__(self, cluster, n_clusters, n_pairs, n_pos, wide=False):
        self.p0, self.p1, self.pos, self.n_pairs, self.n_pos, self.wide = cluster, cluster + n_clusters, 0, n_pairs, n_pos, wide
        self.stride = n_clusters if wide else 2 * n_clusters

    def __iter__(self):
        while self.p0 < self.n_pairs:
            yield self.pos, (2 if (not self.wide and self.p1 < self.n_pairs) else 1), self.p0, self.p1
            self.pos += 1
            if self.pos == self.n_pos:
                self.pos = 0
                self.p0 += self.stride
                if not self.wide:
                    self.p1 += self.stride


class Sim:
    def __init__(self, ops, gemms, jobs, n_pairs, n_clusters, save, seed, mutation=None):
        self.ops, self.gemms, self.jobs, self.n_pairs, self.C, self.save = ops, gemms, jobs, n_pairs, n_clusters, save
        self.wide = any(int(j["ncols"]) > 256 for j in jobs) or int(ops["a_slot"].max()) > 4
        self.e_slot = 8 if self.wide else SLOT_E
        self.mutation = mutation   # deliberately broken protocol variants: the simulation must reject them
        self.rng = random.Random(seed)
        self.nG = len(gemms)
        B = lambda c: [Bar(c) for _ in range(2)]
        # [cta] barriers
        self.full = [[Bar(2 if r == 0 else 1) for _ in range(K_STAGES)] for r in range(2)]
        self.empty = [[Bar(1) for _ in range(K_STAGES)] for r in range(2)]
        self.acc_full = [B(1) for r in range(2)]
        self.epi_done = B(2 * EPI_WARPS)           # leader only
        self.save_ready = [B(EPI_WARPS) for r in range(2)]
        self.save_free = [B(1) for r in range(2)]
        # state: smem contents are version tags
        self.stage = [[None] * K_STAGES for _ in range(2)]            # (w_off)
        self.slot = [[[None] * 9 for _ in range(2)] for _ in range(2)]  # [cta][lane][slot] = (tile_pair, producing job)
        self.acc = [[None, None] for _ in range(2)]                      # [cta][lane] = (pair, gemm) when complete
        self.acc_drained = [[True, True] for _ in range(2)]
        self.inflight = []        # MMAs issued, not yet completed: dicts
        self.async_q = []         # in-order completion queue of the MMA thread: ("mma", rec) | ("commit", fn)
        self.store_reads = [[set(), set()] for _ in range(2)]   # [cta][lane] slots a pending bulk store still reads

    # ---- async tensor pipe: completes in issue order at random times
    def pump(self):
        while self.async_q and self.rng.random() < 0.6:
            kind, x = self.async_q.pop(0)
            if kind == "mma":
                self.inflight.remove(x)
            else:
                x()

    def reading(self, cta, lane, slot):
        return any(m["lane"] == lane and m["a_slot"] == slot for m in self.inflight)

    def shareable(self, g):
        return int(self.gemms[g]["op_end"]) - int(self.gemms[g]["op_begin"]) <= K_STAGES

    # ---- roles (generators yield a predicate to wait on)
    def producer(self, cta, relay):
        stage, phase = 0, 0
        for g, nl, _, _ in Sched(0, self.C, self.n_pairs, self.nG, self.wide):
            reps = 2 if (nl == 2 and not self.shareable(g)) else 1
            for _ in range(reps):
                for i in range(int(self.gemms[g]["op_begin"]), int(self.gemms[g]["op_end"])):
                    if not relay:
                        if self.mutation != "no_empty_wait":
                            yield lambda s=stage, p=phase: self.empty[cta][s].done(p ^ 1)
                        assert not any(m["stage"] == stage for m in self.inflight), "ring stage overwritten while an MMA reads it"
                        self.stage[cta][stage] = int(self.ops[i]["w_off"])
                        self.full[cta][stage].arrive()            # expect_tx arrive + bytes landed (collapsed)
                    else:
                        yield lambda s=stage, p=phase: self.full[1][s].done(p)
                        self.full[0][stage].arrive()               # remote arrive on the leader
                    stage += 1
                    if stage == K_STAGES:
                        stage, phase = 0, phase ^ 1

    def mma(self):
        stage, phase, done_phase = 0, 0, [0, 0]
        for g, nl, pr0, pr1 in Sched(0, self.C, self.n_pairs, self.nG, self.wide):
            shared = nl == 2 and self.shareable(g)
            stage0, phase0 = stage, phase
            for ln in range(nl):
                reuse, release = shared and ln == 1, (not shared) or ln == 1
                if self.mutation == "release_on_first_lane":
                    release = (not shared) or ln == 0
                if reuse:
                    stage, phase = stage0, phase0
                pair = pr1 if ln else pr0
                if self.mutation != "no_epi_done_wait":
                    yield lambda l=ln: self.epi_done[l].done(done_phase[l])
                done_phase[ln] ^= 1
                for cta in range(2):
                    assert self.acc_drained[cta][ln], "accumulator overwritten before the epilogue drained it"
                    self.acc_drained[cta][ln] = False
                    self.acc[cta][ln] = None
                ob, oe = int(self.gemms[g]["op_begin"]), int(self.gemms[g]["op_end"])
                for i in range(ob, oe):
                    op = self.ops[i]
                    if not reuse:
                        yield lambda s=stage, p=phase: self.full[0][s].done(p)
                    for cta in range(2):
                        assert self.stage[cta][stage] == int(op["w_off"]), f"wrong weight chunk in ring stage (gemm {g} op {i} cta {cta})"
                        tag = self.slot[cta][ln][int(op["a_slot"])]
                        want_job = self.expected_writer(g, int(op["a_slot"]))
                        assert tag == (pair, want_job), f"A operand slot {op['a_slot']} holds {tag}, expected {(pair, want_job)} (gemm {g})"
                    rec = dict(lane=ln, a_slot=int(op["a_slot"]), stage=stage)
                    self.inflight.append(rec)
                    self.async_q.append(("mma", rec))
                    if release:
                        self.async_q.append(("commit", lambda s=stage: [self.empty[c][s].arrive() for c in range(2)]))
                    if i + 1 == oe:
                        def fin(l=ln, pr=pair, gg=g):
                            for c in range(2):
                                self.acc[c][l] = (pr, gg)
                                self.acc_full[c][l].arrive()
                        self.async_q.append(("commit", fin))
                    stage += 1
                    if stage == K_STAGES:
                        stage, phase = 0, phase ^ 1
                    yield None

    def expected_writer(self, g, slot):
        """Index of the job (0 = prologue) whose output GEMM g must find in `slot`."""
        for jg in range(g, -1, -1):   # job jg is the epilogue of GEMM jg-1; jobs 0..g precede GEMM g
            j = self.jobs[jg]
            if slot == self.e_slot:
                if jg == 0 or int(j["enc"]) != 0:
                    return jg
            elif jg > 0 and int(j["kind"]) in HIDDEN_KINDS and int(j["out_slot"]) <= slot < int(j["out_slot"]) + (int(j["ncols"]) + 63) // 64:
                return jg
        raise AssertionError(f"no producer for slot {slot} before gemm {g}")

    def epilogue(self, cta, warp):
        aph, sph = [0, 0], [0, 0]
        stride = self.C if self.wide else 2 * self.C
        for p, nl, pr0, pr1 in Sched(0, self.C, self.n_pairs, self.nG, self.wide):
            for ln in range(nl):
                pair = pr1 if ln else pr0
                first_tile, has_next, last = pair < stride, pair + stride < self.n_pairs, p == self.nG - 1

                def write(slot, tag):
                    assert not self.reading(cta, ln, slot), f"slot {slot} rewritten while an MMA reads it"
                    assert slot not in self.store_reads[cta][ln], f"slot {slot} rewritten while a bulk store reads it"
                    self.slot[cta][ln][slot] = tag

                def signal():
                    self.epi_done[ln].arrive()

                if p == 0 and first_tile:
                    if self.save:
                        yield lambda: self.save_free[cta][ln].done(sph[ln] ^ 1)
                        sph[ln] ^= 1
                    write(self.e_slot, (pair, 0))
                    signal()
                    if self.save:
                        self.save_ready[cta][ln].arrive()
                if self.save:
                    if self.mutation != "no_save_free_wait":
                        yield lambda: self.save_free[cta][ln].done(sph[ln] ^ 1)
                    sph[ln] ^= 1
                if last and has_next:
                    write(self.e_slot, (pair + stride, 0))
                j = self.jobs[p + 1]
                yield lambda: self.acc_full[cta][ln].done(aph[ln])
                aph[ln] ^= 1
                assert self.acc[cta][ln] == (pair, p), "epilogue read an accumulator that is not complete"
                yield None   # (tcgen05.ld of all warps interleave)
                if int(j["kind"]) in HIDDEN_KINDS:
                    for s in range(int(j["out_slot"]), int(j["out_slot"]) + (int(j["ncols"]) + 63) // 64):
                        write(s, (pair, p + 1))
                if int(j["enc"]) != 0:
                    write(self.e_slot, (pair, p + 1))
                self.drain_count[cta][ln] += 1
                if self.drain_count[cta][ln] == EPI_WARPS:
                    self.drain_count[cta][ln] = 0
                    self.acc_drained[cta][ln] = True
                if not last or has_next:
                    signal()
                if self.save:
                    self.save_ready[cta][ln].arrive()

    def store(self, cta):
        rph = [0, 0]
        stride = self.C if self.wide else 2 * self.C
        for p, nl, pr0, pr1 in Sched(0, self.C, self.n_pairs, self.nG, self.wide):
            for ln in range(nl):
                pair = pr1 if ln else pr0
                steps = []
                if p == 0 and pair < stride:
                    steps.append({self.e_slot} if int(self.jobs[0]["enc_save_slot"]) >= 0 else set())
                j = self.jobs[p + 1]
                rd = set()
                if int(j["save_slot"]) >= 0:
                    rd |= set(range(int(j["out_slot"]), int(j["out_slot"]) + int(j["ncols"]) // 64))
                if int(j["enc_save_slot"]) >= 0 or (p == self.nG - 1 and pair + stride < self.n_pairs and int(self.jobs[0]["enc_save_slot"]) >= 0):
                    rd.add(self.e_slot)
                steps.append(rd)
                for rd in steps:
                    yield lambda: self.save_ready[cta][ln].done(rph[ln])
                    rph[ln] ^= 1
                    self.store_reads[cta][ln] = set(rd)
                    yield None   # the copy engine reads shared memory for a while
                    yield None
                    self.store_reads[cta][ln] = set()
                    self.save_free[cta][ln].arrive()

    def run(self):
        self.drain_count = [[0, 0] for _ in range(2)]
        roles = [self.mma(), self.producer(0, False), self.producer(1, False), self.producer(1, True)]
        roles += [self.epilogue(c, w) for c in range(2) for w in range(EPI_WARPS)]
        if self.save:
            roles += [self.store(0), self.store(1)]
        waiting = {}
        live = list(range(len(roles)))
        idle = 0
        while live:
            self.pump()
            r = self.rng.choice(live)
            cond = waiting.get(r)
            if cond is not None and not cond():
                idle += 1
                if idle > 200000:
                    while self.async_q:   # let the tensor pipe finish before declaring a deadlock
                        self.pump()
                    if not any(waiting.get(x) is None or waiting[x]() for x in live):
                        raise AssertionError(f"deadlock: {len(live)} roles blocked")
                    idle = 0
                continue
            idle = 0
            try:
                waiting[r] = next(roles[r])
            except StopIteration:
                live.remove(r)
                waiting.pop(r, None)
        while self.async_q:
            self.pump()
        assert not self.inflight


GEOMS = {
    "ns256": dict(hidden=256),
    "ns128": dict(hidden=128),
    "ns64": dict(hidden=64),
    "as_shipped": dict(hidden=100, xyz_freqs=0, dir_freqs=-1, skip_layer=0, use_rgb_head=0),
    "noskip_nodir": dict(hidden=256, skip_layer=0, dir_freqs=-1),
    "ns512": dict(hidden=512),
}


@pytest.mark.parametrize("name", list(GEOMS))
@pytest.mark.parametrize("program", [0, 1, 2])
def test_lane_program_shape(name, program):
    ops, gemms, jobs = lane_plan(nb.default_config(**GEOMS[name]), program)
    assert len(jobs) == len(gemms) + 1 and len(ops) <= 160 and len(gemms) <= 16     # kernel-parameter table capacities
    assert int(gemms[0]["op_begin"]) == 0 and int(gemms[-1]["op_end"]) == len(ops)
    assert all(int(a["op_end"]) == int(b["op_begin"]) for a, b in zip(gemms[:-1], gemms[1:]))
    assert all(int(o["n"]) % 16 == 0 and int(o["n"]) <= 256 and (int(o["kflags"]) & 7) in (1, 2, 4) for o in ops)   # cta_group::2 shapes
    assert all(int(ops[int(gm["op_begin"])]["kflags"]) & 0x80 for gm in gemms)        # every GEMM starts by overwriting its accumulator
    assert all(int(j["ncols"]) in (16, 64, 128, 192, 256, 512) for j in jobs[1:])
    if program != 1:   # training programs save every hidden panel they produce
        assert all(int(j["save_slot"]) >= 0 for j in jobs[1:] if int(j["kind"]) in HIDDEN_KINDS)


@pytest.mark.parametrize("name", ["ns256", "ns64", "as_shipped", "ns512"])
@pytest.mark.parametrize("program", [0, 1, 2])
@pytest.mark.parametrize("n_pairs,n_clusters", [(1, 1), (2, 1), (3, 1), (5, 2), (4, 1)])
def test_protocol_random_interleavings(name, program, n_pairs, n_clusters):
    ops, gemms, jobs = lane_plan(nb.default_config(**GEOMS[name]), program)
    for seed in range(3):
        Sim(ops, gemms, jobs, n_pairs, n_clusters, save=program != 1, seed=seed).run()


@pytest.mark.parametrize("mutation", ["no_empty_wait", "no_epi_done_wait", "release_on_first_lane", "no_save_free_wait"])
def test_simulation_rejects_broken_protocols(mutation):
    """The checker itself is checked: each protocol bug must trip an assertion in at least one interleaving."""
    ops, gemms, jobs = lane_plan(nb.default_config(hidden=256), 0)
    caught = 0
    for seed in range(6):
        try:
            Sim(ops, gemms, jobs, 4, 1, save=True, seed=seed, mutation=mutation).run()
        except AssertionError:
            caught += 1
    assert caught > 0, f"mutation {mutation} went unnoticed"
