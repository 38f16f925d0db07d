"""Helpers shared by the plan tests: fetch the per-tile program tables from libnerf_b200.so
(host-only debug entry points) and emulate them with numpy.

The emulation runs the SAME tables the CUDA kernels execute (MMA ops, epilogue jobs, packed
weight chunks, weight-gradient units) but in float32 numpy, so the result can be compared with
the torch oracle: it proves the host-built schedule computes the reference MLP
(src/model.rs:97-131) and its gradients, independently of the device mechanics.
"""
import ctypes

import numpy as np

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib

MMA_OP = np.dtype([("w_off", "<u4"), ("n", "<u2"), ("a_slot", "u1"), ("acc", "u1"), ("flags", "u1"), ("wait0", "u1"),
                   ("wait1", "u1"), ("kcount", "u1")])
EPI_JOB = np.dtype([("kind", "u1"), ("acc", "u1"), ("ncols", "u1"), ("out_slot", "u1"), ("ready_bar", "u1"), ("enc", "u1"),
                    ("enc_bar", "u1"), ("flags", "u1"), ("save_slot", "<i2"), ("enc_save_slot", "<i2"), ("mask_slot", "<i2"),
                    ("mask_word0", "<u2"), ("bias_off", "<u2"), ("acc_col", "<u2")])
PACK_CHUNK = np.dtype([("dst_off", "<u4"), ("n_rows", "<i4"), ("src_base", "<i8"), ("row_stride", "<i4"),
                       ("col_stride", "<i4"), ("valid_rows", "<i4"), ("valid_cols", "<i4")])
WGRAD_UNIT = np.dtype([("n_p", "u1"), ("n_q", "u1"), ("pad", "u1", (2,)), ("p_slot", "<i2", (6,)), ("q_slot", "<i2", (4,)),
                       ("sg_slot", "<i2"), ("pad1", "<i2"), ("m_valid", "<i4"), ("n_valid", "<i4"), ("pad2", "<i4"), ("w_base", "<i8"),
                       ("w_row_stride", "<i4"), ("pad3", "<i4"), ("b_base", "<i8"), ("sg_w_base", "<i8"), ("sg_b_base", "<i8")])
PACK_BIAS = np.dtype([("dst_off", "<u4"), ("pad", "<u4"), ("src_base", "<i8"), ("count", "<i4"), ("padded", "<i4")])

NONE = 0xFF
OP_FIRST, OP_COMMIT = 1, 2
(EK_PROLOGUE_FWD, EK_RELU, EK_LINEAR, EK_SIGMA, EK_RGBA, EK_PROLOGUE_BWD, EK_DMASK, EK_DCOPY) = range(8)
ENC_NONE, ENC_X, ENC_D, ENC_DSIGMA = range(4)
NUM_STAGES = 4
BAR_FULL, BAR_EMPTY, BAR_ACC_FULL, BAR_ACC_FREE, BAR_READY = 0, 4, 8, 10, 12
SLOT_E = 4
NUM_SLOTS = 5
JOB_WAIT, JOB_RELEASE = 1, 2


def get_plan(cfg, program):
    lib = _lib.load()
    cap = 4096
    ops = np.zeros(cap, MMA_OP)
    jobs = np.zeros(cap, EPI_JOB)
    chunks = np.zeros(cap, PACK_CHUNK)
    units = np.zeros(64, WGRAD_UNIT)
    n = [ctypes.c_int32(cap), ctypes.c_int32(cap), ctypes.c_int32(cap), ctypes.c_int32(64)]
    info = (ctypes.c_int32 * 16)()
    rc = lib.nerf_debug_plan(ctypes.byref(cfg), program, ops.ctypes.data, ctypes.byref(n[0]), jobs.ctypes.data,
                             ctypes.byref(n[1]), chunks.ctypes.data, ctypes.byref(n[2]), units.ctypes.data, ctypes.byref(n[3]), info)
    if rc != 0:
        raise nb.NerfError(rc, "nerf_debug_plan")
    info = list(info)
    assert info[0] == MMA_OP.itemsize and info[1] == EPI_JOB.itemsize and info[2] == PACK_CHUNK.itemsize
    assert info[3] == WGRAD_UNIT.itemsize, (info[3], WGRAD_UNIT.itemsize)
    biases = np.zeros(64, PACK_BIAS)
    nbias = ctypes.c_int32(64)
    assert lib.nerf_debug_plan_biases(ctypes.byref(cfg), biases.ctypes.data, ctypes.byref(nbias)) == 0
    return dict(ops=ops[:n[0].value], jobs=jobs[:n[1].value], chunks=chunks[:n[2].value], units=units[:n[3].value],
                wpack_bytes=info[4], act_slots=info[5], grad_slots=info[6], mask_slots=info[7], bias_floats=info[8],
                n_params=info[9], biases=biases[:nbias.value])


def chunk_matrix(pc, params):
    """Logical [n_rows, 64] matrix of one packed weight chunk."""
    m = np.zeros((pc["n_rows"], 64), dtype=np.float32)
    vr, vc = int(pc["valid_rows"]), int(pc["valid_cols"])
    if vr and vc:
        r = np.arange(vr)[:, None]
        c = np.arange(vc)[None, :]
        m[:vr, :vc] = params[pc["src_base"] + r * pc["row_stride"] + c * pc["col_stride"]]
    return m


def padded_bias(plan, params):
    b = np.zeros(plan["bias_floats"], dtype=np.float32)
    for e in plan["biases"]:
        b[e["dst_off"]:e["dst_off"] + e["count"]] = params[e["src_base"]:e["src_base"] + e["count"]]
    return b


def emulate_chain(plan, params, posenc_x, posenc_d, d_sigma=None, d_rgba=None, rgba=None, masks=None):
    """Run one tile's program sequentially. posenc_x [128,<=64], posenc_d [128,<=32] are the encoded
    inputs. Returns dict(sigma, rgba, act{slot: panel}, grad{slot: panel}, masks{slot: [128,256] bool})."""
    ops, jobs, chunks = plan["ops"], plan["jobs"], plan["chunks"]
    bias = padded_bias(plan, params)
    wide = int(ops["a_slot"].max()) > SLOT_E      # hidden 257..512: hidden panels 0..7, encoded inputs in slot 8
    SLOT_E_ = 8 if wide else SLOT_E
    slots = np.zeros((9, 128, 64), dtype=np.float32)
    acc = np.zeros((128, 512), dtype=np.float32)   # TMEM columns: two 256-column accumulator sets
    out = dict(sigma=None, rgba=None, saved={}, masks={} if masks is None else masks)

    def pad(a, w):
        r = np.zeros((128, w), dtype=np.float32)
        r[:, :a.shape[1]] = a
        return r

    def run_job(j):
        k = j["kind"]
        if k in (EK_RELU, EK_LINEAR, EK_DMASK, EK_DCOPY):
            nc = int(j["ncols"])
            c0 = int(j["acc_col"])
            v = acc[:, c0:c0 + nc].copy()
            if k in (EK_RELU, EK_LINEAR):
                v = v + bias[j["bias_off"]:j["bias_off"] + nc][None, :]
            if k == EK_RELU:
                if j["mask_slot"] >= 0:
                    m = out["masks"].setdefault(int(j["mask_slot"]), np.zeros((128, 512), dtype=bool))
                    m[:, 32 * j["mask_word0"]:32 * j["mask_word0"] + nc] = ~np.signbit(v)
                v = np.maximum(v, 0)
            if k == EK_DMASK:
                m = out["masks"][int(j["mask_slot"])][:, 32 * j["mask_word0"]:32 * j["mask_word0"] + nc]
                v = v * m
            for p in range(nc // 64):
                slots[j["out_slot"] + p] = v[:, 64 * p:64 * p + 64]
                if j["save_slot"] >= 0:
                    out["saved"][int(j["save_slot"]) + p] = slots[j["out_slot"] + p].copy()
        elif k == EK_SIGMA:
            out["sigma"] = acc[:, int(j["acc_col"])] + bias[j["bias_off"]]
        elif k == EK_RGBA:
            c0 = int(j["acc_col"])
            out["rgba"] = 1.0 / (1.0 + np.exp(-(acc[:, c0:c0 + 4] + bias[j["bias_off"]:j["bias_off"] + 4][None, :])))
        wrote_e = True
        if k == EK_PROLOGUE_FWD or j["enc"] == ENC_X:
            slots[SLOT_E_] = pad(posenc_x, 64)
        elif j["enc"] == ENC_D:
            slots[SLOT_E_] = pad(posenc_d, 64)
        elif j["enc"] == ENC_DSIGMA:
            slots[SLOT_E_] = 0
            slots[SLOT_E_][:, 0] = d_sigma
        elif k == EK_PROLOGUE_BWD:
            slots[SLOT_E_] = 0
            slots[SLOT_E_][:, :4] = d_rgba * rgba * (1 - rgba)
        else:
            wrote_e = False
        if wrote_e and j["enc_save_slot"] >= 0:
            out["saved"][int(j["enc_save_slot"])] = slots[SLOT_E_].copy()

    ji = 0
    while ji < len(jobs) and jobs[ji]["acc"] == NONE:
        run_job(jobs[ji]); ji += 1
    for op, pc in zip(ops, chunks):
        assert op["w_off"] == pc["dst_off"] and op["n"] == pc["n_rows"]
        w = chunk_matrix(pc, params)
        kk = 16 * int(op["kcount"])
        prod = slots[op["a_slot"]][:, :kk] @ w[:, :kk].T
        n = int(op["n"])
        c0 = 256 * int(op["acc"])
        if op["flags"] & OP_FIRST:
            acc[:, c0:c0 + n] = prod
        else:
            acc[:, c0:c0 + n] += prod
        if op["flags"] & OP_COMMIT:
            assert (wide or jobs[ji]["acc"] == op["acc"]) and jobs[ji]["flags"] & JOB_WAIT, "job/commit order mismatch"
            while True:   # all jobs of this GEMM, then any prologue-type jobs
                last = jobs[ji]["flags"] & JOB_RELEASE
                run_job(jobs[ji]); ji += 1
                if last:
                    break
            while ji < len(jobs) and jobs[ji]["acc"] == NONE:
                run_job(jobs[ji]); ji += 1
    assert ji == len(jobs)
    return out


def emulate_wgrad(plan, act, grad, n_params):
    g = np.zeros(n_params, dtype=np.float64)
    for u in plan["units"]:
        P = np.concatenate([act[int(s)] for s in u["p_slot"][:u["n_p"]]], axis=1).astype(np.float64)
        Q = np.concatenate([grad[int(s)] for s in u["q_slot"][:u["n_q"]]], axis=1).astype(np.float64)
        dwt = P.T @ Q  # [in, out]
        mv, nv = int(u["m_valid"]), int(u["n_valid"])
        for n in range(nv):
            base = int(u["w_base"]) + n * int(u["w_row_stride"])
            g[base:base + mv] += dwt[:mv, n]
        if u["b_base"] >= 0:
            g[int(u["b_base"]):int(u["b_base"]) + nv] += Q[:, :nv].sum(0)
        if u["sg_slot"] >= 0:   # sigma row of fc8 from the same P panels (CUDA cores in k_wgrad)
            ds = grad[int(u["sg_slot"])][:, 0].astype(np.float64)
            g[int(u["sg_w_base"]):int(u["sg_w_base"]) + mv] += P[:, :mv].T @ ds
            if u["sg_b_base"] >= 0:
                g[int(u["sg_b_base"])] += ds.sum()
        assert int(u["n_p"]) <= 4 or (int(u["n_q"]) <= 2 and u["sg_slot"] < 0), "5 M-side panels need N <= 128 (TMEM 3 x 128 columns)"
    return g.astype(np.float32)


# ---------------------------------------------------------------------------- protocol simulator
class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        assert self.pending >= n, "mbarrier arrival overflow (1:1 pairing violated)"
        self.pending -= n
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.phase & 1) != parity


def simulate_protocol(plan, n_tiles, rng, max_steps=2_000_000):
    """Random-interleaving simulation of the producer / MMA / epilogue roles against the mbarrier
    protocol of k_chain. Raises AssertionError on deadlock, on a read of a stale or too-new slot
    version, on an accumulator or ring-stage hazard. Returns the number of scheduler steps."""
    ops, jobs = plan["ops"], plan["jobs"]
    bars = [Barrier(1) for _ in range(2 * NUM_STAGES)] + [Barrier(1), Barrier(1), Barrier(4), Barrier(4)] + [Barrier(4) for _ in range(3)]

    # ---- logical (sequential) semantics: expected slot / accumulator versions per op and job
    slot_ver = [0] * NUM_SLOTS
    acc_ver = [0, 0]
    op_expect, job_expect_acc, job_writes = [], [], []
    seq = []  # (tile, 'op'/'job', index)
    for t in range(n_tiles):
        ji = 0

        def logical_job(ji):
            j = jobs[ji]
            writes = []
            if j["out_slot"] != NONE and j["kind"] in (EK_RELU, EK_LINEAR, EK_DMASK, EK_DCOPY):
                for p in range(int(j["ncols"]) // 64):
                    slot_ver[j["out_slot"] + p] += 1
                    writes.append((int(j["out_slot"]) + p, slot_ver[j["out_slot"] + p]))
            if j["kind"] in (EK_PROLOGUE_FWD, EK_PROLOGUE_BWD) or j["enc"] != ENC_NONE:
                slot_ver[SLOT_E] += 1
                writes.append((SLOT_E, slot_ver[SLOT_E]))
            job_writes.append(writes)
            job_expect_acc.append(acc_ver[j["acc"]] if j["acc"] != NONE else None)

        while ji < len(jobs) and jobs[ji]["acc"] == NONE:
            logical_job(ji); ji += 1
        for oi, op in enumerate(ops):
            if op["flags"] & OP_FIRST:
                acc_ver[op["acc"]] += 1
            op_expect.append((slot_ver[op["a_slot"]], acc_ver[op["acc"]]))
            if op["flags"] & OP_COMMIT:
                while True:
                    last = jobs[ji]["flags"] & JOB_RELEASE
                    logical_job(ji); ji += 1
                    if last:
                        break
                while ji < len(jobs) and jobs[ji]["acc"] == NONE:
                    logical_job(ji); ji += 1
        assert ji == len(jobs)

    # ---- concurrent state
    cur_slot_ver = [0] * NUM_SLOTS
    cur_acc_ver = [0, 0]       # version being accumulated / last written
    stage_content = [None] * NUM_STAGES    # global op index whose chunk is in the stage (after load completes)
    n_total_ops = n_tiles * len(ops)
    n_total_jobs = n_tiles * len(jobs)

    prod = dict(i=0, stage=0, phase=0)
    loads = []               # in-flight bulk loads: (stage, global op index)
    mma = dict(i=0, stage=0, phase=0, wph=(1 << BAR_ACC_FREE) | (1 << (BAR_ACC_FREE + 1)), sub=0)
    inflight = []            # issued, incomplete MMA ops (FIFO): dict(gi, commits=[bar ids])
    epi = dict(i=0, aph=0, sub=0)
    steps = 0
    while prod["i"] < n_total_ops or mma["i"] < n_total_ops or epi["i"] < n_total_jobs or inflight or loads:
        steps += 1
        assert steps < max_steps, "simulation did not terminate"
        actions = []
        # producer
        if prod["i"] < n_total_ops and bars[BAR_EMPTY + prod["stage"]].passed(prod["phase"] ^ 1):
            actions.append("prod")
        if loads:
            actions.append("load_done")
        # mma
        if mma["i"] < n_total_ops:
            op = ops[mma["i"] % len(ops)]
            ok = True
            if mma["sub"] == 0:
                for w in (op["wait0"], op["wait1"]):
                    if w != NONE and not bars[w].passed((mma["wph"] >> int(w)) & 1):
                        ok = False
                if ok and not bars[BAR_FULL + mma["stage"]].passed(mma["phase"]):
                    ok = False
            if ok:
                actions.append("mma")
        if inflight:
            actions.append("mma_done")
        # epilogue
        if epi["i"] < n_total_jobs:
            j = jobs[epi["i"] % len(jobs)]
            if epi["sub"] == 0 and j["acc"] != NONE and (j["flags"] & JOB_WAIT):
                if bars[BAR_ACC_FULL + j["acc"]].passed((epi["aph"] >> int(j["acc"])) & 1):
                    actions.append("epi")
            else:
                actions.append("epi")
        assert actions, f"deadlock: prod={prod} mma={mma} epi={epi} inflight={len(inflight)}"
        a = actions[rng.integers(len(actions))]
        if a == "prod":
            st = prod["stage"]
            # the stage may only be refilled once the op that read it has completed
            assert stage_content[st] is None or stage_content[st] == "free", "ring stage overwritten while in use"
            loads.append((st, prod["i"]))
            stage_content[st] = "loading"
            prod["i"] += 1
            prod["stage"] += 1
            if prod["stage"] == NUM_STAGES:
                prod["stage"], prod["phase"] = 0, prod["phase"] ^ 1
        elif a == "load_done":
            st, gi = loads.pop(rng.integers(len(loads)))
            stage_content[st] = gi
            bars[BAR_FULL + st].arrive()
        elif a == "mma":
            gi = mma["i"]
            op = ops[gi % len(ops)]
            for w in (op["wait0"], op["wait1"]):
                if w != NONE:
                    mma["wph"] ^= 1 << int(w)
            st = mma["stage"]
            assert stage_content[st] == gi, f"op {gi} found chunk of op {stage_content[st]} in its ring stage"
            exp_slot, exp_acc = op_expect[gi]
            assert cur_slot_ver[op["a_slot"]] == exp_slot, f"op {gi}: A slot {op['a_slot']} version {cur_slot_ver[op['a_slot']]} != expected {exp_slot}"
            if op["flags"] & OP_FIRST:
                cur_acc_ver[op["acc"]] += 1
            assert cur_acc_ver[op["acc"]] == exp_acc, f"op {gi}: accumulator version mismatch"
            commits = [BAR_EMPTY + st]
            if op["flags"] & OP_COMMIT:
                commits.append(BAR_ACC_FULL + int(op["acc"]))
            inflight.append(dict(gi=gi, commits=commits, slot=int(op["a_slot"]), slot_ver=exp_slot, stage=st, acc=int(op["acc"]), acc_ver=exp_acc))
            mma["i"] += 1
            mma["stage"] += 1
            if mma["stage"] == NUM_STAGES:
                mma["stage"], mma["phase"] = 0, mma["phase"] ^ 1
        elif a == "mma_done":
            o = inflight.pop(0)   # tensor-core ops complete in issue order
            assert cur_slot_ver[o["slot"]] == o["slot_ver"], f"slot {o['slot']} overwritten while op {o['gi']} was reading it"
            assert cur_acc_ver[o["acc"]] == o["acc_ver"], "accumulator overwritten while an op was accumulating into it"
            stage_content[o["stage"]] = "free"
            for b in o["commits"]:
                bars[b].arrive()
        elif a == "epi":
            gj = epi["i"]
            j = jobs[gj % len(jobs)]
            if epi["sub"] == 0:
                if j["acc"] != NONE:
                    if j["flags"] & JOB_WAIT:
                        epi["aph"] ^= 1 << int(j["acc"])
                    assert cur_acc_ver[j["acc"]] == job_expect_acc[gj], f"job {gj}: accumulator version mismatch"
                    assert not any(o["acc"] == j["acc"] and o["acc_ver"] == job_expect_acc[gj] for o in inflight), "job reads an accumulator with MMAs in flight"
                    epi["sub"] = 1
                else:
                    epi["sub"] = 2
            elif epi["sub"] == 1:   # accumulator read: release it
                assert cur_acc_ver[j["acc"]] == job_expect_acc[gj], "accumulator overwritten while the epilogue was reading it"
                if j["flags"] & JOB_RELEASE:
                    bars[BAR_ACC_FREE + j["acc"]].arrive(4)
                epi["sub"] = 2
            elif epi["sub"] == 2:   # panel writes
                for slot, ver in job_writes[gj]:
                    assert not any(o["slot"] == slot for o in inflight), f"job {gj} writes slot {slot} while an MMA reads it"
                    assert cur_slot_ver[slot] == ver - 1, "slot written out of order"
                    cur_slot_ver[slot] = ver
                if j["ready_bar"] != NONE:
                    bars[j["ready_bar"]].arrive(4)
                if j["enc_bar"] != NONE:
                    bars[j["enc_bar"]].arrive(4)
                epi["sub"] = 0
                epi["i"] += 1
    return steps
