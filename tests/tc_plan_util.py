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
