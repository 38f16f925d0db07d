"""TS-mode chain programs (mlp_tc3.cu): fetch the host-built tables and (1) emulate one tile in float32 numpy,
(2) simulate the kernel's roles -- weight producer, the two lanes' MMA issuers, the tensor pipe, the epilogue,
the saver warps -- under random interleavings against the barrier protocol.

The tables are the ones the CUDA kernel receives as parameters (TsOp / TsStep / PackChunk, mlp_tc.h).
"""
import ctypes

import numpy as np

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib
from tests.tc_plan_util import (PACK_CHUNK, chunk_matrix, padded_bias, EK_PROLOGUE_FWD, EK_RELU, EK_LINEAR, EK_SIGMA, EK_RGBA,
                                EK_PROLOGUE_BWD, EK_DMASK, EK_DCOPY, ENC_NONE, ENC_X, ENC_D, ENC_DSIGMA, Barrier)

TS_OP = np.dtype([("w_off", "<u4"), ("n", "u1"), ("a_src", "u1"), ("kcount", "u1"), ("first", "u1")])
TS_STEP = np.dtype([("op_begin", "<u2"), ("op_end", "<u2"), ("kind", "u1"), ("enc", "u1"), ("ncols", "u1"), ("a_col", "u1"),
                    ("final_step", "u1"), ("writes_a", "u1"), ("mask_word0", "u1"), ("pre_enc", "u1"), ("bias_off", "<u2"),
                    ("save_slot", "<i2"), ("enc_save_slot", "<i2"), ("mask_slot", "<i2"), ("b_enc", "u1"), ("pad", "u1"),
                    ("b_save_slot", "<i2")])
A_SMEM, A_SMEM_B = 0xFF, 0xFE
PRE_NONE, PRE_A, PRE_B = 0, 1, 2
STAGES = 16


def get_ts_plan(cfg, program):
    lib = _lib.load()
    cap = 512
    ops, steps, chunks = np.zeros(cap, TS_OP), np.zeros(cap, TS_STEP), np.zeros(cap, PACK_CHUNK)
    n = [ctypes.c_int32(cap) for _ in range(3)]
    info = (ctypes.c_int32 * 8)()
    rc = lib.nerf_debug_ts_plan(ctypes.byref(cfg), program, ops.ctypes.data, ctypes.byref(n[0]), steps.ctypes.data, ctypes.byref(n[1]),
                                chunks.ctypes.data, ctypes.byref(n[2]), info)
    if rc != 0:
        raise nb.NerfError(rc, "nerf_debug_ts_plan")
    assert (info[0], info[1], info[2]) == (TS_OP.itemsize, TS_STEP.itemsize, PACK_CHUNK.itemsize), list(info)[:3]
    return dict(ops=ops[:n[0].value], steps=steps[:n[1].value], chunks=chunks[:n[2].value], wpack_bytes=int(info[3]))


def _panel(a, w=64):
    r = np.zeros((128, w), np.float32)
    r[:, :a.shape[1]] = a
    return r


def write_panel(pj, panel_b, x_enc, d_enc, d_sigma, d_rgba, rgba):
    """What write_enc puts into slot E_A (panel_b False) / E_B of a tile (mlp_tc3.cu write_panel)."""
    kind, enc = (EK_RELU, int(pj["b_enc"])) if panel_b else (int(pj["kind"]), int(pj["enc"]))
    if kind == EK_PROLOGUE_FWD or enc == ENC_X:
        return _panel(x_enc)
    if enc == ENC_D:
        return _panel(d_enc)
    p = np.zeros((128, 64), np.float32)
    if enc == ENC_DSIGMA:
        p[:, 0] = d_sigma
    elif kind == EK_PROLOGUE_BWD:
        p[:, :4] = d_rgba * rgba * (1 - rgba)
    return p


def emulate_ts_chain(ts, bias, params, x_enc, d_enc, d_sigma=None, d_rgba=None, rgba=None, masks=None):
    """One tile of a TS program in float32. Returns dict(sigma, rgba, saved{slot: [128,64]}, masks{slot: [128,256] bool})."""
    ops, steps, chunks = ts["ops"], ts["steps"], ts["chunks"]
    pj = steps[0]
    out = dict(sigma=None, rgba=None, saved={}, masks={} if masks is None else masks)
    e = [write_panel(pj, False, x_enc, d_enc, d_sigma, d_rgba, rgba), None]
    if pj["enc_save_slot"] >= 0:
        out["saved"][int(pj["enc_save_slot"])] = e[0].copy()
    if pj["b_enc"] != ENC_NONE:
        e[1] = write_panel(pj, True, x_enc, d_enc, d_sigma, d_rgba, rgba)
        if pj["b_save_slot"] >= 0:
            out["saved"][int(pj["b_save_slot"])] = e[1].copy()
    act = np.zeros((128, 256), np.float32)      # the lane's activation columns (256 bf16 features)
    stash = None
    for st in steps[1:]:
        acc = None
        for oi in range(int(st["op_begin"]), int(st["op_end"])):
            op, pc = ops[oi], chunks[oi]
            assert op["w_off"] == pc["dst_off"] and op["n"] == pc["n_rows"] == st["ncols"]
            assert bool(op["first"]) == (oi == st["op_begin"])
            w = chunk_matrix(pc, params)
            kk = 16 * int(op["kcount"])
            if op["a_src"] == A_SMEM:
                a = e[0]
            elif op["a_src"] == A_SMEM_B:
                a = e[1]
            else:
                assert op["kcount"] == 4
                a = act[:, 64 * int(op["a_src"]):64 * int(op["a_src"]) + 64]
            prod = a[:, :kk] @ w[:, :kk].T
            acc = prod if op["first"] else acc + prod
        k, nc = int(st["kind"]), int(st["ncols"])
        if k == EK_SIGMA:
            out["sigma"] = acc[:, 0] + bias[st["bias_off"]]
            continue
        if k == EK_RGBA:
            out["rgba"] = 1.0 / (1.0 + np.exp(-(acc[:, :4] + bias[st["bias_off"]:st["bias_off"] + 4][None, :])))
            continue
        v = acc
        if k in (EK_RELU, EK_LINEAR):
            v = v + bias[st["bias_off"]:st["bias_off"] + nc][None, :]
        if k == EK_RELU:
            v = np.maximum(v, 0)
            if st["mask_slot"] >= 0:   # (the saver warps derive the mask from the stored activation: != 0)
                m = out["masks"].setdefault(int(st["mask_slot"]), np.zeros((128, 256), bool))
                m[:, 32 * int(st["mask_word0"]):32 * int(st["mask_word0"]) + nc] = v != 0
        if k == EK_DMASK:
            v = v * out["masks"][int(st["mask_slot"])][:, 32 * int(st["mask_word0"]):32 * int(st["mask_word0"]) + nc]
        if st["save_slot"] >= 0:
            for p in range(nc // 64):
                out["saved"][int(st["save_slot"]) + p] = v[:, 64 * p:64 * p + 64].copy()
        assert st["writes_a"], "every non-head step hands its output on through the activation columns"
        f0 = 2 * int(st["a_col"])                # first feature of this step's output
        if not st["final_step"]:
            assert stash is None and f0 == 0
            stash = (f0, v)                      # waits in registers until the layer's MMAs are done
        else:
            if stash is not None:
                act[:, stash[0]:stash[0] + stash[1].shape[1]] = stash[1]
                stash = None
            act[:, f0:f0 + nc] = v
    assert stash is None
    return out


# ---------------------------------------------------------------------------------------------- protocol simulation
def simulate_ts_protocol(ts, n_pairs, rng, train, max_steps=4_000_000, broken=None):
    """Random-interleaving simulation of one CTA pair running `n_pairs` pair tiles on two lanes (mlp_tc3.cu):
    weight producer (16-stage ring, one FULL barrier per step, EMPTY counts both lanes), two MMA issuers serialised
    by a lock, an in-order tensor pipe, the epilogue (accumulator hand-back, first-half stash, in-place activation
    writes, slot-E panels of the NEXT tile written by the marked steps) and -- when training -- the saver warps.
    Asserts: no deadlock; every MMA reads the activation / slot-E / weight version the sequential semantics say;
    nothing is overwritten while an MMA in flight (or a saver) still reads it; no barrier runs two phases ahead.
    `broken` switches on one deliberate protocol bug (the tests check that each is caught):
      "early_final_signal"  a layer's final step releases the lane before its activations are written
      "no_act_saved_wait"   the epilogue overwrites activations without waiting for the saver warps
      "empty_count_1"       a weight stage is released when ONE lane's MMAs on it are done"""
    ops, steps = ts["ops"], ts["steps"][1:]
    pj = ts["steps"][0]
    has_b = pj["b_enc"] != ENC_NONE
    n_steps = len(steps)
    lanes = [ln for ln in range(2) if ln < n_pairs]
    tiles = {ln: list(range(ln, n_pairs, 2)) for ln in lanes}            # lane ln runs pair tiles ln, ln + 2, ...
    n_groups = len(tiles[0]) * n_steps                                   # (lane 1 never outlives lane 0)

    def group(gi):          # -> step index, tile index (per lane position), lanes alive
        t, p = divmod(gi, n_steps)
        return p, t, [ln for ln in lanes if t < len(tiles[ln])]

    # ---- sequential semantics: version of (lane, object) each op / step expects
    # objects: 'A' activations (version = number of final writes), 'EA', 'EB' (version = tile count), 'acc' (version per step)
    stage_of, cum = [], 0
    for gi in range(n_groups):
        p = gi % n_steps
        stage_of.append(cum)
        cum += int(steps[p]["op_end"]) - int(steps[p]["op_begin"])
    full = [Barrier(1) for _ in range(16)]
    empty = [Barrier(1 if broken == "empty_count_1" else 2) for _ in range(STAGES)]
    acc_full = [Barrier(1), Barrier(1)]
    epi_done = [Barrier(1), Barrier(1)]
    act_ready = [Barrier(1), Barrier(1)]
    act_saved = [Barrier(1), Barrier(1)]
    ver = {(ln, o): 0 for ln in lanes for o in ("A", "EA", "EB", "acc")}
    readers = {(ln, o): 0 for ln in lanes for o in ("A", "EA", "EB")}   # MMAs in flight / saver reading the object
    stage_owner = [None] * STAGES        # group index whose chunk the stage holds (None = free)
    stage_users = [0] * STAGES
    lock = [None]

    prod = dict(gi=0, oi=0, armed=False)
    loads = []
    iss = {ln: dict(gi=0, state=0, done_phase=0) for ln in lanes}
    pipe = []                 # FIFO of issued MMA groups: dict(ln, gi, stages, reads)
    epi = dict(gi=0, lane_i=0, sub=0, aph=[0, 0], sph=[0, 0], a_writes={ln: 0 for ln in lanes}, second_pass=False)
    sav = dict(gi=0, lane_i=0, rph=[0, 0], sub=0) if train else None
    # expected versions (sequential): activations are rewritten once per final step, E panels once per tile
    def expected(ln, gi, obj):
        p, t, _ = group(gi)
        if obj == "A":
            return t * sum(int(s["final_step"]) and int(s["writes_a"]) for s in steps) + sum(
                int(s["final_step"]) and int(s["writes_a"]) for s in steps[:p])
        return t + 1          # slot-E panels of tile t: version t + 1 (the first tile's are written up front)

    # prologue: both lanes' first panels
    for ln in lanes:
        ver[(ln, "EA")] = 1
        if has_b:
            ver[(ln, "EB")] = 1
        epi_done[ln].arrive()

    steps_done = 0
    while True:
        work_left = (epi["gi"] < n_groups or pipe or loads or prod["gi"] < n_groups or any(i["gi"] < n_groups for i in iss.values())
                     or (sav is not None and sav["gi"] < n_groups))
        if not work_left:
            break
        steps_done += 1
        assert steps_done < max_steps, "simulation did not terminate"
        acts = []
        # producer: next chunk of group prod.gi
        if prod["gi"] < n_groups:
            p = prod["gi"] % n_steps
            n_ops = int(steps[p]["op_end"]) - int(steps[p]["op_begin"])
            stg = (stage_of[prod["gi"]] + prod["oi"]) % STAGES
            use = (stage_of[prod["gi"]] + prod["oi"]) // STAGES
            if empty[stg].passed((use & 1) ^ 1):
                acts.append("prod")
        if loads:
            acts.append("load_done")
        for ln in lanes:
            it = iss[ln]
            if it["gi"] >= n_groups:
                continue
            p, t, alive = group(it["gi"])
            if ln not in alive:
                acts.append(("skip", ln))
            elif it["state"] == 0 and full[it["gi"] % 16].passed((it["gi"] // 16) & 1):
                acts.append(("wfull", ln))
            elif it["state"] == 1 and epi_done[ln].passed(it["done_phase"]):
                acts.append(("wepi", ln))
            elif it["state"] == 2 and (lock[0] is None or len(alive) == 1):
                acts.append(("issue", ln))
        if pipe:
            acts.append("mma_done")
        if epi["gi"] < n_groups:
            p, t, alive = group(epi["gi"])
            ln = alive[epi["lane_i"]]
            if epi["sub"] == 0:
                if acc_full[ln].passed(epi["aph"][ln]):
                    acts.append("epi")
            elif epi["sub"] == 2 and train and steps[p]["save_slot"] >= 0 and broken != "no_act_saved_wait":
                if act_saved[ln].passed(epi["sph"][ln] ^ 1):
                    acts.append("epi")
            else:
                acts.append("epi")
        if sav is not None and sav["gi"] < n_groups:
            p, t, alive = group(sav["gi"])
            st = steps[p]
            if not (st["final_step"] and st["save_slot"] >= 0):
                acts.append("sav_skip")
            else:
                ln = alive[sav["lane_i"]]
                if sav["sub"] == 1 or act_ready[ln].passed(sav["rph"][ln]):
                    acts.append("sav")
        assert acts, f"deadlock: prod={prod} iss={iss} epi={epi} sav={sav} pipe={len(pipe)}"
        a = acts[rng.integers(len(acts))]

        if a == "prod":
            gi, oi = prod["gi"], prod["oi"]
            p = gi % n_steps
            n_ops = int(steps[p]["op_end"]) - int(steps[p]["op_begin"])
            stg = (stage_of[gi] + oi) % STAGES
            assert stage_owner[stg] is None and stage_users[stg] == 0, "ring stage overwritten while in use"
            stage_owner[stg] = "loading"
            loads.append((stg, gi, oi == n_ops - 1))
            prod["oi"] += 1
            if prod["oi"] == n_ops:
                prod["gi"], prod["oi"] = gi + 1, 0
        elif a == "load_done":
            # chunks of one step all complete on the step's FULL barrier (bytes): model = barrier arrives when the LAST lands
            idx = rng.integers(len(loads))
            stg, gi, _ = loads.pop(idx)
            stage_owner[stg] = gi
            if not any(l[1] == gi for l in loads) and (prod["gi"] > gi):
                full[gi % 16].arrive()
        elif isinstance(a, tuple) and a[0] == "skip":
            iss[a[1]]["gi"] += 1
        elif isinstance(a, tuple) and a[0] == "wfull":
            iss[a[1]]["state"] = 1
        elif isinstance(a, tuple) and a[0] == "wepi":
            it = iss[a[1]]
            it["done_phase"] ^= 1
            it["state"] = 2
        elif isinstance(a, tuple) and a[0] == "issue":
            ln = a[1]
            it = iss[ln]
            gi = it["gi"]
            p, t, alive = group(gi)
            st = steps[p]
            n_ops = int(st["op_end"]) - int(st["op_begin"])
            stages = [(stage_of[gi] + i) % STAGES for i in range(n_ops)]
            reads = set()
            for i, oi in enumerate(range(int(st["op_begin"]), int(st["op_end"]))):
                assert stage_owner[stages[i]] == gi, f"group {gi} op {i}: stage holds {stage_owner[stages[i]]}"
                src = int(ops[oi]["a_src"])
                obj = "EA" if src == A_SMEM else ("EB" if src == A_SMEM_B else "A")
                want = expected(ln, gi, obj)
                assert ver[(ln, obj)] == want, f"lane {ln} group {gi} (step {p}) reads {obj} v{ver[(ln, obj)]}, expected v{want}"
                reads.add(obj)
            for o in reads:
                readers[(ln, o)] += 1
            for s_ in stages:
                stage_users[s_] += 1
            ver[(ln, "acc")] += 1
            pipe.append(dict(ln=ln, gi=gi, stages=stages, reads=reads, lone=len(alive) == 1, acc_ver=ver[(ln, "acc")]))
            it["gi"], it["state"] = gi + 1, 0
        elif a == "mma_done":
            o = pipe.pop(0)
            for obj in o["reads"]:
                readers[(o["ln"], obj)] -= 1
            for s_ in o["stages"]:
                stage_users[s_] -= 1
                if broken == "empty_count_1":
                    if empty[s_].pending == empty[s_].count and stage_users[s_] == 0 and stage_owner[s_] is None:
                        continue           # (second lane arriving on an already released stage)
                    empty[s_].arrive(1)
                    stage_owner[s_] = None if True else stage_owner[s_]
                    continue
                empty[s_].arrive(2 if o["lone"] else 1)
                if empty[s_].pending == empty[s_].count:      # phase just completed: both lanes are done with the stage
                    stage_owner[s_] = None
            assert ver[(o["ln"], "acc")] == o["acc_ver"], "accumulator overwritten under an MMA"
            acc_full[o["ln"]].arrive()
        elif a == "epi":
            p, t, alive = group(epi["gi"])
            ln = alive[epi["lane_i"]]
            st = steps[p]
            last = epi["gi"] == n_groups - 1 or (p == n_steps - 1 and t + 1 >= len(tiles[ln]))
            small = st["kind"] in (EK_SIGMA, EK_RGBA)
            if epi["sub"] == 0:        # accumulator in registers
                epi["aph"][ln] ^= 1
                assert not any(o["ln"] == ln for o in pipe if o["acc_ver"] == ver[(ln, "acc")] and False)
                early_bug = broken == "early_final_signal" and st["final_step"] and not small
                if (not st["final_step"] or early_bug) and not last:
                    epi_done[ln].arrive()            # early signal
                epi["sub"] = 1 if (st["final_step"] and not small) else 3
                if epi["sub"] == 1:
                    epi["sub"] = 2
            elif epi["sub"] == 2:      # final step: activations written in place (after ACT_SAVED when training)
                if train and st["save_slot"] >= 0:
                    epi["sph"][ln] ^= 1
                assert readers[(ln, "A")] == 0, f"lane {ln}: activations overwritten while MMAs / savers still read them"
                ver[(ln, "A")] += 1
                if not last and broken != "early_final_signal":
                    epi_done[ln].arrive()
                if train and st["save_slot"] >= 0:
                    act_ready[ln].arrive()
                epi["sub"] = 3
            else:                      # after the signal: slot-E panel of the lane's next tile, then the next item
                if st["final_step"] and small and not last:
                    pass
                if st["pre_enc"] != PRE_NONE and t + 1 < len(tiles[ln]):
                    obj = "EB" if st["pre_enc"] == PRE_B else "EA"
                    assert readers[(ln, obj)] == 0, f"lane {ln}: slot {obj} rewritten while an MMA reads it"
                    assert ver[(ln, obj)] == t + 1, f"lane {ln}: slot {obj} rewritten twice for one tile"
                    # every reader of this tile's panel must already have been ISSUED (and, checked above, completed)
                    src = A_SMEM_B if obj == "EB" else A_SMEM
                    last_reader = max(p_ for p_ in range(n_steps)
                                      if any(int(ops[oi]["a_src"]) == src for oi in range(int(steps[p_]["op_begin"]), int(steps[p_]["op_end"]))))
                    assert iss[ln]["gi"] > t * n_steps + last_reader, f"lane {ln}: slot {obj} rewritten before its last reader of tile {t} was issued"
                    ver[(ln, obj)] += 1
                if st["final_step"] and small and not last:
                    epi_done[ln].arrive()
                epi["sub"] = 0
                epi["lane_i"] += 1
                if epi["lane_i"] == len(alive):
                    epi["lane_i"] = 0
                    epi["gi"] += 1
        elif a == "sav_skip":
            sav["gi"] += 1
        elif a == "sav":
            p, t, alive = group(sav["gi"])
            ln = alive[sav["lane_i"]]
            if sav["sub"] == 0:        # ACT_READY seen: reading the activations
                sav["rph"][ln] ^= 1
                readers[(ln, "A")] += 1
                sav["sub"] = 1
            else:
                readers[(ln, "A")] -= 1
                act_saved[ln].arrive()
                sav["sub"] = 0
                sav["lane_i"] += 1
                if sav["lane_i"] == len(alive):
                    sav["lane_i"] = 0
                    sav["gi"] += 1
    return steps_done
