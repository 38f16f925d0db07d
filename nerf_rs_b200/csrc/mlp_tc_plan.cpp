// mlp_tc_plan.cpp -- host-side builder of the per-tile programs the fused MLP kernels run.
//
// Network: DensityNet fc1..fc8 and RadianceNet fc9/fc10 (src/model.rs:44-67, 85-131) with the
// north-star options (posenc input, skip into fc6, direction into fc9). The builder emits, for
// one 128-sample tile:
//   * MMA ops   -- one per (GEMM, N block <=128, 64-wide K panel), in issue order, each naming the
//                  smem slot of its A operand, the packed weight chunk it consumes from the ring,
//                  and the mbarriers the issuing thread must wait on first;
//   * epilogue jobs -- one per accumulator block, in order;
//   * pack chunks   -- how to gather each weight chunk from the flat [out,in] f32 blob.
// Every GEMM issues full-width MMAs (N = layer width, <= 256) into one of two TMEM accumulator sets;
// consecutive GEMMs alternate sets, so a layer's epilogue (reading set s, rewriting the hidden
// panels in place) overlaps the next layer's MMAs (writing set s^1) panel pair by panel pair.
// tests/test_tc_plan.py simulates the three roles against these tables to prove the schedule
// is deadlock-free and hazard-free for every supported geometry.
#include <cstdio>
#include <cstring>

#include "mlp_tc.h"

namespace {

struct Builder {
    const NetGeom &g;
    TcProgram &prog;
    int np, np2;
    int set = 0;                  // accumulator set the next GEMM uses
    bool pending[3] = {false, false, false};  // unconsumed "ready" event per slot group
    std::string err;

    bool wide = false;            // hidden 257..512: 8 hidden panels, slot E = 8, N = 512 GEMMs issued as two 256-column halves
    int e_slot = TC_SLOT_E;

    Builder(const NetGeom &g_, TcProgram &p) : g(g_), prog(p) {
        np = g.Wp / 64;
        np2 = g.W2p / 64;
        wide = np > 4;
        e_slot = wide ? TC_SLOT_E_WIDE : TC_SLOT_E;
    }
    int group_of_slot(int slot) const { return slot == e_slot ? 2 : slot / 2; }

    // K-panel input descriptor for one GEMM
    struct KIn {
        int slot, kcount;
        // weight source for rows (n index) r and columns c of this K panel:
        //   element = src_base + r * row_stride + c * col_stride
        int64_t src_base;
        int row_stride, col_stride;
        int valid_cols;
    };

    void signal(int slot, EpiJob &j, bool enc) {
        if (wide) return;   // the per-slot-group ready events are a v1 (one tile per CTA) mechanism; wide runs on the pair kernel only
        const int grp = group_of_slot(slot);
        if (pending[grp]) err = "schedule error: ready event not consumed before the next one";
        pending[grp] = true;
        if (enc) j.enc_bar = (uint8_t)(TC_BAR_READY + grp);
        else j.ready_bar = (uint8_t)(TC_BAR_READY + grp);
    }

    // Emit the MMA ops of one GEMM (or of one 256-column half of a wide GEMM): one op per 64-wide K panel, N = n_mma,
    // into accumulator set `acc`. row0 = first weight row (output index) of this half.
    void emit_ops(const std::vector<KIn> &kin, int acc, int n_mma, int valid_rows, int row0 = 0, bool gemm_start = true,
                  bool commit_last = true) {
        for (size_t i = 0; i < kin.size(); ++i) {
            const KIn &k = kin[i];
            MmaOp op;
            memset(&op, 0, sizeof(op));
            op.w_off = prog.wpack_bytes;
            op.n = (uint16_t)n_mma;
            op.a_slot = (uint8_t)k.slot;
            op.acc = (uint8_t)acc;
            op.kcount = (uint8_t)k.kcount;
            op.flags = 0;
            op.wait0 = op.wait1 = TC_NONE;
            if (i == 0) {
                op.flags |= TC_OP_FIRST;
                if (gemm_start) {
                    op.flags |= TC_OP_GEMM_START;
                    op.wait0 = (uint8_t)(TC_BAR_ACC_FREE + acc);
                }
            }
            if (i + 1 == kin.size() && commit_last) op.flags |= TC_OP_COMMIT_ACC;
            if (!wide) {
                const int grp = group_of_slot(k.slot);
                if (pending[grp]) {
                    op.wait1 = (uint8_t)(TC_BAR_READY + grp);
                    pending[grp] = false;
                }
            }
            prog.ops.push_back(op);
            PackChunk pc;
            pc.dst_off = prog.wpack_bytes;
            pc.n_rows = n_mma;
            pc.src_base = k.src_base + (int64_t)row0 * k.row_stride;
            pc.row_stride = k.row_stride;
            pc.col_stride = k.col_stride;
            pc.valid_rows = valid_rows < 0 ? 0 : (valid_rows > n_mma ? n_mma : valid_rows);
            pc.valid_cols = k.valid_cols;
            prog.chunks.push_back(pc);
            prog.wpack_bytes += (uint32_t)n_mma * 128u;
        }
    }

    // Hidden-panel K inputs: panels [0, n_panels) of the current activation (slots 0..), weight columns
    // start at col0 (in units of the logical matrix's K index), `valid` K entries in total.
    void hidden_kin(std::vector<KIn> &kin, int n_panels, int64_t src_base, int row_stride, int col_stride, int col0,
                    int valid) {
        for (int p = 0; p < n_panels; ++p) {
            KIn k;
            k.slot = p;
            k.kcount = 4;
            k.src_base = src_base + (int64_t)(col0 + 64 * p) * col_stride;
            k.row_stride = row_stride;
            k.col_stride = col_stride;
            int v = valid - 64 * p;
            k.valid_cols = v < 0 ? 0 : (v > 64 ? 64 : v);
            kin.push_back(k);
        }
    }

    // A GEMM whose output is a hidden activation of n_out_panels panels, written IN PLACE over the
    // panels it read (safe: its epilogue starts only after all of its MMAs have completed, and the next
    // GEMM accumulates into the other TMEM set). The epilogue runs as one job per 128-column block so
    // that the next GEMM can start on panels 0,1 while block 1 is still being converted.
    void emit_hidden_gemm(const std::vector<KIn> &kin, int n_out_panels, int n_valid, uint8_t kind, uint32_t bias_off,
                          int save_slot0, int mask_slot, uint8_t enc, int enc_save_slot, bool consumed) {
        const int nblocks = (n_out_panels + 1) / 2;
        int acc = set;
        if (n_out_panels > 4) {
            // wide: two 256-column halves sharing the K inputs; the accumulator is TMEM columns [0, 512)
            acc = 0;
            emit_ops(kin, 0, 256, n_valid, 0, true, false);
            emit_ops(kin, 1, 256, n_valid - 256, 256, false, true);
        } else {
            if (wide) acc = 0;
            else set ^= 1;
            emit_ops(kin, acc, 64 * n_out_panels, n_valid);
        }
        for (int b = 0; b < nblocks; ++b) {
            const int p0 = 2 * b;
            const int pn = (n_out_panels - p0) > 2 ? 2 : (n_out_panels - p0);
            EpiJob j;
            memset(&j, 0, sizeof(j));
            j.kind = kind;
            j.acc = (uint8_t)acc;
            j.ncols = (uint8_t)(64 * pn);
            j.out_slot = (uint8_t)p0;
            j.ready_bar = TC_NONE;
            j.enc = ENC_NONE;
            j.enc_bar = TC_NONE;
            j.flags = (uint8_t)((b == 0 ? TC_JOB_WAIT_ACC : 0u) | (b == nblocks - 1 ? TC_JOB_RELEASE_ACC : 0u));
            j.save_slot = (int16_t)(save_slot0 < 0 ? -1 : save_slot0 + p0);
            j.enc_save_slot = -1;
            j.mask_slot = (int16_t)mask_slot;
            j.mask_word0 = (uint16_t)(4 * b);
            j.bias_off = (uint16_t)(bias_off + 128u * b);
            j.acc_col = (uint16_t)(acc * 256 + 128 * b);   // (wide: acc == 0, blocks 0..3 span both halves)
            if (consumed) signal(j.out_slot, j, false);
            if (b == 0 && enc != ENC_NONE) {
                j.enc = enc;
                j.enc_save_slot = (int16_t)enc_save_slot;
                signal(e_slot, j, true);
            }
            prog.jobs.push_back(j);
        }
    }

    void emit_small_gemm(const std::vector<KIn> &kin, int valid_rows, uint8_t kind, uint32_t bias_off) {
        const int acc = wide ? 0 : set;
        if (!wide) set ^= 1;
        emit_ops(kin, acc, 16, valid_rows);
        EpiJob j;
        memset(&j, 0, sizeof(j));
        j.kind = kind;
        j.acc = (uint8_t)acc;
        j.ncols = 16;
        j.out_slot = TC_NONE;
        j.ready_bar = TC_NONE;
        j.enc = ENC_NONE;
        j.enc_bar = TC_NONE;
        j.flags = TC_JOB_WAIT_ACC | TC_JOB_RELEASE_ACC;
        j.save_slot = j.enc_save_slot = j.mask_slot = -1;
        j.bias_off = (uint16_t)bias_off;
        j.acc_col = (uint16_t)(acc * 256);
        prog.jobs.push_back(j);
    }

    void emit_prologue(uint8_t kind, uint8_t enc, int enc_save_slot) {
        EpiJob j;
        memset(&j, 0, sizeof(j));
        j.kind = kind;
        j.acc = TC_NONE;
        j.out_slot = TC_NONE;
        j.ready_bar = TC_NONE;
        j.enc = enc;
        j.enc_bar = TC_NONE;
        j.save_slot = -1;
        j.mask_slot = -1;
        j.enc_save_slot = (int16_t)enc_save_slot;
        signal(e_slot, j, true);
        prog.jobs.push_back(j);
    }
};

struct Slots {
    int np, np2;
    // activation area
    int X() const { return 0; }
    int D() const { return 1; }
    int H(int l) const { return 2 + (l - 1) * np; }      // l = 1..7
    int feat() const { return 2 + 7 * np; }
    int h9() const { return 2 + 8 * np; }
    int act_slots() const { return 2 + 8 * np + np2; }
    // gradient area
    int R() const { return 0; }
    int Sg() const { return 1; }
    int dP9() const { return 2; }
    int dFeat() const { return 2 + np2; }
    int dP(int l) const { return 2 + np2 + np + (7 - l) * np; }  // l = 7..1
    int grad_slots() const { return 2 + np2 + 8 * np; }
};

}  // namespace

bool tc_build_plan(const NetGeom &g, TcPlan &plan, std::string &err) {
    const int npanels = g.Wp / 64;
    if (g.Wp % 64 || (npanels != 1 && npanels != 2 && npanels != 4 && npanels != 8) || g.W2p % 64 || g.W2p < 64 || g.W2p > 256 ||
        g.Cx > 64 || g.Cd > 32) {
        err = "tcgen05 MLP supports hidden <= 512 (padded to 64, 128, 256 or 512), xyz_freqs <= 10, dir_freqs <= 4";
        return false;
    }
    if (g.skip_layer && (g.skip_layer < 1 || g.skip_layer > 6)) {
        err = "skip_layer must be in 1..6";
        return false;
    }
    const int np = g.Wp / 64, np2 = g.W2p / 64;
    Slots sl{np, np2};
    plan = TcPlan();
    plan.np = np;
    plan.np2 = np2;
    plan.act_slots = sl.act_slots();
    plan.grad_slots = sl.grad_slots();
    plan.mask_slots = 8;

    // ---- padded biases (forward jobs index into this array)
    uint32_t boff = 0;
    uint32_t bias_l[8], bias_s = 0, bias_f = 0, bias_9 = 0, bias_10 = 0;
    for (int l = 1; l <= 7; ++l) {
        bias_l[l] = boff;
        plan.biases.push_back({boff, g.L[l - 1].b_off, g.W, g.Wp});
        boff += g.Wp;
    }
    bias_s = boff; plan.biases.push_back({boff, g.L[7].b_off, 1, 16}); boff += 16;
    bias_f = boff; plan.biases.push_back({boff, g.L[7].b_off + 1, g.W, g.Wp}); boff += g.Wp;
    bias_9 = boff; plan.biases.push_back({boff, g.L[8].b_off, g.W2, g.W2p}); boff += g.W2p;
    bias_10 = boff; plan.biases.push_back({boff, g.L[9].b_off, 4, 16}); boff += 16;
    plan.bias_floats = boff;

    // ---- forward programs (train saves panels + masks, infer does not)
    for (int train = 0; train < 2; ++train) {
        TcProgram &P = train ? plan.fwd_train : plan.fwd_infer;
        P.wide = np > 4;
        Builder B(g, P);
        B.emit_prologue(EK_PROLOGUE_FWD, ENC_X, train ? sl.X() : -1);
        for (int l = 1; l <= 7; ++l) {
            const LayerGeom &L = g.L[l - 1];
            std::vector<Builder::KIn> kin;
            const bool skip = g.skip_layer && l == g.skip_layer + 1;
            if (l == 1 || skip) kin.push_back({B.e_slot, 4, L.w_off, L.in_dim, 1, g.Cx});
            if (l > 1) B.hidden_kin(kin, np, L.w_off, L.in_dim, 1, skip ? g.Cx : 0, g.W);
            B.emit_hidden_gemm(kin, np, g.W, EK_RELU, bias_l[l], train ? sl.H(l) : -1, train ? (l - 1) : -1, ENC_NONE, -1,
                               true);
        }
        {   // fc8 sigma row (out row 0) as its own N=16 GEMM, issued before the feature GEMM rewrites h7 in place
            const LayerGeom &L = g.L[7];
            std::vector<Builder::KIn> kin;
            B.hidden_kin(kin, np, L.w_off, L.in_dim, 1, 0, g.W);
            B.emit_small_gemm(kin, 1, EK_SIGMA, bias_s);
        }
        if (g.use_rgb_head) {
            {   // fc8 features: out rows 1..W, no activation
                const LayerGeom &L = g.L[7];
                std::vector<Builder::KIn> kin;
                B.hidden_kin(kin, np, L.w_off + L.in_dim, L.in_dim, 1, 0, g.W);
                B.emit_hidden_gemm(kin, np, g.W, EK_LINEAR, bias_f, train ? sl.feat() : -1, -1, g.Cd ? ENC_D : ENC_NONE,
                                   train ? sl.D() : -1, true);
            }
            {   // fc9 on [feat | dir]
                const LayerGeom &L = g.L[8];
                std::vector<Builder::KIn> kin;
                if (g.Cd) kin.push_back({B.e_slot, 2, L.w_off + g.W, L.in_dim, 1, g.Cd});
                B.hidden_kin(kin, np, L.w_off, L.in_dim, 1, 0, g.W);
                B.emit_hidden_gemm(kin, np2, g.W2, EK_RELU, bias_9, train ? sl.h9() : -1, train ? 7 : -1, ENC_NONE, -1, true);
            }
            {   // fc10 + sigmoid
                const LayerGeom &L = g.L[9];
                std::vector<Builder::KIn> kin;
                B.hidden_kin(kin, np2, L.w_off, L.in_dim, 1, 0, g.W2);
                B.emit_small_gemm(kin, 4, EK_RGBA, bias_10);
            }
        } else {
            // sigma-only network: nothing consumes h7 after the sigma GEMM
        }
        if (!B.err.empty()) { err = B.err; return false; }
        for (int i = 0; i < 3; ++i)
            if (B.pending[i]) { err = "schedule error: unconsumed ready event at tile end (fwd)"; return false; }
    }

    // ---- backward dgrad chain
    {
        TcProgram &P = plan.bwd;
        P.wide = np > 4;
        Builder B(g, P);
        if (g.use_rgb_head) {
            B.emit_prologue(EK_PROLOGUE_BWD, ENC_NONE, sl.R());
            {   // dH9 = dPre10 (K=16) * W10  -> mask(h9) -> dPre9
                const LayerGeom &L = g.L[9];
                std::vector<Builder::KIn> kin;
                // B chunk rows = in index (h9), cols = out index (4 valid): element = w[(c)*in_dim + r]
                kin.push_back({B.e_slot, 1, L.w_off, 1, L.in_dim, 4});
                B.emit_hidden_gemm(kin, np2, g.W2, EK_DMASK, 0, sl.dP9(), 7, ENC_DSIGMA, sl.Sg(), true);
            }
            {   // dFeat = dPre9 * W9[:, 0:W]
                const LayerGeom &L = g.L[8];
                std::vector<Builder::KIn> kin;
                B.hidden_kin(kin, np2, L.w_off, 1, L.in_dim, 0, g.W2);
                B.emit_hidden_gemm(kin, np, g.W, EK_DCOPY, 0, sl.dFeat(), -1, ENC_NONE, -1, true);
            }
        } else {
            B.emit_prologue(EK_PROLOGUE_BWD, ENC_DSIGMA, sl.Sg());
        }
        {   // dH7 = [dsigma | dFeat] * W8 -> mask(h7) -> dPre7
            const LayerGeom &L = g.L[7];
            std::vector<Builder::KIn> kin;
            kin.push_back({B.e_slot, 1, L.w_off, 1, L.in_dim, 1});
            if (g.use_rgb_head) B.hidden_kin(kin, np, L.w_off + L.in_dim, 1, L.in_dim, 0, g.W);
            B.emit_hidden_gemm(kin, np, g.W, EK_DMASK, 0, sl.dP(7), 6, ENC_NONE, -1, true);
        }
        for (int l = 7; l >= 2; --l) {  // dH_{l-1} = dPre_l * W_l -> mask(h_{l-1}) -> dPre_{l-1}
            const LayerGeom &L = g.L[l - 1];
            const bool skip = g.skip_layer && l == g.skip_layer + 1;
            std::vector<Builder::KIn> kin;
            B.hidden_kin(kin, np, L.w_off + (skip ? g.Cx : 0), 1, L.in_dim, 0, g.W);
            B.emit_hidden_gemm(kin, np, g.W, EK_DMASK, 0, sl.dP(l - 1), l - 2, ENC_NONE, -1, /*consumed=*/l > 2);
        }
        if (!B.err.empty()) { err = B.err; return false; }
        for (int i = 0; i < 3; ++i)
            if (B.pending[i]) { err = "schedule error: unconsumed ready event at tile end (bwd)"; return false; }
    }

    // ---- weight-gradient units: dW^T[in x out] = P^T Q  (P = layer input panels, Q = dPre panels)
    auto add_unit = [&](int n_p, int p0, int n_q, int q0, int m_valid, int n_valid, int64_t w_base, int row_stride,
                        int64_t b_base) {
        WgradUnit u;
        memset(&u, 0, sizeof(u));
        u.n_p = (uint8_t)n_p;
        u.n_q = (uint8_t)n_q;
        for (int i = 0; i < 6; ++i) u.p_slot[i] = (int16_t)(i < n_p ? p0 + i : -1);
        for (int i = 0; i < 4; ++i) u.q_slot[i] = (int16_t)(i < n_q ? q0 + i : -1);
        u.sg_slot = -1;
        u.sg_w_base = 0;
        u.sg_b_base = -1;
        u.m_valid = m_valid;
        u.n_valid = n_valid;
        u.w_base = w_base;
        u.w_row_stride = row_stride;
        u.b_base = b_base;
        plan.units.push_back(u);
    };
    // A layer block [P panels p0..p0+n_p) x [Q panels q0..q0+n_q) is cut into units of at most 4 x 4 panels (TMEM holds a
    // 256 x 256 fp32 block): hidden <= 256 gives one unit per block, hidden 512 four.
    auto add_block = [&](int n_p, int p0, int n_q, int q0, int m_valid, int n_valid, int64_t w_base, int row_stride, int64_t b_base) {
        for (int qa = 0; qa < n_q; qa += 4) {
            for (int pa = 0; pa < n_p; pa += 4) {
                const int np_u = n_p - pa < 4 ? n_p - pa : 4, nq_u = n_q - qa < 4 ? n_q - qa : 4;
                int mv = m_valid - 64 * pa, nv = n_valid - 64 * qa;
                mv = mv < 0 ? 0 : (mv > 64 * np_u ? 64 * np_u : mv);
                nv = nv < 0 ? 0 : (nv > 64 * nq_u ? 64 * nq_u : nv);
                if (mv == 0 || nv == 0) continue;
                add_unit(np_u, p0 + pa, nq_u, q0 + qa, mv, nv, w_base + (int64_t)(64 * qa) * row_stride + 64 * pa, row_stride,
                         (b_base >= 0 && pa == 0) ? b_base + 64 * qa : -1);
            }
        }
    };
    for (int l = 1; l <= 7; ++l) {
        const LayerGeom &L = g.L[l - 1];
        const bool skip = g.skip_layer && l == g.skip_layer + 1;
        if (l == 1) {
            add_block(1, sl.X(), np, sl.dP(1), g.Cx, g.W, L.w_off, L.in_dim, L.b_off);
        } else {
            if (skip) add_block(1, sl.X(), np, sl.dP(l), g.Cx, g.W, L.w_off, L.in_dim, -1);
            add_block(np, sl.H(l - 1), np, sl.dP(l), g.W, g.W, L.w_off + (skip ? g.Cx : 0), L.in_dim, L.b_off);
        }
    }
    {
        const LayerGeom &L = g.L[7];
        if (!g.use_rgb_head) {
            add_block(np, sl.H(7), 1, sl.Sg(), g.W, 1, L.w_off, L.in_dim, L.b_off);  // sigma row
        } else {
            // fc8 = [sigma row ; feature rows]. The feature units hold the H7 panels anyway: the ones of the first Q block also
            // produce their slice of the sigma row on the CUDA cores (a 1-column GEMM would re-read all of H7 for 1/256 of the work)
            const size_t first = plan.units.size();
            add_block(np, sl.H(7), np, sl.dFeat(), g.W, g.W, L.w_off + L.in_dim, L.in_dim, L.b_off + 1);
            for (size_t i = first; i < plan.units.size(); ++i) {
                WgradUnit &u = plan.units[i];
                if (u.q_slot[0] != sl.dFeat()) continue;          // only the units of the first Q block
                const int pa = u.p_slot[0] - sl.H(7);             // first H7 panel of this unit
                u.sg_slot = (int16_t)sl.Sg();
                u.sg_w_base = L.w_off + 64 * pa;
                u.sg_b_base = pa == 0 ? L.b_off : -1;
            }
        }
    }
    if (g.use_rgb_head) {
        const LayerGeom &L9 = g.L[8], &L10 = g.L[9];
        if (g.Cd && np <= 4 && np2 <= 2 && g.W % 64 == 0) {
            // fc9 over [features ; encoded direction] as ONE unit (5 M-side panels x <= 128 columns fit TMEM as 3 x 128):
            // dP9 is read once instead of once per input block
            add_unit(np + 1, sl.feat(), np2, sl.dP9(), g.W + g.Cd, g.W2, L9.w_off, L9.in_dim, L9.b_off);
            plan.units.back().p_slot[np] = (int16_t)sl.D();
        } else {
            add_block(np, sl.feat(), np2, sl.dP9(), g.W, g.W2, L9.w_off, L9.in_dim, L9.b_off);
            if (g.Cd) add_block(1, sl.D(), np2, sl.dP9(), g.Cd, g.W2, L9.w_off + g.W, L9.in_dim, -1);
        }
        add_block(np2, sl.h9(), 1, sl.R(), g.W2, 4, L10.w_off, L10.in_dim, L10.b_off);
    }
    plan.e_slot = np > 4 ? TC_SLOT_E_WIDE : TC_SLOT_E;
    plan.mask_words = (np > 4) ? 16 : 8;

    // ---- TS-mode programs (mlp_tc3.cu), hidden <= 256: every GEMM as steps of at most 128 output columns
    if (np <= 4) {
        struct KSrc { int a_src, kcount; int64_t src_base; int row_stride, col_stride, valid_cols; };
        auto hidden_src = [&](std::vector<KSrc> &k, int n_panels, int64_t src_base, int row_stride, int col_stride, int col0, int valid) {
            for (int p = 0; p < n_panels; ++p) {
                int v = valid - 64 * p;
                k.push_back({p, 4, src_base + (int64_t)(col0 + 64 * p) * col_stride, row_stride, col_stride, v < 0 ? 0 : (v > 64 ? 64 : v)});
            }
        };
        auto emit_prologue = [&](TsProgram &P, uint8_t kind, uint8_t enc, int enc_save_slot, uint8_t b_enc, int b_save_slot) {
            TsStep st;
            memset(&st, 0, sizeof(st));
            st.kind = kind; st.enc = enc;
            st.save_slot = -1; st.mask_slot = -1; st.enc_save_slot = (int16_t)enc_save_slot;
            st.b_enc = b_enc; st.b_save_slot = (int16_t)b_save_slot;
            P.steps.push_back(st);
        };

        // one GEMM of n_out (padded) output columns, `valid` of them real, as ceil(n_out / 128) steps
        uint8_t pending_pre = TS_PRE_NONE;   // set after a GEMM that is the last reader of a slot-E panel: the next step emitted carries the write
        auto emit_gemm = [&](TsProgram &P, const std::vector<KSrc> &kin, int n_out, int valid, uint8_t kind, uint32_t bias_off, int save_slot0,
                             int mask_slot, uint8_t enc, int enc_save_slot, bool writes_a) {
            const bool small = kind == EK_SIGMA || kind == EK_RGBA;
            const int nh = small ? 1 : (n_out + 127) / 128;
            for (int h = 0; h < nh; ++h) {
                const int ncols = small ? 32 : (n_out - 128 * h > 128 ? 128 : n_out - 128 * h);
                const int row0 = 128 * h;
                TsStep st;
                memset(&st, 0, sizeof(st));
                st.op_begin = (uint16_t)P.ops.size();
                for (size_t i = 0; i < kin.size(); ++i) {
                    const KSrc &k = kin[i];
                    TsOp op;
                    op.w_off = P.wpack_bytes;
                    op.n = (uint8_t)ncols;
                    op.a_src = (uint8_t)k.a_src;
                    op.kcount = (uint8_t)k.kcount;
                    op.first = i == 0 ? 1 : 0;
                    P.ops.push_back(op);
                    PackChunk pc;
                    pc.dst_off = P.wpack_bytes;
                    pc.n_rows = ncols;
                    pc.src_base = k.src_base + (int64_t)row0 * k.row_stride;
                    pc.row_stride = k.row_stride;
                    pc.col_stride = k.col_stride;
                    const int vr = valid - row0;
                    pc.valid_rows = vr < 0 ? 0 : (vr > ncols ? ncols : vr);
                    pc.valid_cols = k.valid_cols;
                    P.chunks.push_back(pc);
                    P.wpack_bytes += (uint32_t)ncols * 128u;
                }
                st.op_end = (uint16_t)P.ops.size();
                st.kind = kind;
                st.ncols = (uint8_t)ncols;
                st.a_col = (uint8_t)(64 * h);
                st.final_step = (h == nh - 1 && !small) ? 1 : 0;
                st.writes_a = (writes_a && !small) ? 1 : 0;
                st.mask_word0 = (uint8_t)(4 * h);
                st.bias_off = (uint16_t)(bias_off + 128u * h);
                st.save_slot = (int16_t)(save_slot0 < 0 ? -1 : save_slot0 + 2 * h);
                st.mask_slot = (int16_t)mask_slot;
                (void)enc; (void)enc_save_slot;   // (slot-E panels are written by the pre-encode actions, see mark_pre)
                st.enc = ENC_NONE;
                st.enc_save_slot = -1;
                P.steps.push_back(st);
                if (pending_pre != TS_PRE_NONE && h == 0) { P.steps.back().pre_enc = pending_pre; pending_pre = TS_PRE_NONE; }
            }
        };
        for (int train = 0; train < 2; ++train) {
            TsProgram &P = train ? plan.ts_fwd_train : plan.ts_fwd_infer;
            const bool dir = g.use_rgb_head && g.Cd;
            emit_prologue(P, EK_PROLOGUE_FWD, ENC_X, train ? sl.X() : -1, dir ? ENC_D : ENC_NONE, (dir && train) ? sl.D() : -1);
            pending_pre = TS_PRE_NONE;
            const int last_x_layer = g.skip_layer ? g.skip_layer + 1 : 1;   // the last GEMM that reads the encoded positions
            for (int l = 1; l <= 7; ++l) {
                const LayerGeom &L = g.L[l - 1];
                const bool skip = g.skip_layer && l == g.skip_layer + 1;
                std::vector<KSrc> kin;
                if (l == 1 || skip) kin.push_back({TS_A_SMEM, 4, L.w_off, L.in_dim, 1, g.Cx});
                if (l > 1) hidden_src(kin, np, L.w_off, L.in_dim, 1, skip ? g.Cx : 0, g.W);
                emit_gemm(P, kin, g.Wp, g.W, EK_RELU, bias_l[l], train ? sl.H(l) : -1, train ? (l - 1) : -1, ENC_NONE, -1, true);
                if (l == last_x_layer) pending_pre = TS_PRE_A;   // its successor's epilogue may write the next tile's positions
            }
            {   // fc8 sigma row, before the feature GEMM's epilogue rewrites h7
                const LayerGeom &L = g.L[7];
                std::vector<KSrc> kin;
                hidden_src(kin, np, L.w_off, L.in_dim, 1, 0, g.W);
                emit_gemm(P, kin, 32, 1, EK_SIGMA, bias_s, -1, -1, ENC_NONE, -1, false);
            }
            if (g.use_rgb_head) {
                {
                    const LayerGeom &L = g.L[7];
                    std::vector<KSrc> kin;
                    hidden_src(kin, np, L.w_off + L.in_dim, L.in_dim, 1, 0, g.W);
                    emit_gemm(P, kin, g.Wp, g.W, EK_LINEAR, bias_f, train ? sl.feat() : -1, -1, ENC_NONE, -1, true);
                }
                {
                    const LayerGeom &L = g.L[8];
                    std::vector<KSrc> kin;
                    if (g.Cd) kin.push_back({TS_A_SMEM_B, 2, L.w_off + g.W, L.in_dim, 1, g.Cd});
                    hidden_src(kin, np, L.w_off, L.in_dim, 1, 0, g.W);
                    emit_gemm(P, kin, g.W2p, g.W2, EK_RELU, bias_9, train ? sl.h9() : -1, train ? 7 : -1, ENC_NONE, -1, true);
                    if (g.Cd) pending_pre = TS_PRE_B;   // fc10's epilogue may write the next tile's encoded direction
                }
                {
                    const LayerGeom &L = g.L[9];
                    std::vector<KSrc> kin;
                    hidden_src(kin, np2, L.w_off, L.in_dim, 1, 0, g.W2);
                    emit_gemm(P, kin, 32, 4, EK_RGBA, bias_10, -1, -1, ENC_NONE, -1, false);
                }
            }
        }
        {
            TsProgram &P = plan.ts_bwd;
            if (g.use_rgb_head) {
                emit_prologue(P, EK_PROLOGUE_BWD, ENC_NONE, sl.R(), ENC_DSIGMA, sl.Sg());
                pending_pre = TS_PRE_NONE;
                {   // dH9 = dPre10 (K = 16) * W10 -> mask(h9) -> dPre9
                    const LayerGeom &L = g.L[9];
                    std::vector<KSrc> kin;
                    kin.push_back({TS_A_SMEM, 1, L.w_off, 1, L.in_dim, 4});
                    emit_gemm(P, kin, g.W2p, g.W2, EK_DMASK, 0, sl.dP9(), 7, ENC_NONE, -1, true);
                    pending_pre = TS_PRE_A;
                }
                {   // dFeat = dPre9 * W9[:, 0:W]
                    const LayerGeom &L = g.L[8];
                    std::vector<KSrc> kin;
                    hidden_src(kin, np2, L.w_off, 1, L.in_dim, 0, g.W2);
                    emit_gemm(P, kin, g.Wp, g.W, EK_DCOPY, 0, sl.dFeat(), -1, ENC_NONE, -1, true);
                }
            } else {
                emit_prologue(P, EK_PROLOGUE_BWD, ENC_DSIGMA, sl.Sg(), ENC_NONE, -1);   // (sigma-only network: d(sigma) is panel A)
                pending_pre = TS_PRE_NONE;
            }
            {   // dH7 = [dsigma | dFeat] * W8 -> mask(h7) -> dPre7
                const LayerGeom &L = g.L[7];
                std::vector<KSrc> kin;
                kin.push_back({g.use_rgb_head ? TS_A_SMEM_B : TS_A_SMEM, 1, L.w_off, 1, L.in_dim, 1});
                if (g.use_rgb_head) hidden_src(kin, np, L.w_off + L.in_dim, 1, L.in_dim, 0, g.W);
                emit_gemm(P, kin, g.Wp, g.W, EK_DMASK, 0, sl.dP(7), 6, ENC_NONE, -1, true);
                pending_pre = g.use_rgb_head ? TS_PRE_B : TS_PRE_A;
            }
            for (int l = 7; l >= 2; --l) {
                const LayerGeom &L = g.L[l - 1];
                const bool skip = g.skip_layer && l == g.skip_layer + 1;
                std::vector<KSrc> kin;
                hidden_src(kin, np, L.w_off + (skip ? g.Cx : 0), 1, L.in_dim, 0, g.W);
                emit_gemm(P, kin, g.Wp, g.W, EK_DMASK, 0, sl.dP(l - 1), l - 2, ENC_NONE, -1, true);   // (dPre1 has no consumer, but the saver warps read it from there)
            }
        }
    }
    return true;
}

// Split the weight-gradient units over n_ctas CTAs: each CTA gets up to kWgMaxSeg (unit, tile range) segments of (nearly) equal
// total fitted cost. Pure host logic (tests/test_tc_plan.py checks coverage and balance).
void tc_wgrad_partition(const std::vector<WgradUnit> &units, int n_ctas, int64_t n_tiles, std::vector<WgradWork> &work) {
    const int G = n_ctas;
    const int U = (int)units.size();
    // Cost of one tile of a unit in ns, fitted to tools/wgrad_marks.py: a half-tile ring iteration takes 516 ns + 87 ns per
    // 8 KB half panel -- a stage costs a fixed latency on top of its bytes -- (+ ~80 ns when a one-M-block unit also sums
    // biases: its epilogue pass outlasts its four MMAs). A segment adds its ring fill + accumulator flush (~20 us).
    std::vector<int64_t> tile_cost(U);
    int64_t total = 0;
    for (int i = 0; i < U; ++i) {
        const WgradUnit &u = units[i];
        tile_cost[i] = 2 * (87 * (u.n_p + u.n_q) + 516 + ((u.b_base >= 0 && u.n_p <= 2 && u.n_q >= 4) ? 80 : 0) + (u.sg_slot >= 0 ? 630 : 0));
        total += tile_cost[i] * n_tiles;
    }
    const int64_t seg_cost = 20000;
    // the CTAs walk the units in order, each taking `budget` worth of cost; returns whether everything was placed
    auto place = [&](int64_t budget) -> bool {
        work.assign((size_t)G, WgradWork{});
        int unit = 0;
        int64_t tile = 0;   // next unassigned tile of `unit`
        for (int c = 0; c < G && unit < U; ++c) {
            int64_t left = budget;
            WgradWork &w = work[c];
            while (unit < U && w.n_seg < kWgMaxSeg) {
                const int64_t fit = (left - seg_cost) / tile_cost[unit];
                if (fit < (w.n_seg ? 8 : 1)) break;     // not worth opening another segment for a few tiles
                const int64_t take = fit < n_tiles - tile ? fit : n_tiles - tile;
                w.seg[w.n_seg].unit = unit;
                w.seg[w.n_seg].tile_begin = (int)tile;
                w.seg[w.n_seg].tile_end = (int)(tile + take);
                ++w.n_seg;
                left -= seg_cost + take * tile_cost[unit];
                tile += take;
                if (tile >= n_tiles) { ++unit; tile = 0; }
            }
        }
        return unit >= U;
    };
    int64_t lo = total / G, hi = total + seg_cost * (U + 1) + 1;   // hi: one CTA could take everything (kWgMaxSeg permitting)
    while (!place(hi)) hi *= 2;
    while (hi - lo > 64) {   // smallest budget that places everything
        const int64_t mid = lo + (hi - lo) / 2;
        if (place(mid)) hi = mid; else lo = mid;
    }
    place(hi);
}
