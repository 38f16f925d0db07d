"""Cycle breakdown of the TS-mode chain kernel (k_chain3) from its per-CTA counters.
Needs a library built with NERF_B200_NVCC_EXTRA=-DNERF_TC3_STATS (tools/gpu_tc3_stats.sh copies build/ab_stats.so in place).
Prints, per program, the median over CTAs of: MMA-thread total / waiting for EPI_DONE / waiting for weights, and
epilogue-warp-0 total / waiting for ACC_FULL / tcgen05.ld+wait / waiting for a free staging buffer."""
import ctypes
import sys
import numpy as np
import nerf_rs_b200 as nb


def read_stats(model, ctas=148):
    out = np.zeros((ctas, 16), np.uint64)
    rc = model.lib.nerf_debug_tc3_stats(out.ctypes.data, ctas)
    assert rc == 0, rc
    return out.astype(np.float64)


def show(tag, st):
    lead = st[0::2]   # leader CTAs carry the MMA thread
    med = lambda a: float(np.median(a))
    steps = med(lead[:, 3])
    items = med(st[:, 12])
    print(f"{tag}: MMA thread total {med(lead[:, 0]):.0f} clk, wait EPI_DONE {med(lead[:, 1]):.0f} ({med(lead[:, 1]) / max(steps, 1):.0f}/step), "
          f"wait weights {med(lead[:, 2]):.0f}, MMA issue blocks {med(lead[:, 4]):.0f} ({med(lead[:, 4]) / max(steps, 1):.0f}/step), lane steps {steps:.0f}")
    print(f"{tag}: epilogue total {med(st[:, 8]):.0f} clk, wait ACC_FULL {med(st[:, 9]):.0f} ({med(st[:, 9]) / max(items, 1):.0f}/item), "
          f"ld {med(st[:, 10]):.0f} ({med(st[:, 10]) / max(items, 1):.0f}/item), wait ACT_SAVED {med(st[:, 11]) / max(items, 1):.0f}/item, "
          f"signal {med(st[:, 13]) / max(items, 1):.0f}/item, convert (incl. ld, early signal) {med(st[:, 14]) / max(items, 1):.0f}/item, "
          f"tcgen05.st+wait {med(st[:, 15]) / max(items, 1):.0f}/item, items {items:.0f}")


def show_trace(model, first=0, count=120):
    ev = np.zeros(4096, np.uint64)
    assert model.lib.nerf_debug_tc3_trace(ev.ctypes.data, 4096) == 0
    tags = (ev >> np.uint64(48)).astype(int)
    clk = (ev & np.uint64(0xffffffffffff)).astype(np.int64)
    names = {1: "step", 2: "epi_done", 3: "weights", 4: "op", 5: "acc_commit"}  # (weights are waited for before EPI_DONE)
    t0 = clk[first]
    prev = t0
    line = []
    for i in range(first, min(first + count, 4096)):
        if tags[i] == 0:
            break
        if tags[i] == 1 and line:
            print("  ".join(line)); line = []
        line.append(f"{names.get(tags[i], tags[i])}+{clk[i] - prev}")
        prev = clk[i]
    if line:
        print("  ".join(line))


def show2(tag, st):
    med = lambda a: float(np.median(a))
    print(f"{tag}: epilogue warp 0 time in first halves {med(st[:, 5]):.0f}, second halves {med(st[:, 6]):.0f}, plain steps {med(st[:, 7]):.0f} (of {med(st[:, 8]):.0f})")


def main():
    rays, samples = 4096, 64
    cfg = nb.default_config(image_w=800, image_h=800, num_rays=rays, num_samples=samples, hidden=256)
    m = nb.NeRF(cfg, device=0)
    rng = np.random.default_rng(0)
    m.set_images(rng.random((4, 800 * 800, 4)).astype(np.float32))
    m.set_view_angles(nb.get_view_angles(6))
    m.get_batch(None, None, 64, None, True, 1, want=())
    for _ in range(3):
        m.predict(train=False)
    m.sync()
    show("infer ", read_stats(m)); show2("infer ", read_stats(m))
    if "--trace" in sys.argv:
        show_trace(m, 200, 160)
    for i in range(3):
        m.train_iter(i)
    m.sync()
    show("dgrad ", read_stats(m)); show2("dgrad ", read_stats(m))
    if "--trace" in sys.argv:
        show_trace(m, 200, 200)   # the last k_chain3 launch of a training iteration is the backward chain
    m.predict(train=True)
    m.sync()
    show("fwd-tr", read_stats(m)); show2("fwd-tr", read_stats(m))


if __name__ == "__main__":
    main()
