"""A mixed call sequence through every path that overlaps host copies or the sampler with the previous step (device-resident
prediction + early loss read, pipelined nerf_predict_points, side-stream nerf_get_batch / nerf_train_iter), with pinned host
buffers that are rewritten between calls. Prints a digest of every value that came back; tests/test_gpu_e2e.py runs it twice --
as shipped, and with NERF_B200_STEP_SYNC=1 NERF_B200_NO_SAMPLER_OVERLAP=1 NERF_B200_NO_H2D_OVERLAP=1 (everything on one stream,
every step waited for) -- and the digests must be identical (deterministic weight gradients)."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_rs_b200 as nb  # noqa: E402

R, S, W = 512, 64, 64
CHUNK = int(os.environ.get("PIPELINE_WORKER_CHUNK", "0"))   # > 0: micro-batched steps (the forward re-runs inside nerf_step)
cfg = nb.default_config(image_w=W, image_h=W, num_rays=R, num_samples=S, hidden=128, deterministic_grads=1, max_rays_per_launch=CHUNK)
m = nb.NeRF(cfg)
rng = np.random.default_rng(5)
imgs = rng.random((4, W * W, 4)).astype(np.float32)
m.set_images(imgs)
m.set_view_angles(nb.get_view_angles(6)[:4])
tr = nb.Trainer(m)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h = hashlib.sha256()


def digest(*arrays):
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())


b0 = m.get_batch(None, None, 4, None, True, 1, want=("points", "t", "dirs", "gold"))
pts, tt, dirs, gold = pin(b0["points"].reshape(-1)), pin(b0["t"].reshape(-1)), pin(b0["dirs"].reshape(-1)), pin(b0["gold"].reshape(-1))
idx = pin(np.stack([rng.integers(0, W, R), rng.integers(0, W, R)], 1).astype(np.int64))
vi = pin(rng.integers(0, 4, 4).astype(np.int64))
jit = pin(np.sort(rng.random((R, S), dtype=np.float32), axis=1))
losses = []
for it in range(12):
    mode = it % 4
    if mode == 0:     # device-resident prediction, pipelined upload of the next batch; host buffers rewritten right after the step
        pred, _ = m.predict(pts, tt, dirs, train=True, lazy=True)
        losses.append(tr.step(pred, gold))
        pts *= np.float32(1.001)
        gold[:] = np.clip(gold * np.float32(0.999), 0, 1)
        if it % 8 == 0:
            digest(pred.numpy(), pred.densities())
    elif mode == 1:   # eager prediction straight after a step (once as an inference predict: the step then re-runs the forward itself)
        out, sig = m.predict(pts, tt, dirs, train=(it != 5))
        losses.append(tr.step(out, gold))
        digest(out, sig)
    elif mode == 2:   # host indices / jitter straight after a step, gold read back, resident predict
        b = m.get_batch(idx, vi, 4, jit, True, 0, want=("gold", "t"))
        idx[:, 0] = (idx[:, 0] + 1) % W
        jit[:] = np.sort((jit * np.float32(0.97) + np.float32(0.01)), axis=1)
        pred, _ = m.predict(train=True, lazy=True)
        losses.append(tr.step(pred, b["gold"].reshape(-1)))
        digest(b["gold"], b["t"], m.log_metrics(True, True)["prediction"])
    else:             # fused iterations (sampler on the side stream), then a read of everything
        for k in range(3):
            m.train_iter(100 + 10 * it + k)
        m.sync()
        losses.append(m.last_loss())
        digest(m.get_predictions(True, True)[0])
digest(np.array(losses, np.float32), m.get_weights(), m.get_grads())
print("DIGEST", h.hexdigest(), " ".join(f"{l:.6f}" for l in losses))
