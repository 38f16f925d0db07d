"""tch `.ot` VarStore checkpoints: NeRF::save / NeRF::load (src/model.rs:211-217, called at src/main.rs:48-50, 81-83).

The reference saves with `VarStore::save` -> tch `Tensor::save_multi` -> libtorch `torch::serialize::OutputArchive`:
one `archive.write(name, tensor)` per variable, i.e. a TorchScript module archive whose PARAMETERS are the named
variables. Host-side container plumbing only -- torch is used for the zip/pickle format, never for compute.

Variable names. All ten `nn::linear` layers are created on the same `vs.root()` path (src/model.rs:48-55, 89-90), so every
layer asks for "bias" and "weight". tch resolves a colliding name by appending `__{number of variables registered so
far}`, and `nn::linear` registers the bias before the weight, which gives
    fc1: bias, weight   fc2: bias__2, weight__3   fc3: bias__4, weight__5   ...   fc10: bias__18, weight__19.
tch is not vendored in the reference tree (only a caret range in Cargo.toml), so this scheme is restated from the crate's
published source, not checked against a file the Rust binary wrote: PARITY UNPINNED for the names. The container format
itself is pinned: tests/golden/tch_varstore.ot is written by libtorch's own OutputArchive (tests/golden/make_ot_fixture.cpp).
The reader therefore does not trust the order inside a pair: it sorts variables by their numeric suffix and tells bias from
weight by rank, so a weight-first tch version loads as well.

Unlike the reference, Adam's state can travel in the same file (`adam_m`, `adam_v`, `adam_step`): `VarStore::load` looks its
own variables up by name and ignores extra entries, so such a file still loads in nerf-rs.
"""
import re

import numpy as np

ADAM_KEYS = ("adam_m", "adam_v", "adam_step")


def var_names(n_layers=10):
    """(bias_name, weight_name) per layer, in creation order fc1..fc{n}."""
    out = []
    for layer in range(n_layers):
        k = 2 * layer
        out.append(("bias" if k == 0 else f"bias__{k}", "weight" if k == 0 else f"weight__{k + 1}"))
    return out


def layer_dims(cfg):
    """[(in, out)] of fc1..fc10 for a nerf_config (same rule as the library's build_geom; src/model.rs:48-55, 89-90)."""
    w = cfg.hidden
    cx = 3 + 6 * cfg.xyz_freqs
    cd = 0 if cfg.dir_freqs < 0 else 3 + 6 * cfg.dir_freqs
    dims = []
    for layer in range(1, 8):
        din = cx if layer == 1 else w
        if cfg.skip_layer and layer == cfg.skip_layer + 1:
            din = w + cx
        dims.append((din, w))
    dims += [(w, w + 1), (w + cd, w // 2), (w // 2, 4)]
    return dims


def _holder_class():
    import torch

    class Module(torch.nn.Module):   # an attribute-only module: what OutputArchive builds
        def forward(self):
            return 0

    return Module


def write_ot(path, named):
    """named: iterable of (name, float32/int64 ndarray). Every entry becomes a parameter of the archived module."""
    import torch
    holder = _holder_class()()
    for name, arr in named:
        t = torch.from_numpy(np.ascontiguousarray(arr).copy())
        holder.register_parameter(name, torch.nn.Parameter(t, requires_grad=False))
    torch.jit.script(holder).save(str(path))


def read_ot(path):
    """-> {name: ndarray} of every parameter and buffer in the archive (tch reads both through `named_parameters`)."""
    import torch
    m = torch.jit.load(str(path), map_location="cpu")
    out = {n: t.detach().cpu().numpy() for n, t in m.named_parameters()}
    out.update({n: t.detach().cpu().numpy() for n, t in m.named_buffers() if n not in out})
    return out


def _suffix(name):
    m = re.search(r"__(\d+)$", name)
    return int(m.group(1)) if m else -1


def layers_from_vars(tensors, layer_dims):
    """Pair up the VarStore variables of `tensors` into [(weight [out,in], bias [out])] for fc1..fcN, checking shapes against
    layer_dims = [(in, out)]. Variables are ordered by (suffix, name); within a pair rank decides which is the bias."""
    names = [n for n in tensors if n not in ADAM_KEYS]
    base = [n for n in names if _suffix(n) < 0]
    rest = sorted((n for n in names if _suffix(n) >= 0), key=_suffix)
    # the two unsuffixed variables come first, in whichever order nn::linear made them
    order = sorted(base, key=lambda n: tensors[n].ndim) + rest
    if len(order) != 2 * len(layer_dims):
        raise ValueError(f"expected {2 * len(layer_dims)} VarStore variables, found {len(order)}: {sorted(names)}")
    layers = []
    for i, (din, dout) in enumerate(layer_dims):
        a, b = tensors[order[2 * i]], tensors[order[2 * i + 1]]
        w, bias = (a, b) if a.ndim == 2 else (b, a)
        if w.shape != (dout, din) or bias.shape != (dout,):
            raise ValueError(f"fc{i + 1}: checkpoint has weight {w.shape} / bias {bias.shape}, the model needs ({dout}, {din}) / ({dout},)")
        layers.append((w.astype(np.float32), bias.astype(np.float32)))
    return layers


def flat_from_layers(layers):
    """The library's flat parameter blob: fc1..fcN, each `[out,in]` row-major then the bias."""
    return np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in layers]).astype(np.float32)


def layers_from_flat(flat, layer_dims):
    out, off = [], 0
    for din, dout in layer_dims:
        w = flat[off:off + din * dout].reshape(dout, din)
        off += din * dout
        b = flat[off:off + dout]
        off += dout
        out.append((w, b))
    if off != flat.size:
        raise ValueError("flat parameter blob does not match the layer dimensions")
    return out


def save_varstore(path, flat_weights, layer_dims, adam=None):
    named = []
    for (bn, wn), (w, b) in zip(var_names(len(layer_dims)), layers_from_flat(np.asarray(flat_weights, np.float32), layer_dims)):
        named += [(bn, b), (wn, w)]
    if adam is not None:
        m, v, step = adam
        named += [("adam_m", np.asarray(m, np.float32)), ("adam_v", np.asarray(v, np.float32)), ("adam_step", np.asarray([step], np.int64))]
    write_ot(path, named)


def load_varstore(path, layer_dims):
    """-> (flat_weights, adam or None)."""
    t = read_ot(path)
    flat = flat_from_layers(layers_from_vars(t, layer_dims))
    adam = None
    if all(k in t for k in ADAM_KEYS):
        adam = (t["adam_m"].astype(np.float32), t["adam_v"].astype(np.float32), int(np.asarray(t["adam_step"]).reshape(-1)[0]))
    return flat, adam
