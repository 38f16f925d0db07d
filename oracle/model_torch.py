"""torch-CPU restatement of the reference's tensor path -- TEST INFRASTRUCTURE ONLY.

tch 0.11 (un-vendored, ``Cargo.toml:15``) binds libtorch/ATen; torch 2.11 is the
same ATen, so each reference op is restated with its torch twin:

  DensityNet / RadianceNet   src/model.rs:44-67, 85-131
  NeRF::predict              src/model.rs:152-209
  accumulated_transmittance  src/model.rs:221-232   (the literal S-op graph)
  compositing                src/model.rs:234-249   (incl. the ``.view`` scramble, :241)
  mse_loss                   src/model.rs:296-299
  Trainer::new / step        src/model.rs:306-325   (nn::Adam::default, lr cli.rs:64-65)

PARITY UNPINNED: the reference's tests pin no tensor value and
the binary cannot be built here; closed-form checks stand in.

Config-driven: ``ModelConfig.as_shipped()`` is the literal
reference (W=100, raw xyz, no skip, colours (s,s,s,1), scrambled T);
``ModelConfig()`` is the north-star default (W=256, L=10/4, skip, RGB head used).
"""
from dataclasses import dataclass, replace
import math

import torch

T_FAR = 2.0  # src/ray_sampling.rs:12


@dataclass(frozen=True)
class ModelConfig:
    hidden: int = 256          # HIDDEN_NODES (model.rs:12 ships 100)
    xyz_freqs: int = 10        # 0 -> raw xyz (INDIM = 3, model.rs:11)
    dir_freqs: int = 4         # -1 -> no view-direction input (as shipped)
    skip_layer: int = 5        # concat [x_enc, h] after this layer's ReLU; 0 -> none
    use_rgb_head: bool = True  # False -> colours (sigma,sigma,sigma,1) (model.rs:190-206)
    sigma_relu: bool = False   # reference applies no activation to sigma (model.rs:168-171)
    bug_compat_T_view: bool = False  # model.rs:241 ``stack(...,0).view((R,S))``
    emulate_bf16: bool = False       # round weights/activations where the CUDA path does
    emulate_bf16_grads: bool = False  # also round d(pre-activation) to bf16 in backward

    @staticmethod
    def as_shipped():
        return ModelConfig(hidden=100, xyz_freqs=0, dir_freqs=-1, skip_layer=0, use_rgb_head=False,
                           bug_compat_T_view=True)

    @property
    def cx(self):
        return 3 + 6 * self.xyz_freqs

    @property
    def cd(self):
        return 0 if self.dir_freqs < 0 else 3 + 6 * self.dir_freqs

    def layer_dims(self):
        """[(in, out)] for fc1..fc10, creation order of model.rs:48-55 then :89-90."""
        w, cx, cd = self.hidden, self.cx, self.cd
        dims = []
        for l in range(1, 8):
            k = cx if l == 1 else w
            if self.skip_layer and l == self.skip_layer + 1:
                k = w + cx
            dims.append((k, w))
        dims.append((w, w + 1))          # fc8: sigma || features (model.rs:55)
        dims.append((w + cd, w // 2))    # fc9 (model.rs:89; + direction per :87-88, :175)
        dims.append((w // 2, 4))         # fc10 (model.rs:90), LABELS = 4
        return dims

    def num_params(self):
        return sum(i * o + o for i, o in self.layer_dims())


def init_params(cfg, seed=0):
    """nn.Linear default init per layer, fc1..fc10, after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    params = []
    for i, o in cfg.layer_dims():
        lin = torch.nn.Linear(i, o)
        params.append((lin.weight.detach().clone(), lin.bias.detach().clone()))
    return params


def flatten_params(params):
    """Flat f32 blob: per layer weight[out,in] row-major, then bias[out]."""
    return torch.cat([torch.cat([w.reshape(-1), b.reshape(-1)]) for w, b in params])


def unflatten_params(cfg, flat):
    out, off = [], 0
    for i, o in cfg.layer_dims():
        w = flat[off:off + i * o].reshape(o, i)
        off += i * o
        b = flat[off:off + o]
        off += o
        out.append((w, b))
    assert off == flat.numel()
    return out


def posenc(x, num_freqs):
    """Same convention as oracle.ray_np.posenc, in torch f32."""
    outs = [x]
    for k in range(max(num_freqs, 0)):
        a = x * float(2.0 ** k)
        outs.append(torch.sin(a))
        outs.append(torch.cos(a))
    return torch.cat(outs, dim=-1)


class _RoundBF16(torch.autograd.Function):
    """Forward: round to bf16 (what the kernel stores as the next MMA operand).
    Backward: straight-through."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGradBF16(torch.autograd.Function):
    """Identity forward; the gradient w.r.t. a pre-activation is rounded to bf16
    (the CUDA backward stores it as a bf16 MMA operand)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _rb(cfg, x):
    return _RoundBF16.apply(x) if cfg.emulate_bf16 else x


def _rg(cfg, z):
    return _RoundGradBF16.apply(z) if cfg.emulate_bf16_grads else z


def _lin(cfg, x, wb):
    w, b = wb
    return _rg(cfg, torch.nn.functional.linear(x, _rb(cfg, w), b))


def mlp_forward(cfg, params, x, d=None):
    """x[B, cx] encoded positions, d[B, cd] encoded directions (per sample).

    Returns (sigma[B], rgba[B,4], features[B,W]). model.rs:97-131, 168-178."""
    x = _rb(cfg, x)
    h = x
    for l in range(1, 8):
        h = _rb(cfg, torch.relu(_lin(cfg, h, params[l - 1])))
        if cfg.skip_layer and l == cfg.skip_layer:
            h = torch.cat([x, h], dim=-1)
    df = _lin(cfg, h, params[7])                # [B, W+1], no activation
    sigma = df[:, 0]
    if cfg.sigma_relu:
        sigma = torch.relu(sigma)
    feat = _rb(cfg, df[:, 1:])
    f9 = feat if cfg.cd == 0 else torch.cat([feat, _rb(cfg, d)], dim=-1)
    h9 = _rb(cfg, torch.relu(_lin(cfg, f9, params[8])))
    rgba = torch.sigmoid(_lin(cfg, h9, params[9]))
    return sigma, rgba, feat


def deltas_from_t(t):
    """model.rs:184-187: delta_i = t_{i+1} - t_i, last = T_FAR - t_last."""
    tfar = torch.full((t.shape[0], 1), T_FAR, dtype=t.dtype)
    return torch.cat([t[:, 1:], tfar], dim=1) - t


def accumulated_transmittance(densities, distances, i):
    """model.rs:221-232, literal."""
    if i == 0:
        return torch.ones(densities.shape[0], dtype=densities.dtype)
    return (densities[:, 0:i] * distances[:, 0:i]).sum(dim=1).neg().exp()


def compositing_literal(densities, colors, distances, bug_compat_T_view=False):
    """model.rs:234-249 op for op: S separate transmittance chains, stack, weights, sum.

    bug_compat_T_view=True keeps the reference's ``stack(dim 0).view((R,S))`` (:241),
    which reinterprets the [S,R] stack's memory; False transposes (intended [ray,sample])."""
    r, s = densities.shape
    ts = [accumulated_transmittance(densities, distances, i) for i in range(s)]
    stacked = torch.stack(ts, 0)  # [S, R]
    T = stacked.view(r, s) if bug_compat_T_view else stacked.t()
    weights = (T * (1.0 - (densities * distances).neg().exp())).unsqueeze(2)
    return (weights * colors).sum(dim=1)


def compositing(densities, colors, distances):
    """Scan form of the same formula: a = exp(-sigma*delta),
    T = exclusive cumprod(a), w = T (1 - a). Returns (out[R,4], weights[R,S])."""
    a = (densities * distances).neg().exp()
    T = torch.cumprod(torch.cat([torch.ones_like(a[:, :1]), a[:, :-1]], dim=1), dim=1)
    w = T * (1.0 - a)
    return (w.unsqueeze(2) * colors).sum(dim=1), w


def predict(cfg, params, query_points, distances, num_rays, num_points, dirs=None, literal=True):
    """NeRF::predict (model.rs:152-209).

    query_points[B*3] raw sample positions, distances[B] = t values, dirs[R,3] world
    ray directions (north-star only). Returns (pixels[R,4], sigma[R,S])."""
    b = num_rays * num_points
    assert tuple(query_points.shape) == (b * 3,)      # model.rs:162
    assert tuple(distances.shape) == (b,)             # model.rs:163
    x = posenc(query_points.view(b, 3), cfg.xyz_freqs)
    d = None
    if cfg.cd:
        d = posenc(dirs, cfg.dir_freqs).unsqueeze(1).expand(num_rays, num_points, cfg.cd).reshape(b, cfg.cd)
    sigma, rgba, _ = mlp_forward(cfg, params, x, d)
    sigma = sigma.view(num_rays, num_points)
    delta = deltas_from_t(distances.view(num_rays, num_points))
    if cfg.use_rgb_head:
        colors = rgba.view(num_rays, num_points, 4)
    else:  # model.rs:192-204
        colors = torch.stack([sigma, sigma, sigma, torch.ones_like(sigma)], 0).permute(1, 2, 0)
    if literal or cfg.bug_compat_T_view:
        out = compositing_literal(sigma, colors, delta, cfg.bug_compat_T_view)
    else:
        out, _ = compositing(sigma, colors, delta)
    return out, sigma


def mse_loss(x, y):
    """model.rs:296-299: mean over all R*4 elements (alpha included)."""
    diff = x - y
    return (diff * diff).mean()


class Trainer:
    """Trainer::new / Trainer::step (model.rs:301-325): Adam defaults
    (betas .9/.999, eps 1e-8, no weight decay, no amsgrad), lr from cli.rs:64-65."""

    def __init__(self, cfg, params, lr=5e-4):
        self.cfg = cfg
        self.params = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in params]
        flat = [p for wb in self.params for p in wb]
        self.opt = torch.optim.Adam(flat, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False)

    def predict(self, query_points, distances, num_rays, num_points, dirs=None, literal=True):
        return predict(self.cfg, self.params, query_points, distances, num_rays, num_points, dirs, literal)

    def step(self, predictions, gold):
        assert predictions.dim() == 2 and predictions.shape[1] == 4     # model.rs:315
        assert tuple(gold.shape) == (predictions.shape[0] * 4,)         # model.rs:316
        loss = mse_loss(predictions, gold.view(-1, 4))
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def grads_flat(self):
        def g(p):  # as shipped fc9/fc10 receive no gradient: report zeros
            return (p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
        return torch.cat([torch.cat([g(w), g(b)]) for w, b in self.params])

    def params_flat(self):
        return flatten_params([(w.detach(), b.detach()) for w, b in self.params])


def adam_reference(p, g, m, v, step, lr=5e-4, b1=0.9, b2=0.999, eps=1e-8):
    """One Adam update in the op order of libtorch's Adam (what tch's nn::Adam calls)."""
    m = m * b1 + g * (1 - b1)
    v = v * b2 + g * g * (1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v
