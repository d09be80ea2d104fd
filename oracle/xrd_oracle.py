"""CPU oracle for the /denoise hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (state_dict in, tensors out) restatement of the
reference's inference algorithm.  It exists to *check* the CUDA path; it is
never the thing that is shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product package refuses to run when
its CUDA library is missing and never routes through this file.

Parity pinning: the reference repository has no tests, golden vectors or
fixtures for this path ("parity unpinned" upstream).  The
oracle is therefore pinned against outputs of the reference itself: the script
``tests/golden/make_golden.py`` imports the unmodified reference classes from
/root/reference (build container only), runs them on seeded weights/inputs and
stores the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function below against those vectors.

All arithmetic in the reference lives in the third-party dependency PyTorch
(``torch==2.1.0`` pinned in Backend/requirements.txt:4, not vendored; this
image has 2.11.0).  The torch functional ops used here (conv2d,
conv_transpose2d, group_norm, interpolate, pixel_shuffle, softmax, gelu, silu)
are the same ATen ops the reference's nn.Modules dispatch to.

File:line citations use the short names of SURVEY.md:
  HYB  = Backend/hybrid/hybrid3diffusionspeed.py
  DDIM = Backend/DDIM/DDIMModel.py
  NAF  = Backend/NafNet/NafnetModel.py
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------
# configuration records (mirror the reference constructors' keyword arguments)
# --------------------------------------------------------------------------
@dataclass
class UNetCfg:
    """UNetDiffusion.__init__ kwargs (HYB:309-310, DDIM:167-168)."""
    in_channels: int = 1
    model_channels: int = 48
    channel_mult: Tuple[int, ...] = (1, 2, 3, 4)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (3,)
    time_emb_dim: int = 192
    num_heads: int = 2          # AttentionBlock default (HYB:285)
    groups: int = 8             # nn.GroupNorm(8, C) everywhere (HYB:264,269,288,354)


@dataclass
class NAFCfg:
    """EnhancedNAFNet.__init__ kwargs (HYB:173-174, NAF:233-234)."""
    img_channel: int = 1
    width: int = 32
    middle_blk_num: int = 8
    enc_blk_nums: Tuple[int, ...] = (2, 2, 4, 6)
    dec_blk_nums: Tuple[int, ...] = (2, 2, 2, 2)


def _g(sd: SD, key: str, dtype) -> Tensor:
    return sd[key].to(dtype)


# --------------------------------------------------------------------------
# Enhanced NAFNet
# --------------------------------------------------------------------------
def layernorm2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """Per-pixel LayerNorm over channels, biased variance, eps inside sqrt
    (HYB:108-115, NAF:167-172)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    y = (x - u) / torch.sqrt(s + eps)
    return w[None, :, None, None] * y + b[None, :, None, None]


def simple_gate(x: Tensor) -> Tensor:
    """First half of the channels times second half (HYB:119-121)."""
    a, b = x.chunk(2, dim=1)
    return a * b


def nafblock(sd: SD, p: str, inp: Tensor) -> Tensor:
    """NAFBlock.forward (HYB:152-169, NAF:210-229)."""
    dt = inp.dtype
    c2 = sd[p + "conv2.weight"].shape[0]
    x = layernorm2d(inp, _g(sd, p + "norm1.weight", dt), _g(sd, p + "norm1.bias", dt))
    x = F.conv2d(x, _g(sd, p + "conv1.weight", dt), _g(sd, p + "conv1.bias", dt))
    x = F.conv2d(x, _g(sd, p + "conv2.weight", dt), _g(sd, p + "conv2.bias", dt), padding=1, groups=c2)
    x = simple_gate(x)
    pooled = x.mean(dim=(2, 3), keepdim=True)                       # AdaptiveAvgPool2d(1)
    sca = F.conv2d(pooled, _g(sd, p + "sca.1.weight", dt), _g(sd, p + "sca.1.bias", dt))
    x = x * sca
    x = F.conv2d(x, _g(sd, p + "conv3.weight", dt), _g(sd, p + "conv3.bias", dt))
    y = inp + x * _g(sd, p + "beta", dt)
    x = layernorm2d(y, _g(sd, p + "norm2.weight", dt), _g(sd, p + "norm2.bias", dt))
    x = F.conv2d(x, _g(sd, p + "conv4.weight", dt), _g(sd, p + "conv4.bias", dt))
    x = simple_gate(x)
    x = F.conv2d(x, _g(sd, p + "conv5.weight", dt), _g(sd, p + "conv5.bias", dt))
    return y + x * _g(sd, p + "gamma", dt)


def nafnet_forward(sd: SD, inp: Tensor, cfg: NAFCfg = NAFCfg(), prefix: str = "") -> Tensor:
    """EnhancedNAFNet.forward (HYB:206-238, NAF:275-309)."""
    dt = inp.dtype
    p = prefix
    B, C, H, W = inp.shape
    mult = 2 ** len(cfg.enc_blk_nums)
    ph, pw = (mult - H % mult) % mult, (mult - W % mult) % mult
    inp = F.pad(inp, (0, pw, 0, ph))
    x = F.conv2d(inp, _g(sd, p + "intro.weight", dt), _g(sd, p + "intro.bias", dt), padding=1)
    encs: List[Tensor] = []
    for s, nblk in enumerate(cfg.enc_blk_nums):
        for b in range(nblk):
            x = nafblock(sd, f"{p}encoders.{s}.{b}.", x)
        encs.append(x)
        x = F.conv2d(x, _g(sd, f"{p}downs.{s}.weight", dt), _g(sd, f"{p}downs.{s}.bias", dt), stride=2)
    for b in range(cfg.middle_blk_num):
        x = nafblock(sd, f"{p}middle_blks.{b}.", x)
    for s, nblk in enumerate(cfg.dec_blk_nums):
        x = F.conv2d(x, _g(sd, f"{p}ups.{s}.0.weight", dt))         # 1x1, no bias
        x = F.pixel_shuffle(x, 2)
        skip = encs[len(encs) - 1 - s]
        if x.shape[2:] != skip.shape[2:]:
            x = F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=False)
        x = torch.cat([x, skip], dim=1)
        x = F.conv2d(x, _g(sd, f"{p}skip_convs.{s}.weight", dt), _g(sd, f"{p}skip_convs.{s}.bias", dt))
        for b in range(nblk):
            x = nafblock(sd, f"{p}decoders.{s}.{b}.", x)
    x = F.conv2d(x, _g(sd, p + "ending.weight", dt), _g(sd, p + "ending.bias", dt), padding=1)
    x = x + inp
    return x[:, :, :H, :W]


# --------------------------------------------------------------------------
# conditional UNet
# --------------------------------------------------------------------------
def sinusoidal_embedding(t: Tensor, dim: int) -> Tensor:
    """SinusoidalPositionEmbeddings.forward (HYB:246-253).  ``t`` is expected
    to already carry the working dtype (float32, or float64 for the fp64
    oracle -- the reference module itself always yields float32)."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -k).to(t.dtype)
    ang = t[:, None] * freqs[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def time_mlp(sd: SD, p: str, t: Tensor, cfg: UNetCfg, dt) -> Tensor:
    """UNetDiffusion.time_mlp (HYB:313-318)."""
    e = sinusoidal_embedding(t.to(dt), cfg.model_channels)
    e = F.linear(e, _g(sd, p + "time_mlp.1.weight", dt), _g(sd, p + "time_mlp.1.bias", dt))
    e = F.silu(e)
    return F.linear(e, _g(sd, p + "time_mlp.3.weight", dt), _g(sd, p + "time_mlp.3.bias", dt))


def resblock(sd: SD, p: str, x: Tensor, temb: Tensor, groups: int = 8) -> Tensor:
    """ResidualBlock.forward (HYB:276-281)."""
    dt = x.dtype
    h = F.group_norm(x, groups, _g(sd, p + "block1.0.weight", dt), _g(sd, p + "block1.0.bias", dt), 1e-5)
    h = F.conv2d(F.silu(h), _g(sd, p + "block1.2.weight", dt), _g(sd, p + "block1.2.bias", dt), padding=1)
    te = F.linear(F.silu(temb), _g(sd, p + "time_mlp.1.weight", dt), _g(sd, p + "time_mlp.1.bias", dt))
    h = h + te[:, :, None, None]
    h = F.group_norm(h, groups, _g(sd, p + "block2.0.weight", dt), _g(sd, p + "block2.0.bias", dt), 1e-5)
    h = F.conv2d(F.silu(h), _g(sd, p + "block2.3.weight", dt), _g(sd, p + "block2.3.bias", dt), padding=1)
    if (p + "res_conv.weight") in sd:
        x = F.conv2d(x, _g(sd, p + "res_conv.weight", dt), _g(sd, p + "res_conv.bias", dt))
    return h + x


def attnblock(sd: SD, p: str, x: Tensor, heads: int = 2, groups: int = 8) -> Tensor:
    """AttentionBlock.forward (HYB:292-305; DDIM:143-166 is the same maths with
    the queries processed in chunks of 512)."""
    dt = x.dtype
    b, c, h, w = x.shape
    xn = F.group_norm(x, groups, _g(sd, p + "norm.weight", dt), _g(sd, p + "norm.bias", dt), 1e-5)
    qkv = F.conv2d(xn, _g(sd, p + "qkv.weight", dt), _g(sd, p + "qkv.bias", dt))
    qkv = qkv.reshape(b, 3, heads, c // heads, h * w)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]                      # (b, heads, d, n)
    scale = (c // heads) ** -0.5
    att = torch.matmul(q.transpose(-2, -1), k) * scale             # (b, heads, n_q, n_k)
    att = F.softmax(att, dim=-1)
    out = torch.matmul(att, v.transpose(-2, -1)).transpose(-2, -1) # (b, heads, d, n)
    out = out.reshape(b, c, h, w)
    out = F.conv2d(out, _g(sd, p + "proj.weight", dt), _g(sd, p + "proj.bias", dt))
    return out + x


def unet_layout(cfg: UNetCfg) -> Tuple[List[Tuple[str, int, int]], List[Tuple[str, int, int]]]:
    """The module lists UNetDiffusion.__init__ builds (HYB:322-351), as
    (kind, in_ch, out_ch) tuples with kind in {'res','attn','down','up'}."""
    downs, ups = [], []
    ch = cfg.model_channels
    nres = len(cfg.channel_mult)
    for i in range(nres):
        oc = cfg.model_channels * cfg.channel_mult[i]
        for _ in range(cfg.num_res_blocks):
            downs.append(("res", ch, oc)); ch = oc
            if i in cfg.attention_resolutions:
                downs.append(("attn", ch, ch))
        if i != nres - 1:
            downs.append(("down", ch, ch))
    for i in reversed(range(nres)):
        oc = cfg.model_channels * cfg.channel_mult[i]
        for _ in range(cfg.num_res_blocks + 1):
            ups.append(("res", ch + ch, oc)); ch = oc
            if i in cfg.attention_resolutions:
                ups.append(("attn", ch, ch))
        if i != 0:
            ups.append(("up", ch, ch))
    return downs, ups


def unet_forward(sd: SD, x: Tensor, cond: Tensor, t: Tensor, cfg: UNetCfg = UNetCfg(),
                 prefix: str = "") -> Tensor:
    """UNetDiffusion.forward (HYB:359-388) including the off-by-one skip
    bookkeeping and the bilinear resizes it forces."""
    dt = x.dtype
    p = prefix
    downs, ups = unet_layout(cfg)
    temb = time_mlp(sd, p, t, cfg, dt)
    h = torch.cat([x, cond], dim=1)
    h = F.conv2d(h, _g(sd, p + "in_conv.weight", dt), _g(sd, p + "in_conv.bias", dt), padding=1)
    skips: List[Tensor] = []
    for i, (kind, _, _) in enumerate(downs):
        q = f"{p}downs.{i}."
        if kind == "res":
            h = resblock(sd, q, h, temb, cfg.groups)
        elif kind == "attn":
            h = attnblock(sd, q, h, cfg.num_heads, cfg.groups)
        else:
            h = F.conv2d(h, _g(sd, q + "weight", dt), _g(sd, q + "bias", dt), stride=2, padding=1)
        skips.append(h)
    h = resblock(sd, p + "mid_block1.", h, temb, cfg.groups)
    h = attnblock(sd, p + "mid_attn.", h, cfg.num_heads, cfg.groups)
    h = resblock(sd, p + "mid_block2.", h, temb, cfg.groups)
    for i, (kind, _, _) in enumerate(ups):
        q = f"{p}ups.{i}."
        if kind == "res":
            skip = skips.pop()
            if h.shape[2:] != skip.shape[2:]:
                h = F.interpolate(h, size=skip.shape[2:], mode="bilinear", align_corners=False)
            h = torch.cat([h, skip], dim=1)
            h = resblock(sd, q, h, temb, cfg.groups)
        elif kind == "attn":
            h = attnblock(sd, q, h, cfg.num_heads, cfg.groups)
        else:
            h = F.conv_transpose2d(h, _g(sd, q + "weight", dt), _g(sd, q + "bias", dt), stride=2, padding=1)
    h = F.group_norm(h, cfg.groups, _g(sd, p + "out_conv.0.weight", dt), _g(sd, p + "out_conv.0.bias", dt), 1e-5)
    return F.conv2d(F.silu(h), _g(sd, p + "out_conv.2.weight", dt), _g(sd, p + "out_conv.2.bias", dt), padding=1)


# --------------------------------------------------------------------------
# the sampler
# --------------------------------------------------------------------------
def ddim_timesteps(noise_steps: int, inference_steps: int) -> List[int]:
    """Timestep indices DiffusionDenoiser.denoise visits (HYB:404-406): note
    this is ``noise_steps // inference_steps`` strided, so the number of UNet
    evaluations is not ``inference_steps`` in general."""
    step = max(1, noise_steps // inference_steps)
    return list(reversed(range(0, noise_steps, step)))


def ddim_tables(noise_steps: int = 50, beta_start: float = 1e-4, beta_end: float = 0.02,
                dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """beta / alpha / alpha_hat (HYB:396-398).  The reference builds them in
    float32; pass float64 only for the fp64 study oracle."""
    beta = torch.linspace(beta_start, beta_end, noise_steps, dtype=torch.float32).to(dtype)
    alpha = 1.0 - beta
    return beta, alpha, torch.cumprod(alpha, dim=0)


def ddim_update(x: Tensor, eps: Tensor, alpha_t: Tensor, alpha_hat_t: Tensor) -> Tensor:
    """One reverse update (HYB:410-416): clamp eps to +-5, posterior mean with
    the single-step alpha[t] and no noise term, clamp x to [0,1]."""
    eps = torch.clamp(eps, -5, 5)
    x = (1 / torch.sqrt(alpha_t)) * (x - ((1 - alpha_t) / torch.sqrt(1 - alpha_hat_t)) * eps)
    return torch.clamp(x, 0, 1)


def ddim_denoise(sd: SD, noisy: Tensor, inference_steps: int, noise_steps: int = 50,
                 cfg: UNetCfg = UNetCfg(), prefix: str = "", trace: Optional[dict] = None,
                 teacher_x: Optional[Sequence[Tensor]] = None) -> Tensor:
    """DiffusionDenoiser.denoise (HYB:400-418).  ``trace`` (optional dict)
    receives lists 'x_in' (input of every UNet evaluation) and 'eps' (its raw,
    unclamped output).  ``teacher_x`` replaces the loop state before every
    evaluation (teacher forcing) so per-step eps can be compared in isolation."""
    dt = noisy.dtype
    _, alpha, alpha_hat = ddim_tables(noise_steps, dtype=dt)
    x = noisy.clone()
    for n, i in enumerate(ddim_timesteps(noise_steps, inference_steps)):
        if teacher_x is not None:
            x = teacher_x[n].to(dt)
        t = torch.full((x.shape[0],), i, dtype=torch.long, device=x.device)
        eps = unet_forward(sd, x, noisy, t, cfg, prefix)
        if trace is not None:
            trace.setdefault("x_in", []).append(x.clone())
            trace.setdefault("eps", []).append(eps.clone())
        x = ddim_update(x, eps, alpha[i], alpha_hat[i])
    return x


# --------------------------------------------------------------------------
# router, fusion, hybrid
# --------------------------------------------------------------------------
def _cgg(sd: SD, p: str, x: Tensor, groups: int, stride: int = 1) -> Tensor:
    """conv3x3 -> GroupNorm -> exact (erf) GELU, the router/fusion building
    block (HYB:473-477 etc.)."""
    dt = x.dtype
    x = F.conv2d(x, _g(sd, p + "0.weight", dt), _g(sd, p + "0.bias", dt), stride=stride, padding=1)
    x = F.group_norm(x, groups, _g(sd, p + "1.weight", dt), _g(sd, p + "1.bias", dt), 1e-5)
    return F.gelu(x)


def router_forward(sd: SD, x: Tensor, prefix: str = "") -> Tensor:
    """NoiseAnalyzer.forward (HYB:511-534): soft sigmoid mask, never binarised."""
    dt = x.dtype
    p = prefix
    e1 = _cgg(sd, p + "enc1.", x, 8)
    e2 = _cgg(sd, p + "enc2.", e1, 8, stride=2)
    e3 = _cgg(sd, p + "enc3.", e2, 8, stride=2)
    m = _cgg(sd, p + "mid.", e3, 8)
    d3 = F.conv_transpose2d(m, _g(sd, p + "up3.weight", dt), _g(sd, p + "up3.bias", dt), stride=2)
    if d3.shape[2:] != e2.shape[2:]:
        d3 = F.interpolate(d3, size=e2.shape[2:], mode="bilinear", align_corners=False)
    d3 = _cgg(sd, p + "dec3.", torch.cat([d3, e2], dim=1), 8)
    d2 = F.conv_transpose2d(d3, _g(sd, p + "up2.weight", dt), _g(sd, p + "up2.bias", dt), stride=2)
    if d2.shape[2:] != e1.shape[2:]:
        d2 = F.interpolate(d2, size=e1.shape[2:], mode="bilinear", align_corners=False)
    d2 = _cgg(sd, p + "dec2.", torch.cat([d2, e1], dim=1), 8)
    if d2.shape[2:] != x.shape[2:]:
        d2 = F.interpolate(d2, size=x.shape[2:], mode="bilinear", align_corners=False)
    return torch.sigmoid(F.conv2d(d2, _g(sd, p + "out_conv.weight", dt), _g(sd, p + "out_conv.bias", dt)))


def fusion_forward(sd: SD, naf: Tensor, diff: Tensor, mask: Tensor, prefix: str = "") -> Tensor:
    """FusionModule.forward (HYB:552-557): a 3-layer conv stack, no attention."""
    dt = naf.dtype
    p = prefix
    x = torch.cat([naf, diff, mask], dim=1)
    x = _cgg(sd, p + "conv1.", x, 8)
    x = _cgg(sd, p + "conv2.", x, 4)
    return F.conv2d(x, _g(sd, p + "out_conv.weight", dt), _g(sd, p + "out_conv.bias", dt))


# --------------------------------------------------------------------------
# ExpertDenoiser (the 4th /denoise output; RUN:52-57,127)
# --------------------------------------------------------------------------
def _cbr(sd: SD, p: str, i: int, x: Tensor) -> Tensor:
    """Conv2d(3x3, bias=False) -> BatchNorm2d in eval mode (running statistics, eps 1e-5) -> ReLU
    (DirectUNetModel.py:163-168 and the other Sequential blocks of that class)."""
    dt = x.dtype
    y = F.conv2d(x, _g(sd, f"{p}{i}.weight", dt), None, padding=1)
    b = f"{p}{i + 1}."
    y = F.batch_norm(y, _g(sd, b + "running_mean", dt), _g(sd, b + "running_var", dt), _g(sd, b + "weight", dt), _g(sd, b + "bias", dt),
                     training=False, eps=1e-5)
    return F.relu(y)


def expert_forward(sd: SD, x: Tensor, prefix: str = "") -> Tensor:
    """ExpertDenoiser.forward (DirectUNet/DirectUNetModel.py:232-255)."""
    p, dt = prefix, x.dtype
    pair = lambda q, t: _cbr(sd, q, 3, _cbr(sd, q, 0, t))
    x1 = pair(p + "inc.", x)
    x2 = pair(p + "down1.", x1)
    x3 = pair(p + "down2.", F.max_pool2d(x2, 2))
    x4 = pair(p + "bottleneck.", F.max_pool2d(x3, 2))
    d2 = F.conv_transpose2d(x4, _g(sd, p + "up2.weight", dt), _g(sd, p + "up2.bias", dt), stride=2)
    d2 = pair(p + "upconv2.", torch.cat([d2, x3], dim=1))
    d1 = F.conv_transpose2d(d2, _g(sd, p + "up1.weight", dt), _g(sd, p + "up1.bias", dt), stride=2)
    d1 = pair(p + "upconv1.", torch.cat([d1, x2], dim=1))
    d1 = _cbr(sd, p + "final.", 0, d1)
    return F.conv2d(d1, _g(sd, p + "outc.weight", dt), _g(sd, p + "outc.bias", dt))


def _sanitize(x: Tensor) -> Tensor:
    """nan_to_num(nan=0,posinf=1,neginf=0) + clamp(0,1) (HYB:615-616)."""
    return torch.clamp(torch.nan_to_num(x, nan=0.0, posinf=1.0, neginf=0.0), 0, 1)


def hybrid_forward(sd: SD, noisy: Tensor, inference_steps: int, noise_steps: int = 50,
                   naf_cfg: NAFCfg = NAFCfg(), unet_cfg: UNetCfg = UNetCfg(),
                   parts: Optional[dict] = None, trace: Optional[dict] = None) -> Tensor:
    """HybridDenoisingRouter.forward in eval mode (HYB:610-628).  ``parts``
    (optional dict) receives 'naf', 'diff', 'mask'."""
    naf = _sanitize(nafnet_forward(sd, noisy, naf_cfg, "nafnet."))
    diff = _sanitize(ddim_denoise(sd, noisy, inference_steps, noise_steps, unet_cfg, "diffusion_unet.", trace))
    mask = _sanitize(router_forward(sd, noisy, "router."))
    if parts is not None:
        parts.update(naf=naf, diff=diff, mask=mask)
    return fusion_forward(sd, naf, diff, mask, "fusion.")


# --------------------------------------------------------------------------
# seeded weights / inputs: neutral helpers (no reference arithmetic) that live in synthetic_data.py at the repository
# root so that bench.py's GPU arm can build its workload without importing this file; re-exported for the tests
# --------------------------------------------------------------------------
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthetic_data import is_norm_param, randomize_batchnorm_params, randomize_identity_params, synthetic_xray  # noqa: E402,F401


def synthetic_state_dict(shapes: Dict[str, Sequence[int]], seed: int = 1234) -> SD:
    """A seeded state_dict with the given key -> shape table (tests/golden/meta.json: the reference's 912 hybrid keys), for the
    CPU timing legs of bench.py: the cost of this path does not depend on the values, and building the weights here keeps the
    reference arm free of any import of the product package.  Conv / linear weights ~ N(0, 1/fan_in), norm weights around
    1, biases and NAFBlock beta / gamma small (never the identity)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for k in sorted(shapes):
        shp = tuple(int(v) for v in shapes[k])
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("beta", "gamma"):
            sd[k] = 0.5 * torch.randn(shp, generator=g)
        elif len(shp) >= 2:
            fan_in = 1
            for v in shp[1:]:
                fan_in *= v
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(max(1, fan_in))
        elif leaf == "weight":
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shp, generator=g)
    return sd


# --------------------------------------------------------------------------
# Overlap tiling (BASELINE configs[4]).  No reference counterpart: this is the DEFINITION the CUDA kernels in
# csrc/tiles.cu are checked against (bit-exact; fp32 numpy, same operation order, no fused multiply-add).
# --------------------------------------------------------------------------
def tile_origins(length: int, tile: int, halo: int) -> List[int]:
    if length == tile:
        return [0]
    stride = tile - 2 * halo
    n = -(-(length - tile) // stride) + 1
    return [min(k * stride, length - tile) for k in range(n)]


def tile_weights(k: int, n: int, tile: int, halo: int):
    import numpy as np
    r = max(1, 2 * halo)
    i = np.arange(tile)
    m = np.full(tile, r)
    if k != 0:
        m = np.minimum(m, i + 1)
    if k != n - 1:
        m = np.minimum(m, tile - i)
    return m.astype(np.float32) / np.float32(r)


def extract_tiles(img: Tensor, tile: int, halo: int) -> Tensor:
    b, _, h, w = img.shape
    oy, ox = tile_origins(h, tile, halo), tile_origins(w, tile, halo)
    out = [img[i:i + 1, :, y:y + tile, x:x + tile] for i in range(b) for y in oy for x in ox]
    return torch.cat(out, 0).contiguous()


def blend_tiles(tiles: Tensor, batch: int, height: int, width: int, tile: int, halo: int) -> Tensor:
    import numpy as np
    oy, ox = tile_origins(height, tile, halo), tile_origins(width, tile, halo)
    t = tiles.detach().cpu().numpy().astype(np.float32).reshape(batch, len(oy), len(ox), tile, tile)
    num = np.zeros((batch, height, width), np.float32)
    den = np.zeros((height, width), np.float32)
    for iy, y in enumerate(oy):
        wy = tile_weights(iy, len(oy), tile, halo)
        for ix, x in enumerate(ox):
            w = (wy[:, None] * tile_weights(ix, len(ox), tile, halo)[None, :]).astype(np.float32)
            num[:, y:y + tile, x:x + tile] = num[:, y:y + tile, x:x + tile] + (w[None] * t[:, iy, ix]).astype(np.float32)
            den[y:y + tile, x:x + tile] = den[y:y + tile, x:x + tile] + w
    return torch.from_numpy((num / den[None]).astype(np.float32))[:, None]

