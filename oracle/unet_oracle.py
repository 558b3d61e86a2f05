"""ORACLE (test infrastructure, never shipped or timed as the product): plain PyTorch fp32 restatement of the UNet that
the reference's sampler calls.

What it restates
  * ``ldmseg/models/unet.py:24`` ``class UNet(UNet2DConditionModel)`` -- the arithmetic lives in the third-party
    ``diffusers`` package (PyPI, version un-pinned by the reference: tools/scripts/install_env_manual.sh:11), which is
    NOT vendored under /root/reference and is not installable here (no network).  The block structure below follows
    the published SD-1.4 ``unet/config.json`` and the public diffusers UNet2DConditionModel algorithm.
  * ``ldmseg/models/unet.py:281-436``  forward (timestep expand, conv_in, 4 down blocks with 12 skips, mid block,
    4 up blocks popping 3 skips each, conv_norm_out + SiLU + conv_out).
  * ``ldmseg/models/unet.py:178-233``  modify_encoder (8/12-channel conv_in built from the 4-channel one).
  * ``ldmseg/models/unet.py:83-105``   remove_cross_attention (attn2 / norm2 dropped -> self-attention only); with
    ``cross_attention_dim=768`` the blocks keep norm2 / attn2 against ``encoder_hidden_states`` (unet.py:319-323:
    ``encoder_hid_proj``, learnable ``object_queries``).

One independent known answer pins the SHAPES of this restatement: with cross_attention_dim=768 and in_channels=4 it has
exactly the published 859 520 964 parameters of Stable Diffusion 1.x's UNet (tests/test_host_cpu.py); the arithmetic
itself stays unpinned:

PARITY UNPINNED at the diffusers boundary: the reference holds no test, golden vector or fixture for the UNet and
diffusers cannot be imported here, so nothing independent pins this restatement (DESIGN.md says the same).
Two documented generalisations: noise/latents may be non-square, and Upsample2D uses the
upstream ``upsample_size`` rule (size of the next skip) when a spatial dim is not a multiple of 8 -- identical to the
reference wherever the reference's own forward does not crash.

State-dict key names equal diffusers' (conv_in, time_embedding.linear_1, down_blocks.0.resnets.0.norm1, ...,
``new_conv`` alias after modify_encoder) so reference checkpoints load into it and into the CUDA host mirror.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

USE_SDPA = False  # module switch, set by bench.py's torch_gpu_baseline leg

SD14 = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
    norm_groups=32, norm_eps=1e-5, attn_levels=(True, True, True, False), temb_mult=4,
)


def sinusoid_freqs(half):
    # diffusers get_timestep_embedding: exp(-ln(10000) * arange(half) / (half - downscale_freq_shift)), shift = 0
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    return torch.exp(exponent / half)


def timestep_features(timesteps, dim):
    # flip_sin_to_cos=True: [cos | sin]
    half = dim // 2
    arg = timesteps[:, None].float() * sinusoid_freqs(half).to(timesteps.device)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimeMLP(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class Resnet(nn.Module):
    """diffusers ResnetBlock2D (default time-scale-shift, output_scale_factor 1, dropout 0)."""

    def __init__(self, cin, cout, temb, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        sc = self.conv_shortcut(x) if hasattr(self, "conv_shortcut") else x
        return sc + h


class SelfAttn(nn.Module):
    """diffusers Attention: self-attention (ctx_dim None) or cross-attention against `context` [b, L, ctx_dim]."""

    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, context=None):
        b, n, c = x.shape
        d = c // self.heads
        ctx = x if context is None else context

        def split(t):
            return t.view(b, t.shape[1], self.heads, d).transpose(1, 2)

        q, k, v = split(self.to_q(x)), split(self.to_k(ctx)), split(self.to_v(ctx))
        if USE_SDPA:  # the same arithmetic through torch's fused kernel (what current diffusers calls): bench.py's
            o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)  # GPU-library baseline leg only
        else:
            w = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
            o = (w @ v).transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, 2 * inner)

    def forward(self, x):
        val, gate = self.proj(x).chunk(2, dim=-1)
        return val * F.gelu(gate)  # erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, 4 * dim), nn.Identity(), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class TransformerBlock(nn.Module):
    """BasicTransformerBlock; with ctx_dim None attn2/norm2 are removed (ldmseg/models/unet.py:83-105), otherwise the
    block is norm1/attn1, norm2/attn2 (cross-attention against encoder_hidden_states), norm3/ff as in diffusers."""

    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = SelfAttn(dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        if ctx_dim is not None:  # (created last so that the default path draws the same random weights as before)
            self.norm2 = nn.LayerNorm(dim)
            self.attn2 = SelfAttn(dim, heads, ctx_dim)

    def forward(self, x, context=None):
        x = x + self.attn1(self.norm1(x))
        if hasattr(self, "attn2"):
            x = x + self.attn2(self.norm2(x), context)
        return x + self.ff(self.norm3(x))


class Transformer2D(nn.Module):
    def __init__(self, dim, heads, groups, ctx_dim=None):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([TransformerBlock(dim, heads, ctx_dim)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        t = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = self.transformer_blocks[0](t, context)
        t = t.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self.proj_out(t) + x


class ConvHolder(nn.Module):
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1)


class DownStage(nn.Module):
    def __init__(self, cin, cout, n, temb, heads, groups, eps, attn, down, ctx_dim=None):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(n)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2D(cout, heads, groups, ctx_dim) for _ in range(n)])
        if down:
            self.downsamplers = nn.ModuleList([ConvHolder(cout, 2)])

    def forward(self, x, emb, context=None):
        skips = []
        for i, r in enumerate(self.resnets):
            x = r(x, emb)
            if hasattr(self, "attentions"):
                x = self.attentions[i](x, context)
            skips.append(x)
        if hasattr(self, "downsamplers"):
            x = self.downsamplers[0].conv(x)
            skips.append(x)
        return x, skips


class MidStage(nn.Module):
    def __init__(self, c, temb, heads, groups, eps, ctx_dim=None):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(c, c, temb, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2D(c, heads, groups, ctx_dim)])

    def forward(self, x, emb, context=None):
        x = self.resnets[0](x, emb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, emb)


class UpStage(nn.Module):
    def __init__(self, cins, cout, temb, heads, groups, eps, attn, up, ctx_dim=None):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(ci, cout, temb, groups, eps) for ci in cins])
        if attn:
            self.attentions = nn.ModuleList([Transformer2D(cout, heads, groups, ctx_dim) for _ in cins])
        if up:
            self.upsamplers = nn.ModuleList([ConvHolder(cout, 1)])

    def forward(self, x, emb, skips, upsample_size, context=None):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), emb)
            if hasattr(self, "attentions"):
                x = self.attentions[i](x, context)
        if hasattr(self, "upsamplers"):
            if upsample_size is None:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            else:
                x = F.interpolate(x, size=upsample_size, mode="nearest")
            x = self.upsamplers[0].conv(x)
        return x


class UNetOracle(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(SD14)
        c.update(cfg)
        self.cfg = c
        ch = list(c["block_out_channels"])
        n, heads, groups, eps = c["layers_per_block"], c["heads"], c["norm_groups"], c["norm_eps"]
        temb = ch[0] * c["temb_mult"]
        xd = c.get("cross_attention_dim")  # None: cross-attention removed (the default path)
        self.conv_in = nn.Conv2d(c["in_channels"], ch[0], 3, padding=1)
        self.time_embedding = TimeMLP(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        prev = ch[0]
        for i, co in enumerate(ch):
            self.down_blocks.append(DownStage(prev, co, n, temb, heads, groups, eps, c["attn_levels"][i], i < len(ch) - 1, xd))
            prev = co
        self.mid_block = MidStage(ch[-1], temb, heads, groups, eps, xd)
        self.up_blocks = nn.ModuleList()
        rev = ch[::-1]
        prev = rev[0]
        for i, co in enumerate(rev):
            skip_in = rev[min(i + 1, len(ch) - 1)]
            cins = []
            for j in range(n + 1):
                res_skip = skip_in if j == n else co
                res_in = prev if j == 0 else co
                cins.append(res_in + res_skip)
            attn = c["attn_levels"][::-1][i]
            self.up_blocks.append(UpStage(cins, co, temb, heads, groups, eps, attn, i < len(ch) - 1, xd))
            prev = co
        self.conv_norm_out = nn.GroupNorm(groups, ch[0], eps=eps)
        self.conv_out = nn.Conv2d(ch[0], c["out_channels"], 3, padding=1)

    # ldmseg/models/unet.py:178-233 (in_channels == 8 branch)
    @torch.no_grad()
    def modify_encoder(self, in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=0,
                       init_mode_cond="zero", **unused):
        assert in_channels in (4, 8)
        if in_channels != 8:
            return
        old = self.conv_in
        new = nn.Conv2d(in_channels + cond_channels, old.out_channels, 3, padding=1)
        for sl, mode in ((slice(0, 4), init_mode_seg), (slice(4, 8), init_mode_image)):
            if mode == "copy":
                new.weight[:, sl] = old.weight
            elif mode == "div":
                new.weight[:, sl] = old.weight  # the reference's "/ 2." is applied to a discarded temporary
            elif mode == "mean":
                new.weight[:, sl] = old.weight.mean(dim=1, keepdim=True).repeat(1, 4, 1, 1)
            elif mode == "zero":
                new.weight[:, sl] = 0
            elif mode != "random":
                raise NotImplementedError(mode)
        new.bias.copy_(old.bias)
        if cond_channels > 0:
            if init_mode_cond == "zero":
                new.weight[:, 8:] = 0
            elif init_mode_cond != "random":
                raise NotImplementedError(init_mode_cond)
        self.new_conv = new
        self.conv_in = new

    # ldmseg/models/unet.py:122-123 and the 'learnable' descriptor branch (descriptors.py:89-91)
    def modify_encoder_hidden_state_proj(self, in_channels, out_channels):
        self.encoder_hid_proj = nn.Linear(in_channels, out_channels)

    def define_learnable_embeddings(self, num_queries, dim):
        self.object_queries = nn.Embedding(num_queries, dim)

    def forward(self, sample, timestep, encoder_hidden_states=None):
        """Returns the epsilon prediction tensor (the reference wraps it in UNetOutput(sample=...))."""
        ctx = encoder_hidden_states
        if self.cfg.get("cross_attention_dim") is None:
            assert ctx is None, "cross-attention is removed on this path (base.yaml:71)"
        if hasattr(self, "encoder_hid_proj") and ctx is not None:   # unet.py:319-320
            ctx = self.encoder_hid_proj(ctx)
        if hasattr(self, "object_queries"):                          # unet.py:322-323
            ctx = self.object_queries.weight.unsqueeze(0).repeat(sample.shape[0], 1, 1)
        ts = torch.as_tensor(timestep, device=sample.device).expand(sample.shape[0])
        emb = self.time_embedding(timestep_features(ts, self.conv_in.out_channels).to(sample.dtype))
        n_up = len(self.up_blocks) - 1
        need_size = any(s % (2 ** n_up) != 0 for s in sample.shape[-2:])
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, s = blk(x, emb, ctx)
            skips.extend(s)
        x = self.mid_block(x, emb, ctx)
        for i, blk in enumerate(self.up_blocks):
            n_res = len(blk.resnets)
            size = None
            if need_size and i < n_up:
                size = skips[-n_res - 1].shape[-2:]
            x = blk(x, emb, skips, size, ctx)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def build_unet(seed=0, model_kwargs=None, learnable_queries=None, **cfg):
    """Random-init oracle UNet the way tools/main_ldm.py:147-160 builds it (from_pretrained -> random init here,
    remove_cross_attention, modify_encoder(**model_kwargs))."""
    torch.manual_seed(seed)
    net = UNetOracle(**cfg)
    if learnable_queries:  # descriptors.py:89-91: image_descriptors == 'learnable'
        net.define_learnable_embeddings(*learnable_queries)
    mk = dict(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=0)
    mk.update(model_kwargs or {})
    net.modify_encoder(**mk)
    return net.eval()
