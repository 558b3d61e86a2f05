"""ORACLE (test infrastructure, never shipped or timed as the product): plain PyTorch fp32 restatement of the RGB VAE
encoder the reference runs in front of the sampler.

What it restates
  * ``ldmseg/models/vae.py:36-39`` ``class GeneralVAEImage(AutoencoderKL)`` -- the arithmetic lives in the third-party
    ``diffusers`` package (PyPI, version un-pinned by the reference: tools/scripts/install_env_manual.sh:11), which is
    NOT vendored under /root/reference and is not installable here. The structure below follows the published SD-1.4
    ``vae/config.json`` (in 3, block_out_channels (128, 256, 512, 512), layers_per_block 2, latent 4, 32 groups, SiLU)
    and the public diffusers algorithm: ``Encoder`` (conv_in, DownEncoderBlock2D x 4 with ResnetBlock2D(eps 1e-6, no
    time embedding) and Downsample2D(padding=0) = F.pad(x, (0, 1, 0, 1)) + Conv2d(stride 2), UNetMidBlock2D with one
    single-head attention over the 512 channels, GroupNorm + SiLU + conv_out to 2*latent channels), then ``quant_conv``
    (1x1) and ``DiagonalGaussianDistribution`` (mean | logvar).
  * ``ldmseg/trainers/trainers_ldm_cond.py:336-396`` ``encode_inputs``: optional bilinear resize of the image,
    ``2 * images - 1``, ``encode(...).latent_dist.mode()``, optional bilinear resize of the latents, ``* scaling_factor``
    (restated verbatim as a free function; the class needs datasets / DDP to construct).
  * call sites: ``tools/main_ldm.py:138-140`` (from_pretrained, decoder dropped, set_scaling_factor),
    ``trainers_ldm_cond.py:1234-1239`` (rgb_latents of compute_pq).

PARITY UNPINNED at the diffusers boundary, exactly as for the UNet (oracle/unet_oracle.py): the reference holds no test
or fixture for the VAE and diffusers cannot be imported here.

State-dict key names equal diffusers' (``encoder.conv_in``, ``encoder.down_blocks.0.resnets.0.norm1``,
``encoder.mid_block.attentions.0.to_q`` ..., ``quant_conv``) so SD-1.4 VAE checkpoints load into it and into the CUDA
host mirror; the pre-0.15 attention names (``query / key / value / proj_attn``) are accepted by the mirror too.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

SD14_VAE = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                norm_num_groups=32)


class Resnet(nn.Module):
    """diffusers ResnetBlock2D(temb_channels=None, eps=1e-6, output_scale_factor=1)."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if hasattr(self, "conv_shortcut"):
            x = self.conv_shortcut(x)
        return x + h


class Downsample(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=0)."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if j == 0 else cout, cout, groups) for j in range(n)])
        if down:
            self.downsamplers = nn.ModuleList([Downsample(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if hasattr(self, "downsamplers"):
            x = self.downsamplers[0](x)
        return x


class MidAttention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, bias=True, norm_num_groups, eps=1e-6, residual_connection=True)."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        B, C, H, W = x.shape
        t = self.group_norm(x.view(B, C, H * W)).transpose(1, 2)      # [B, HW, C]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * (C ** -0.5), dim=-1)
        o = self.to_out[0](torch.bmm(p, v))
        return o.transpose(1, 2).reshape(B, C, H, W) + x


class MidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(c, c, groups), Resnet(c, c, groups)])
        self.attentions = nn.ModuleList([MidAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        blocks, c = [], boc[0]
        for i, co in enumerate(boc):
            blocks.append(DownBlock(c, co, layers_per_block, norm_num_groups, down=i < len(boc) - 1))
            c = co
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = MidBlock(c, norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VAEImageOracle(nn.Module):
    """AutoencoderKL with the decoder dropped (tools/main_ldm.py:139): encode(x) -> moments [B, 2*latent, H/8, W/8]."""

    def __init__(self, scaling_factor=0.18215, **cfg):
        super().__init__()
        c = dict(SD14_VAE)
        c.update(cfg)
        self.cfg = c
        self.encoder = Encoder(**c)
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)
        self.scaling_factor = scaling_factor

    def moments(self, x):
        return self.quant_conv(self.encoder(x))

    def mode(self, x):
        return torch.chunk(self.moments(x), 2, dim=1)[0]


def build_vae_image(seed=0, **cfg):
    torch.manual_seed(seed)
    return VAEImageOracle(**cfg).eval()


@torch.no_grad()
def encode_inputs(vae, images, scaling_factor=None, resize=None, latent_size=None):
    """trainers_ldm_cond.py:336-396 with sample_posterior=False. `latent_size` is the (h, w) the latents are resized to
    when `resize` is given (the reference's (latent_size, latent_size), generalised to non-square: SURVEY fact 7)."""
    if scaling_factor is None:
        scaling_factor = vae.scaling_factor
    if resize is not None:
        images = F.interpolate(images, size=resize, mode="bilinear", align_corners=False)
    images = 2. * images - 1.
    latents = vae.mode(images.float())
    latents_mean = latents.clone()
    if resize is not None:
        latents = F.interpolate(latents, size=latent_size, mode="bilinear", align_corners=False)
        latents_mean = F.interpolate(latents_mean, size=latent_size, mode="bilinear", align_corners=False)
    return latents * scaling_factor, latents_mean * scaling_factor
