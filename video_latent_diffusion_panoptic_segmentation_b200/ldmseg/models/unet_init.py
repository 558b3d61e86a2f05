"""Shapes and random initialisation of the UNet / seg-AE parameters with the reference's state-dict key names.

There is no network for the SD-1.4 checkpoint that ``UNet.from_pretrained`` loads in the reference
(tools/main_ldm.py:147), so benchmarks and smoke tests use random weights of the same architecture
(SD-1.4 unet/config.json with cross-attention removed). Initialisation follows torch's defaults
(Conv/Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias; norms: weight 1, bias 0).
"""
import math

import torch


def unet_param_shapes(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, in_channels=4, out_channels=4,
                      attn_levels=(True, True, True, False), temb_mult=4, cross_attention_dim=None):
    """Ordered {key: shape} of diffusers' UNet2DConditionModel; without attn2/norm2 (unet.py:83-105) unless
    cross_attention_dim is given (768 in the SD-1.4 config)."""
    ch = list(block_out_channels)
    temb = ch[0] * temb_mult
    S = {}

    def conv(name, cin, cout, k):
        S[name + ".weight"], S[name + ".bias"] = (cout, cin, k, k), (cout,)

    def lin(name, cin, cout, bias=True):
        S[name + ".weight"] = (cout, cin)
        if bias:
            S[name + ".bias"] = (cout,)

    def norm(name, c):
        S[name + ".weight"], S[name + ".bias"] = (c,), (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", temb, cout)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def transformer(name, c):
        norm(name + ".norm", c)
        conv(name + ".proj_in", c, c, 1)
        tb = name + ".transformer_blocks.0"
        norm(tb + ".norm1", c)
        for p in ("to_q", "to_k", "to_v"):
            lin(f"{tb}.attn1.{p}", c, c, bias=False)
        lin(tb + ".attn1.to_out.0", c, c)
        if cross_attention_dim:
            norm(tb + ".norm2", c)
            lin(tb + ".attn2.to_q", c, c, bias=False)
            lin(tb + ".attn2.to_k", cross_attention_dim, c, bias=False)
            lin(tb + ".attn2.to_v", cross_attention_dim, c, bias=False)
            lin(tb + ".attn2.to_out.0", c, c)
        norm(tb + ".norm3", c)
        lin(tb + ".ff.net.0.proj", c, 8 * c)
        lin(tb + ".ff.net.2", 4 * c, c)
        conv(name + ".proj_out", c, c, 1)

    conv("conv_in", in_channels, ch[0], 3)
    lin("time_embedding.linear_1", ch[0], temb)
    lin("time_embedding.linear_2", temb, temb)
    prev = ch[0]
    for i, co in enumerate(ch):
        for j in range(layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if attn_levels[i]:
            for j in range(layers_per_block):
                transformer(f"down_blocks.{i}.attentions.{j}", co)
        if i < len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        prev = co
    resnet("mid_block.resnets.0", ch[-1], ch[-1])
    resnet("mid_block.resnets.1", ch[-1], ch[-1])
    transformer("mid_block.attentions.0", ch[-1])
    rev = ch[::-1]
    prev = rev[0]
    for i, co in enumerate(rev):
        skip_in = rev[min(i + 1, len(ch) - 1)]
        for j in range(layers_per_block + 1):
            res_skip = skip_in if j == layers_per_block else co
            res_in = prev if j == 0 else co
            resnet(f"up_blocks.{i}.resnets.{j}", res_in + res_skip, co)
        if attn_levels[::-1][i]:
            for j in range(layers_per_block + 1):
                transformer(f"up_blocks.{i}.attentions.{j}", co)
        if i < len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    norm("conv_norm_out", ch[0])
    conv("conv_out", ch[0], out_channels, 3)
    return S


def seg_decoder_param_shapes(out_channels=128, int_channels=256, latent_channels=4, num_upscalers=2,
                             upscale_channels=256, **unused):
    """decoder.* keys of GeneralVAESeg (ldmseg/models/vae.py:124-173) with num_mid_blocks=0."""
    S = {"decoder.0.weight": (int_channels, latent_channels, 3, 3), "decoder.0.bias": (int_channels,)}
    idx, dim = 2, upscale_channels
    for i in range(num_upscalers):
        cin = int_channels if i == 0 else dim
        S[f"decoder.{idx}.weight"], S[f"decoder.{idx}.bias"] = (cin, dim, 2, 2), (dim,)       # ConvTranspose2d
        S[f"decoder.{idx + 1}.weight"], S[f"decoder.{idx + 1}.bias"] = (dim,), (dim,)         # LayerNorm2d
        idx += 3
    S[f"decoder.{idx}.weight"], S[f"decoder.{idx}.bias"] = (dim,), (dim,)                     # GroupNorm
    S[f"decoder.{idx + 2}.weight"], S[f"decoder.{idx + 2}.bias"] = (out_channels, dim, 3, 3), (out_channels,)
    return S


def vae_image_param_shapes(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                           layers_per_block=2, **unused):
    """encoder.* / quant_conv.* keys of diffusers' AutoencoderKL (SD-1.4 vae/config.json) -- the half of
    GeneralVAEImage that tools/main_ldm.py:138-140 keeps."""
    S = {}

    def conv(name, cin, cout, k):
        S[name + ".weight"], S[name + ".bias"] = (cout, cin, k, k), (cout,)

    def lin(name, cin, cout):
        S[name + ".weight"], S[name + ".bias"] = (cout, cin), (cout,)

    def norm(name, c):
        S[name + ".weight"], S[name + ".bias"] = (c,), (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    boc = list(block_out_channels)
    conv("encoder.conv_in", in_channels, boc[0], 3)
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        c = co
    resnet("encoder.mid_block.resnets.0", c, c)
    resnet("encoder.mid_block.resnets.1", c, c)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", c, c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", c, 2 * latent_channels, 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    return S


def random_state_dict(shapes, seed=0, transposed_conv_keys=()):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in shapes.items():
        base = key.rsplit(".", 1)[0]
        wshape = shapes.get(base + ".weight", shape)
        if len(wshape) == 1:  # normalisation layer
            sd[key] = torch.ones(shape) if key.endswith(".weight") else torch.zeros(shape)
            continue
        if base in transposed_conv_keys:  # ConvTranspose2d: fan_in is computed from dim 1 by torch
            fan_in = wshape[1] * math.prod(wshape[2:])
        else:
            fan_in = math.prod(wshape[1:])
        bound = 1.0 / math.sqrt(fan_in)
        sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def random_unet_state_dict(seed=0, **cfg):
    return random_state_dict(unet_param_shapes(**cfg), seed)


def random_seg_decoder_state_dict(seed=1, **cfg):
    shapes = seg_decoder_param_shapes(**cfg)
    n_up = cfg.get("num_upscalers", 2)
    tkeys = tuple(f"decoder.{2 + 3 * i}" for i in range(n_up))
    return random_state_dict(shapes, seed, transposed_conv_keys=tkeys)


def random_vae_image_state_dict(seed=2, **cfg):
    return random_state_dict(vae_image_param_shapes(**cfg), seed)


# ------------------------------------------------------------------------------------------------ "trained-like" recipe
# Plain random init makes the whole tail degenerate: the UNet's epsilon (std 0.3) is a 2 % correction of a final latent
# that is 15 x the initial noise, the decoder's softmax never reaches mask_th, every pixel becomes void and PQ / DVPQ
# are 0 = 0. A trained LDMSeg model differs in four ways, and the recipe restores exactly those on top of the random
# weights (no layer is bypassed or shrunk; every kernel runs on the same dense shapes):
#   1. the prediction, not the initial noise, determines the final latent     -> conv_out weight and bias x conv_out_gain
#   2. the prediction follows the image                                       -> init_mode_image='copy' (a reference
#      mode, ldmseg/models/unet.py:178-233) and image latents that are piecewise constant over drifting Voronoi cells
#      (ldmseg/data/synthetic.py)
#   3. the decoder is smooth: its two ConvTranspose2d(2, 2) layers do not invent a different class for each of the 16
#      sub-pixel positions of a latent pixel                                   -> the four taps of every channel pair
#      are equal (tap (0, 0) of the random draw): up-sampling + channel mix
#   4. the classifier head is fitted to its input: a random 256 -> 128 projection slices the features along directions
#      in which they hardly vary, so the two best classes are within 1 % of each other on a few percent of the pixels
#      and the id map is speckle. One-shot stand-in for training the head (fit_seg_head): the last conv becomes a
#      nearest-centroid classifier over k-means centroids of the features of ONE teacher frame (3x3 box filter on all
#      nine taps, classes beyond the centroids switched off by their bias) -- coherent segments that persist from
#      frame to frame, margins like a trained model's.
TRAINED_LIKE = dict(conv_out_gain=32.0, rgb_amplitude=300.0, regions=40, drift=0.5,
                    head_centroids=48, head_gain=40.0, head_offset=1.0, head_iters=15, head_stride=7,
                    head_skip_labels=(12, 13, 14, 15, 16, 17, 18))
TRAINED_LIKE_MODEL_KWARGS = dict(in_channels=8, init_mode_seg="copy", init_mode_image="copy", cond_channels=0)


def trained_like_unet_(sd, conv_out_gain=TRAINED_LIKE["conv_out_gain"]):
    """In place on a UNet state dict (diffusers key names): epsilon x conv_out_gain."""
    sd["conv_out.weight"] = sd["conv_out.weight"] * conv_out_gain
    sd["conv_out.bias"] = sd["conv_out.bias"] * conv_out_gain
    return sd


def _decoder_conv_indices(sd):
    return sorted(int(k.split(".")[1]) for k in sd if k.startswith("decoder.") and k.endswith(".weight")
                  and sd[k].dim() == 4)


def trained_like_seg_decoder_(sd):
    """In place on a seg-AE state dict: the ConvTranspose2d layers ([cin, cout, 2, 2]) get four equal taps."""
    for i in _decoder_conv_indices(sd)[1:-1]:
        wt = sd[f"decoder.{i}.weight"]
        sd[f"decoder.{i}.weight"] = wt[:, :, :1, :1].expand_as(wt).clone()
    return sd


@torch.no_grad()
def fit_seg_head(features, out_channels=128, centroids=TRAINED_LIKE["head_centroids"], gain=TRAINED_LIKE["head_gain"],
                 offset=TRAINED_LIKE["head_offset"], iters=TRAINED_LIKE["head_iters"], stride=TRAINED_LIKE["head_stride"],
                 skip_labels=TRAINED_LIKE["head_skip_labels"]):
    """features: [1, C, H, W] f32, the input of the decoder's last conv (GroupNorm + SiLU output) for one teacher
    frame. Returns (weight [out, C, 3, 3], bias [out]) on the CPU:
        logit_c(x) = g * ( p_c . (box3x3(f)(x) - mu) - |p_c|^2 / 2 - offset * mean|p|^2 ),    g = gain / mean|p|^2
    with p_c the k-means centroids (Lloyd, `iters` rounds, started from evenly spaced samples: deterministic) of the
    centred, box-filtered features sampled every `stride`-th pixel. argmax_c is the nearest centroid. The offset sets
    how much of the plane a class claims through its sigmoid, i.e. the ratio of the merge's overlap test
    (argmax area / sigmoid area >= overlap_th, trainers_ldm_cond.py:1316-1321): with `offset` = 1 every class passes
    with a ratio > 1.3, far from the 0.5 threshold (at offset 0.5 the ratios sit between 0.3 and 0.7 and a 0.5 % id
    difference decides whether a whole segment survives). Centroid k becomes class ids[k], ids = 0..127 without
    `skip_labels`: the Cityscapes evaluator splits the labels 11..18 into connected components and counts every
    unmatched one as a false positive (cityscapes_pap_eval.py:96-105,166-174); a random UNet's texture gives such a
    label ~15 stray components per frame, so with 8 of 48 slots on thing ids the PQ is dominated by ~120 tiny false
    positives per frame whose number moves by 1-2 % under a 0.5 % id difference. One slot (11) stays a thing class.
    The classes without a centroid get bias -1e4."""
    import torch.nn.functional as F
    f = features.float()
    C = f.shape[1]
    fb = F.avg_pool2d(f[:1], 3, 1, 1, count_include_pad=True)      # = conv with w / 9 on all nine taps, zero padding
    X = fb[0].permute(1, 2, 0).reshape(-1, C)[::stride]
    mu = X.mean(0)
    Xc = X - mu
    P = Xc[torch.linspace(0, Xc.shape[0] - 1, centroids, device=Xc.device).long()].clone()
    for _ in range(iters):
        d = (Xc * Xc).sum(1, keepdim=True) - 2.0 * Xc @ P.T + (P * P).sum(1)[None]
        a = d.argmin(1)
        for k in range(centroids):
            m = a == k
            if bool(m.any()):
                P[k] = Xc[m].mean(0)
    n2 = (P * P).sum(1)
    g = gain / float(n2.mean())
    weight = torch.zeros((out_channels, C, 3, 3), dtype=torch.float32)
    bias = torch.full((out_channels,), -1.0e4, dtype=torch.float32)
    ids = torch.tensor([c for c in range(out_channels) if c not in set(skip_labels)][:centroids])
    weight[ids] = (g * P / 9.0).cpu()[:, :, None, None].expand(centroids, C, 3, 3)
    bias[ids] = (g * (-(P @ mu) - 0.5 * n2 - offset * n2.mean())).cpu()
    return weight, bias


def set_seg_head_(sd, weight, bias):
    """In place: the last decoder conv of a seg-AE state dict := (weight, bias)."""
    last = _decoder_conv_indices(sd)[-1]
    assert sd[f"decoder.{last}.weight"].shape == weight.shape, (sd[f"decoder.{last}.weight"].shape, weight.shape)
    sd[f"decoder.{last}.weight"], sd[f"decoder.{last}.bias"] = weight.clone(), bias.clone()
    return sd
