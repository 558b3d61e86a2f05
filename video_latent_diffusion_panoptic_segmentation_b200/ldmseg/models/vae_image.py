"""B200 host mirror of the RGB VAE ``GeneralVAEImage`` (ldmseg/models/vae.py:36-39 = diffusers ``AutoencoderKL``;
SURVEY section 8(f) rank 1): the step in front of the sampler that turns a frame into ``rgb_latents``.

Only the encoder half exists, as in the reference's sampling entry point (tools/main_ldm.py:138-140 replaces the
decoder by ``nn.Identity``). Same call surface: ``from_pretrained(path, subfolder="vae")``, ``set_scaling_factor``,
``encode(x).latent_dist.mode() / .sample()``, ``load_state_dict`` with diffusers key names. Every tensor op is a
kernel behind the C ABI (ops.py); PyTorch only owns the memory; there is no CPU fallback.

Data flow (NHWC bf16 activations, fp32 accumulate):
  conv_in 3 -> 128 fused with an affine map of the image (``2 * images - 1`` of encode_inputs)    few-channel kernel
  4 x DownEncoderBlock2D: 2 x [GroupNorm(eps 1e-6)+SiLU, conv3x3, GroupNorm+SiLU, conv3x3 + shortcut]   implicit GEMM;
      the 1x1 shortcut and the residual add ride in the second conv's launch sequence like the UNet's resnets
      Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + conv3x3 stride 2                         implicit GEMM (a_stride 2)
  mid block: resnet, single-head attention over the 512 channels, resnet
      one 512-wide head does not fit the fused flash kernels (their O accumulator lives in TMEM: 512 fp32 columns =
      all of it), so it runs unfused per image: S = Q K^T (GEMM, fp32 out), row softmax, O = P V (GEMM against V^T,
      which a GEMM with the roles of weights and activations swapped writes directly). It runs once per frame, not
      once per DDIM step: 115 GFLOP next to the sampler's 80 TFLOP.
  GroupNorm + SiLU, conv_out 512 -> 8 with ``quant_conv`` (1x1) folded into its weights -> fp32 NCHW moments
      (implicit GEMM with the planar fp32 epilogue, as the UNet's conv_out)
"""
import torch

from .unet import _Arena
from .vae import DiagonalGaussianDistribution, EncoderOutput
from ... import _lib as L
from ... import ops

bf16, f32 = torch.bfloat16, torch.float32

SD14_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, act_fn="silu", scaling_factor=0.18215)

_OLD_ATTN_KEYS = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


class GeneralVAEImage:
    def __init__(self, device="cuda", **config):
        cfg = dict(SD14_VAE_CONFIG)
        cfg.update(config)
        if cfg["act_fn"] != "silu":
            raise NotImplementedError("only act_fn='silu' (the SD VAE config) is built")
        for c in cfg["block_out_channels"]:
            if c % 64 != 0:
                raise NotImplementedError("block_out_channels must be multiples of 64 (tensor-core conv K blocks)")
        if 2 * cfg["latent_channels"] > 8:
            raise NotImplementedError("2 * latent_channels must be <= 8")
        from types import SimpleNamespace
        self.config = SimpleNamespace(**cfg)
        self.scaling_factor = cfg["scaling_factor"]
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.decoder = None  # tools/main_ldm.py:139 drops it
        self._sd, self._packed, self._plans = None, None, {}

    # ------------------------------------------------------------------ construction (main_ldm.py:138-140)
    @classmethod
    def from_pretrained(cls, path=None, subfolder="vae", cache_dir=None, device="cuda", state_dict=None, **config):
        """The reference loads the SD VAE from disk through diffusers; here the caller passes the diffusers-keyed
        ``state_dict`` or a ``torch.save``d one at ``path``."""
        net = cls(device=device, **config)
        if state_dict is not None:
            net.load_state_dict(state_dict)
        elif path is not None:
            data = torch.load(path, map_location="cpu")
            net.load_state_dict(data["vae_image"] if "vae_image" in data else data)
        return net

    def set_scaling_factor(self, scaling_factor):
        """vae.py:38-39"""
        self.scaling_factor = scaling_factor

    def load_state_dict(self, sd, strict=True):
        out = {}
        for k, v in sd.items():
            k = k.replace("module.", "")
            if k.startswith(("decoder.", "post_quant_conv.")):
                continue  # the decoder is dropped on this path
            if ".attentions." in k:
                parts = k.split(".")
                if parts[-2] in _OLD_ATTN_KEYS:  # diffusers < 0.15 AttentionBlock names
                    k = ".".join(parts[:-2] + [_OLD_ATTN_KEYS[parts[-2]], parts[-1]])
            out[k] = v.detach().to("cpu", f32)
        self._sd, self._packed, self._plans = out, None, {}
        return "<All keys matched successfully>"

    def state_dict(self):
        return dict(self._sd)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(self._sd.values())

    # ------------------------------------------------------------------ weights
    def _pack(self):
        sd, dev, cfg = self._sd, self.device, self.config
        P = {}

        def dv(t, dtype=f32):
            return t.to(dev, dtype).contiguous()

        def conv3(name):
            w = sd[name + ".weight"]
            return dv(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), bf16), dv(sd[name + ".bias"])

        def lin(name):
            w = sd[name + ".weight"]
            return dv(w.reshape(w.shape[0], -1), bf16), dv(sd[name + ".bias"])

        def norm(name):
            return dv(sd[name + ".weight"]), dv(sd[name + ".bias"])

        P["conv_in"] = (dv(ops.pack_small_cin_weight(sd["encoder.conv_in.weight"])), dv(sd["encoder.conv_in.bias"]))
        for r in [k[: -len(".norm1.weight")] for k in sd if k.endswith(".norm1.weight")]:
            P[r + ".norm1"], P[r + ".norm2"] = norm(r + ".norm1"), norm(r + ".norm2")
            P[r + ".conv1"], P[r + ".conv2"] = conv3(r + ".conv1"), conv3(r + ".conv2")
            if r + ".conv_shortcut.weight" in sd:
                P[r + ".conv_shortcut"] = lin(r + ".conv_shortcut")
        for k in sd:
            if k.endswith("downsamplers.0.conv.weight"):
                P[k[: -len(".weight")]] = conv3(k[: -len(".weight")])
        a = "encoder.mid_block.attentions.0"
        P[a + ".group_norm"] = norm(a + ".group_norm")
        P[a + ".to_q"], P[a + ".to_k"] = lin(a + ".to_q"), lin(a + ".to_k")
        P[a + ".to_v"] = dv(sd[a + ".to_v.weight"], bf16)
        # softmax rows sum to one: P (V + 1 b_v^T) = P V + b_v, so the value bias moves into the output projection
        wo, bo, bv = sd[a + ".to_out.0.weight"].double(), sd[a + ".to_out.0.bias"].double(), sd[a + ".to_v.bias"].double()
        P[a + ".to_out"] = (dv(wo.float(), bf16), dv((bo + wo @ bv).float()))
        P["conv_norm_out"] = norm("encoder.conv_norm_out")
        # quant_conv (1x1) folded into conv_out: W'[o] = sum_m Wq[o, m] W[m], b' = Wq b + bq
        wq = sd["quant_conv.weight"].double().reshape(sd["quant_conv.weight"].shape[0], -1)
        w, b = sd["encoder.conv_out.weight"].double(), sd["encoder.conv_out.bias"].double()
        wf = torch.einsum("om,mckl->ockl", wq, w)  # [2*latent, C, 3, 3] -> tap-major rows, zero-padded to 8 outputs
        bfold = wq @ b + sd["quant_conv.bias"].double()
        wop = torch.zeros((8, 9 * wf.shape[1]), dtype=torch.float64)
        wop[: wf.shape[0]] = wf.permute(0, 2, 3, 1).reshape(wf.shape[0], -1)
        bop = torch.zeros((8,), dtype=torch.float64)
        bop[: wf.shape[0]] = bfold
        P["conv_out"] = (dv(wop.float(), bf16), dv(bop.float()))
        self._packed = P

    # ------------------------------------------------------------------ plan
    def _build_plan(self, B, H, W):
        if self._packed is None:
            self._pack()
        P, cfg, dev = self._packed, self.config, self.device
        boc, groups, nres = list(cfg.block_out_channels), cfg.norm_num_groups, cfg.layers_per_block
        arena = _Arena(dev)
        plan = []

        def add(fn, *a, **k):
            plan.append((fn, a, k))

        from types import SimpleNamespace
        st = SimpleNamespace()
        st.image = torch.zeros((B, cfg.in_channels, H, W), dtype=f32, device=dev)
        st.affine = [1.0, 0.0]  # (scale, shift) of the image samples, read when the plan runs
        st.gn_stats = ops.gn_scratch(B, groups, dev)

        def resnet(name, x, h, w):
            cin, cout = x.shape[-1], P[name + ".conv1"][0].shape[0]
            t0 = arena.alloc((B, h, w, cin))
            add(ops.groupnorm, x, *P[name + ".norm1"], t0, st.gn_stats, groups=groups, eps=1e-6, silu=True)
            t1 = arena.alloc((B, h, w, cout))
            add(ops.gemm, t0, P[name + ".conv1"][0], t1, taps=9, bias=P[name + ".conv1"][1])
            arena.release(t0)
            t2 = arena.alloc((B, h, w, cout))
            add(ops.groupnorm, t1, *P[name + ".norm2"], t2, st.gn_stats, groups=groups, eps=1e-6, silu=True)
            arena.release(t1)
            if name + ".conv_shortcut" in P:
                sc = arena.alloc((B, h, w, cout))
                add(ops.gemm, x, P[name + ".conv_shortcut"][0], sc, taps=1, bias=P[name + ".conv_shortcut"][1])
            else:
                sc = x
            out = arena.alloc((B, h, w, cout))
            add(ops.gemm, t2, P[name + ".conv2"][0], out, taps=9, bias=P[name + ".conv2"][1],
                residual=sc.view(B * h * w, cout))
            arena.release(t2)
            if sc is not x:
                arena.release(sc)
            return out

        x = arena.alloc((B, H, W, boc[0]))
        st.conv_in_slot = len(plan)
        add(self._conv_in, st, x)
        h, w = H, W
        for i, co in enumerate(boc):
            for j in range(nres):
                y = resnet(f"encoder.down_blocks.{i}.resnets.{j}", x, h, w)
                arena.release(x)
                x = y
            if i < len(boc) - 1:
                oh, ow = (h - 2) // 2 + 1, (w - 2) // 2 + 1
                y = arena.alloc((B, oh, ow, co))
                wd_, bd_ = P[f"encoder.down_blocks.{i}.downsamplers.0.conv"]
                # F.pad(x, (0, 1, 0, 1)) + Conv2d(3x3, stride 2, padding 0): the A map steps by two pixels and its
                # out-of-range row / column reads as zero (at 384x1248 the im2col buffer of the first one was 2.2 GB)
                add(ops.gemm, x, wd_, y, taps=9, bias=bd_, a_stride=2, a_pad=0)
                arena.release(x)
                x, h, w = y, oh, ow

        y = resnet("encoder.mid_block.resnets.0", x, h, w)
        arena.release(x)
        x = y
        # single-head attention over C channels, unfused (see the module docstring)
        a = "encoder.mid_block.attentions.0"
        C, seq, M = x.shape[-1], h * w, B * h * w
        if seq % 32 != 0:
            raise NotImplementedError(f"VAE mid-block attention needs (H/8)*(W/8) % 32 == 0 (got {h}x{w})")
        t = arena.alloc((B, seq, C))
        add(ops.groupnorm, x, *P[a + ".group_norm"], t, st.gn_stats, groups=groups, eps=1e-6, silu=False)
        q, k = arena.alloc((B, seq, C)), arena.alloc((B, seq, C))
        add(ops.gemm, t.view(M, C), P[a + ".to_q"][0], q.view(M, C), bias=P[a + ".to_q"][1])
        add(ops.gemm, t.view(M, C), P[a + ".to_k"][0], k.view(M, C), bias=P[a + ".to_k"][1])
        vt = arena.alloc((C, seq))
        s = arena.alloc((seq, seq), f32)
        p = arena.alloc((seq, seq))
        o = arena.alloc((B, seq, C))
        for b in range(B):
            add(ops.gemm, P[a + ".to_v"], t[b], vt)                       # V^T = W_v X^T  [C, seq]
            add(ops.gemm, q[b], k[b], s, flags=L.LDM_GEMM_OUT_F32)        # S = Q K^T      [seq, seq] fp32
            add(ops.softmax_rows, s, p, float(C) ** -0.5)
            add(ops.gemm, p, vt, o[b])                                     # O = P V        [seq, C]
        for buf in (t, q, k, vt, s, p):
            arena.release(buf)
        y = arena.alloc((B, h, w, C))
        add(ops.gemm, o.view(M, C), P[a + ".to_out"][0], y.view(M, C), bias=P[a + ".to_out"][1],
            residual=x.view(M, C))
        arena.release(o)
        arena.release(x)
        x = y
        y = resnet("encoder.mid_block.resnets.1", x, h, w)
        arena.release(x)
        x = y

        t = arena.alloc((B, h, w, x.shape[-1]))
        add(ops.groupnorm, x, *P["conv_norm_out"], t, st.gn_stats, groups=groups, eps=1e-6, silu=True)
        st.moments = torch.empty((B, 2 * cfg.latent_channels, h, w), dtype=f32, device=dev)
        add(ops.gemm, t, P["conv_out"][0], st.moments, taps=9, bias=P["conv_out"][1], flags=L.LDM_GEMM_OUT_NCHW_F32,
            block_n=32, n_store=2 * cfg.latent_channels)
        st.plan, st.arena_bytes = plan, arena.total
        return st

    def _conv_in(self, st, out):
        P = self._packed
        ops.conv3x3_small_cin([st.image], P["conv_in"][0], P["conv_in"][1], out, scale=st.affine[0], shift=st.affine[1])

    # ------------------------------------------------------------------ encode
    @torch.no_grad()
    def encode_moments(self, x, scale=1.0, shift=0.0):
        """x f32 NCHW [B, 3, H, W] on the GPU, mapped to x*scale + shift on the fly -> moments f32 NCHW
        [B, 2*latent, H/8, W/8] (= quant_conv(encoder(x)) of diffusers AutoencoderKL.encode)."""
        if not x.is_cuda:
            raise L.LdmError("GeneralVAEImage.encode needs CUDA tensors: there is no CPU fallback")
        if self._sd is None:
            raise L.LdmError("GeneralVAEImage: no weights loaded")
        B, cin, H, W = x.shape
        if cin != self.config.in_channels:
            raise ValueError(f"image has {cin} channels, conv_in expects {self.config.in_channels}")
        key = (B, H, W)
        st = self._plans.get(key)
        if st is None:
            st = self._build_plan(B, H, W)
            self._plans = {key: st}  # keep one shape resident (full-resolution activations are GBs at B = 8)
        st.image.copy_(x)
        st.affine[0], st.affine[1] = float(scale), float(shift)
        for fn, a, k in st.plan:
            fn(*a, **k)
        return st.moments.clone()

    def encode(self, x, return_dict=True):
        posterior = DiagonalGaussianDistribution(self.encode_moments(x))
        if not return_dict:
            return (posterior,)
        return EncoderOutput(latent_dist=posterior)

    def decode(self, *a, **k):
        raise NotImplementedError("the RGB VAE decoder is dropped on the sampling path (tools/main_ldm.py:139)")
