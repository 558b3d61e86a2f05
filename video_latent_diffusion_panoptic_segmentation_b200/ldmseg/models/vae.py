"""B200 host mirror of the segmentation auto-encoder ``GeneralVAESeg`` (ldmseg/models/vae.py:42-307), decoder side.

``decode(z, interpolate=True)`` keeps the reference signature and returns NCHW fp32 logits; internally it is
  conv3x3 4->int (few-channel kernel, fused 1/scaling_factor)            vae.py:134
  [ConvTranspose2d(2,2) + LayerNorm2d + SiLU] x num_upscalers            vae.py:152-160  (one tcgen05 GEMM each,
                                                                          pixel-shuffle + LN + SiLU in the epilogue)
  GroupNorm + SiLU                                                        vae.py:163-164
  conv3x3 dim->num_classes (implicit GEMM, fp32 NHWC logits)              vae.py:165
  bilinear x interpolation_factor                                         vae.py:271
The sampler's integer tail consumes the NHWC logits *before* the bilinear step (``decode_nhwc``) and fuses the
up-sampling into the argmax kernel, so the 245 MB/frame full-resolution logits never hit HBM on that path.
The encoder / posteriors (vae.py:175-266) are SURVEY section 8(f) rank 3 ("next") and raise NotImplementedError.
"""
import torch

from .. import utils as U
from ... import _lib as L
from ... import ops

bf16, f32 = torch.bfloat16, torch.float32


class EncoderOutput(U.OutputDict):
    latent_dist: object


class VAEOutput(U.OutputDict):
    sample: torch.Tensor
    posterior: object


class GeneralVAESeg:
    def __init__(self, in_channels=3, int_channels=256, out_channels=19, block_out_channels=(32, 64, 128, 256),
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215, pretrained_path=None, encoder=None,
                 num_mid_blocks=0, num_latents=2, num_upscalers=1, upscale_channels=256, parametrization="gaussian",
                 fuse_rgb=False, resize_input=False, act_fn="none", clamp_output=False, freeze_codebook=False,
                 skip_encoder=False, device="cuda"):
        if num_mid_blocks > 0:
            raise NotImplementedError("num_mid_blocks > 0 is not the default config (base.yaml:29)")
        assert parametrization in ["gaussian", "discrete_gumbel_softmax", "discrete_codebook", "auto"]
        self.in_channels, self.int_channels, self.out_channels = in_channels, int_channels, out_channels
        self.latent_channels, self.norm_num_groups = latent_channels, norm_num_groups
        self.num_upscalers, self.upscale_channels = num_upscalers, upscale_channels
        self.scaling_factor = scaling_factor
        self.downsample_factor = 2 ** (len(block_out_channels) - 1)
        self.interpolation_factor = self.downsample_factor // (2 ** num_upscalers)
        self.parametrization, self.num_latents, self.act_fn, self.clamp_output = parametrization, num_latents, act_fn, clamp_output
        self.device = torch.device(device)
        self.dtype = torch.float32
        self._sd, self._packed, self._bufs = None, None, {}
        if pretrained_path is not None:
            self.load_pretrained(pretrained_path)

    # ------------------------------------------------------------------ weights (vae.py:117-122)
    def load_pretrained(self, pretrained_path):
        data = torch.load(pretrained_path, map_location="cpu")
        sd = {k.replace("module.", ""): v for k, v in data["vae"].items()}
        return self.load_state_dict(sd)

    def load_state_dict(self, sd, strict=True):
        self._sd = {k: v.detach().to("cpu", f32) for k, v in sd.items()}
        self._packed, self._bufs = None, {}
        return "<All keys matched successfully>"

    def state_dict(self):
        return dict(self._sd)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def _pack(self):
        sd, dev = self._sd, self.device
        P = {}
        P["in_w"], P["in_b"] = ops.pack_small_cin_weight(sd["decoder.0.weight"]).to(dev), sd["decoder.0.bias"].to(dev)
        idx = 2
        P["ups"] = []
        for _ in range(self.num_upscalers):
            w = sd[f"decoder.{idx}.weight"]  # ConvTranspose2d: [cin, cout, 2, 2]
            cout = w.shape[1]
            wp = w.permute(2, 3, 1, 0).reshape(4 * cout, w.shape[0]).contiguous().to(dev, bf16)  # [(dy,dx,co), ci]
            bp = sd[f"decoder.{idx}.bias"].repeat(4).contiguous().to(dev)
            g, b = sd[f"decoder.{idx + 1}.weight"].to(dev), sd[f"decoder.{idx + 1}.bias"].to(dev)
            P["ups"].append((wp, bp, g, b, cout))
            idx += 3
        P["gn"] = (sd[f"decoder.{idx}.weight"].to(dev), sd[f"decoder.{idx}.bias"].to(dev))
        w = sd[f"decoder.{idx + 2}.weight"]
        P["out_w"] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(dev, bf16)
        P["out_b"] = sd[f"decoder.{idx + 2}.bias"].to(dev)
        self._packed = P

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode_nhwc(self, z, scale=1.0):
        """z f32 NCHW [B,latent,h,w] (multiplied by `scale` on the fly) -> fp32 NHWC logits [B, s*h, s*w, out] with
        s = 2**num_upscalers, i.e. ``self.decoder(z * scale)`` of vae.py:269 in channels-last layout."""
        if self._packed is None:
            self._pack()
        if not z.is_cuda:
            raise L.LdmError("GeneralVAESeg.decode needs CUDA tensors: there is no CPU fallback")
        P = self._packed
        z = z.contiguous().float()
        B, _, h, w = z.shape
        key = (B, h, w)
        if key not in self._bufs:
            dev = self.device
            bufs = {"x0": torch.empty((B, h, w, self.int_channels), dtype=bf16, device=dev), "ups": []}
            hh, ww = h, w
            for (_, _, _, _, cout) in P["ups"]:
                hh, ww = 2 * hh, 2 * ww
                bufs["ups"].append(torch.empty((B, hh, ww, cout), dtype=bf16, device=dev))
            bufs["gn"] = torch.empty_like(bufs["ups"][-1])
            bufs["stats"] = ops.gn_scratch(B, self.norm_num_groups, dev)
            self._bufs = {key: bufs}  # keep one shape resident
        bufs = self._bufs[key]
        ops.conv3x3_small_cin([z], P["in_w"], P["in_b"], bufs["x0"], scale=float(scale))
        x = bufs["x0"]
        for (wp, bp, g, b, cout), out in zip(P["ups"], bufs["ups"]):
            ops.gemm(x, wp, out, bias=bp, flags=L.LDM_GEMM_CONVT_LN_SILU, block_n=cout, ln=(g, b, 1e-6))
            x = out
        ops.groupnorm(x, *P["gn"], bufs["gn"], bufs["stats"], groups=self.norm_num_groups, eps=1e-5, silu=True)
        logits = torch.empty(x.shape[:3] + (self.out_channels,), dtype=f32, device=self.device)
        ops.gemm(bufs["gn"], P["out_w"], logits, taps=9, bias=P["out_b"], flags=L.LDM_GEMM_OUT_F32)
        return logits

    @torch.no_grad()
    def decode(self, z, interpolate=True):
        """vae.py:268-272 -> NCHW fp32 [B, out, f*4h, f*4w] (f = interpolation_factor if interpolate else 1)."""
        logits = self.decode_nhwc(z)
        up = self.interpolation_factor if interpolate else 1
        B, H, W, C = logits.shape
        out = torch.empty((B, C, H * up, W * up), dtype=f32, device=self.device)
        ops.bilinear_up_nchw(logits, out, up)
        return out

    def encode(self, semseg):
        raise NotImplementedError("seg-AE encoder is SURVEY section 8(f) rank 3 (next); only decode() is on the path")

    def forward(self, sample, sample_posterior=True, return_dict=True, generator=None, rgb_sample=None,
                valid_mask=None):
        raise NotImplementedError("GeneralVAESeg.forward needs the encoder (SURVEY section 8(f) rank 3)")

    __call__ = forward
