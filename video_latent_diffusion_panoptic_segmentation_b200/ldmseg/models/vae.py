"""B200 host mirror of the segmentation auto-encoder ``GeneralVAESeg`` (ldmseg/models/vae.py:42-307).

``decode(z, interpolate=True)`` keeps the reference signature and returns NCHW fp32 logits; internally it is
  conv3x3 4->int (few-channel kernel, fused 1/scaling_factor)            vae.py:134
  [ConvTranspose2d(2,2) + LayerNorm2d + SiLU] x num_upscalers            vae.py:152-160  (one tcgen05 GEMM each,
                                                                          pixel-shuffle + LN + SiLU in the epilogue)
  GroupNorm + SiLU                                                        vae.py:163-164
  conv3x3 dim->num_classes (implicit GEMM, fp32 NHWC logits)              vae.py:165
  bilinear x interpolation_factor                                         vae.py:271
The sampler's integer tail consumes the NHWC logits *before* the bilinear step (``decode_nhwc``) and fuses the
up-sampling into the argmax kernel, so the 245 MB/frame full-resolution logits never hit HBM on that path.
``encode(semseg)`` (vae.py:175-266, default topology: no resize_input / skip_encoder /
mid blocks, gaussian parametrisation) runs the bit planes through
  conv3x3 in->c0 + SiLU (few-channel kernel)                              vae.py:191-195
  [conv3x3 ci->ci ; conv3x3 stride 2 ci->ci+1 + SiLU] per level           vae.py:199-208  (implicit GEMMs; stride 2 = a_stride)
  conv3x3 c_last->int ; GroupNorm(eps 1e-6) + SiLU ; conv3x3 int->2*latent  vae.py:215-237
and returns ``EncoderOutput(latent_dist=DiagonalGaussianDistribution)`` (vae.py:371-425). Channel counts below 64 are
zero-padded to 64 in the packed weights (the tensor-core conv wants K blocks of 64 channels).
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


class RangeDict(U.OutputDict):
    """vae.py:22-24"""
    min: torch.Tensor
    max: torch.Tensor


class DiagonalGaussianDistribution:
    """vae.py:371-425 (mean | logvar halves of the encoder's moments; clamp, activation, mode / sample / kl)."""

    def __init__(self, parameters, clamp_output=False, act_fn="none"):
        self.parameters = parameters
        if clamp_output:
            parameters = torch.clamp(parameters, -5.0, 5.0)
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.mean = self.to_range(self.mean, act_fn)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        self.clamp_output, self.act_fn = clamp_output, act_fn

    @staticmethod
    def to_range(x, act_fn):
        if act_fn == "sigmoid":
            return 2 * torch.sigmoid(x) - 1
        if act_fn == "tanh":
            return torch.tanh(x)
        if act_fn == "clip":
            return torch.clamp(x, -1, 1)
        if act_fn == "none":
            return x
        raise NotImplementedError

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def kl(self):
        return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])

    def get_range(self):
        return RangeDict(min=self.mean.min(), max=self.mean.max())

    def __str__(self) -> str:
        return (f"DiagonalGaussianDistribution(mean={self.mean}, var={self.var}, "
                f"clamp_output={self.clamp_output}, act_fn={self.act_fn})")


def _pad64(c):
    return (c + 63) // 64 * 64


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
        self.block_out_channels = tuple(block_out_channels)
        self.enc_in_channels = in_channels + (3 if fuse_rgb else 0)
        self._enc_unbuilt = ("resize_input" if resize_input else "") or ("skip_encoder" if skip_encoder else "") or \
                            ("external encoder" if encoder is not None else "")
        self.scaling_factor = scaling_factor
        self.downsample_factor = 2 ** (len(block_out_channels) - 1)
        self.interpolation_factor = self.downsample_factor // (2 ** num_upscalers)
        self.parametrization, self.num_latents, self.act_fn, self.clamp_output = parametrization, num_latents, act_fn, clamp_output
        self.device = torch.device(device)
        self.dtype = torch.float32
        self._sd, self._packed, self._bufs = None, None, {}
        self._enc_packed, self._enc_bufs = None, {}
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
        self._enc_packed, self._enc_bufs = None, {}
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
    def decode_features(self, z, scale=1.0):
        """The input of the decoder's last conv (GroupNorm + SiLU output, vae.py:163-164) as f32 NCHW [B, dim, 4h, 4w]:
        what unet_init.fit_seg_head fits the synthetic classifier head to (not on the sampling path)."""
        self.decode_nhwc(z, scale)
        return self._bufs[(z.shape[0], z.shape[2], z.shape[3])]["gn"].float().permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def decode(self, z, interpolate=True):
        """vae.py:268-272 -> NCHW fp32 [B, out, f*4h, f*4w] (f = interpolation_factor if interpolate else 1)."""
        logits = self.decode_nhwc(z)
        up = self.interpolation_factor if interpolate else 1
        B, H, W, C = logits.shape
        out = torch.empty((B, C, H * up, W * up), dtype=f32, device=self.device)
        ops.bilinear_up_nchw(logits, out, up)
        return out

    # ------------------------------------------------------------------ encode (vae.py:175-266)
    def _pack_encoder(self):
        sd, dev, boc = self._sd, self.device, self.block_out_channels

        def conv_w(name, cin_pad, cout_pad):  # [cout, cin, 3, 3] -> bf16 [cout_pad, 9 * cin_pad], tap-major / channel-minor
            w, b = sd[name + ".weight"], sd[name + ".bias"]
            wp = torch.zeros((cout_pad, 3, 3, cin_pad), dtype=f32)
            wp[: w.shape[0], :, :, : w.shape[1]] = w.permute(0, 2, 3, 1)
            bp = torch.zeros(cout_pad, dtype=f32)
            bp[: b.shape[0]] = b
            return wp.reshape(cout_pad, -1).contiguous().to(dev, bf16), bp.to(dev)

        P = {}
        w0, b0 = sd["encoder.0.weight"], sd["encoder.0.bias"]
        c0p = _pad64(boc[0])
        w0p = torch.zeros((c0p,) + tuple(w0.shape[1:]), dtype=f32)
        w0p[: w0.shape[0]] = w0
        b0p = torch.zeros(c0p, dtype=f32)
        b0p[: b0.shape[0]] = b0
        P["in_w"], P["in_b"] = ops.pack_small_cin_weight(w0p).to(dev), b0p.to(dev)
        idx, P["down"] = 2, []
        for i in range(len(boc) - 1):
            cip, cop = _pad64(boc[i]), _pad64(boc[i + 1])
            P["down"].append((conv_w(f"encoder.{idx}", cip, cip), conv_w(f"encoder.{idx + 1}", cip, cop), cip, cop))
            idx += 3
        P["mid"] = conv_w(f"encoder.{idx}", _pad64(boc[-1]), self.int_channels)
        P["gn"] = (sd[f"encoder.{idx + 2}.weight"].to(dev), sd[f"encoder.{idx + 2}.bias"].to(dev))
        P["out_w"], P["out_b"] = sd[f"encoder.{idx + 4}.weight"].contiguous().to(dev), sd[f"encoder.{idx + 4}.bias"].to(dev)
        self._enc_packed = P

    @torch.no_grad()
    def encode_moments(self, semseg):
        """semseg f32 NCHW [B, in_channels, H, W] on the GPU -> moments f32 NCHW [B, 2*latent, H/f, W/f]."""
        if self._enc_unbuilt:
            raise NotImplementedError(f"GeneralVAESeg.encode: the {self._enc_unbuilt} variant is not built")
        if self.int_channels % 64 != 0 or 2 * self.latent_channels > 8:
            raise NotImplementedError("GeneralVAESeg.encode: int_channels must be a multiple of 64 and 2*latent <= 8")
        if not semseg.is_cuda:
            raise L.LdmError("GeneralVAESeg.encode needs CUDA tensors: there is no CPU fallback")
        if self._enc_packed is None:
            self._pack_encoder()
        P, dev = self._enc_packed, self.device
        x_in = semseg.contiguous().float()
        B, cin, H, W = x_in.shape
        if cin != self.enc_in_channels:
            raise L.LdmError(f"GeneralVAESeg.encode: {cin} input channels, expected {self.enc_in_channels}")
        key = (B, H, W)
        if key not in self._enc_bufs:
            bufs = {"x0": torch.empty((B, H, W, P["in_b"].numel()), dtype=bf16, device=dev), "lv": []}
            h, w = H, W
            for (_, _, cip, cop) in P["down"]:
                oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                bufs["lv"].append((torch.empty((B, h, w, cip), dtype=bf16, device=dev),
                                   torch.empty((B, oh, ow, cop), dtype=bf16, device=dev)))
                h, w = oh, ow
            bufs["mid"] = torch.empty((B, h, w, self.int_channels), dtype=bf16, device=dev)
            bufs["gn"] = torch.empty_like(bufs["mid"])
            bufs["stats"] = ops.gn_scratch(B, self.norm_num_groups, dev)
            self._enc_bufs = {key: bufs}
        bufs = self._enc_bufs[key]
        x = ops.conv3x3_small_cin([x_in], P["in_w"], P["in_b"], bufs["x0"], silu=True)
        for ((wa, ba), (wb, bb), cip, cop), (ta, tb) in zip(P["down"], bufs["lv"]):
            ops.gemm(x, wa, ta, taps=9, bias=ba)
            # Conv2d(3x3, stride 2, padding 1) + SiLU as an implicit GEMM (element-strided A map, no im2col buffer)
            ops.gemm(ta, wb, tb, taps=9, bias=bb, flags=L.LDM_GEMM_SILU, a_stride=2, a_pad=1)
            x = tb
        ops.gemm(x, P["mid"][0], bufs["mid"], taps=9, bias=P["mid"][1])
        ops.groupnorm(bufs["mid"], *P["gn"], bufs["gn"], bufs["stats"], groups=self.norm_num_groups, eps=1e-6, silu=True)
        moments = torch.empty((B, P["out_b"].numel(), x.shape[1], x.shape[2]), dtype=f32, device=dev)
        ops.conv_out(bufs["gn"], P["out_w"], P["out_b"], moments)
        return moments

    def encode(self, semseg):
        moments = self.encode_moments(semseg)
        if self.parametrization != "gaussian":
            raise NotImplementedError(f"GeneralVAESeg.encode: parametrization {self.parametrization!r} is not built "
                                      "(base.yaml uses 'gaussian')")
        return EncoderOutput(latent_dist=DiagonalGaussianDistribution(moments, clamp_output=self.clamp_output,
                                                                      act_fn=self.act_fn))

    def forward(self, sample, sample_posterior=True, return_dict=True, generator=None, rgb_sample=None,
                valid_mask=None):
        """vae.py:274-307: encode -> sample / mode of the posterior -> (mask) -> decode without the bilinear step."""
        x = sample if rgb_sample is None else torch.cat([sample, rgb_sample], dim=1)
        posterior = self.encode(x).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        if valid_mask is not None:
            z = z * valid_mask[:, None]
        dec = self.decode(z, interpolate=False)
        if not return_dict:
            return (dec,)
        return VAEOutput(sample=dec, posterior=posterior)

    __call__ = forward
