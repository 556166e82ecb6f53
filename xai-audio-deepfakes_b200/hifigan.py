"""HiFi-GAN V1 generator on tcgen05 (the vocoder hifigan.py:106-110,180 runs through SpeechBrain).

SpeechBrain is not vendored in the reference and not installed here, and the checkpoint
(``speechbrain/tts-hifigan-libritts-16kHz``) is unreachable offline, so the architecture is restated from
the published V1 design that SpeechBrain's ``HifiganGenerator`` implements: conv_pre(80->512, k7);
4 x [LeakyReLU(0.1) -> ConvTranspose1d (x8, x8, x2, x2; k16, k16, k4, k4) -> mean of 3 ResBlock1
(k 3/7/11, dilations 1/3/5, two convs each)]; LeakyReLU(0.01) -> conv_post(32->1, k7) -> tanh;
``inference_padding`` = 5 replicate-padded mel frames per side.  conv_pre, the ResBlock convs and conv_post are
``speechbrain.nnet.CNN.Conv1d`` modules, whose "same" padding is REFLECT by default (``padding_mode="reflect"``); the
transposed convs pad with zeros.  ``HifiganConfig.pad_reflect = True`` (the default) follows that; ``False`` is the
original HiFi-GAN's zero padding.  Weights here are seeded random (N(0, 0.01^2), like the original init) with
weight-norm already folded; a SpeechBrain checkpoint loads through ``load_speechbrain_state_dict`` (weight_g / weight_v
folded, ``.conv.`` wrapper names stripped).  **Parity for this module is unpinned by the
reference** (see oracle/vocoder.py); it is checked against a torch-fp32 restatement with the same weights.

Execution: activations are channels-last bf16 [B][L][C]; every conv is ``adv_conv1d_bf16`` - an implicit
GEMM (M = positions, N = C_out, K = taps*C_in) on tcgen05 with fp32 TMEM accumulation, LeakyReLU fused
into the operand gather, bias / residual fused into the epilogue.  A ConvTranspose1d with stride s is the
same kernel with 3 taps and N = s*C_out phase-stacked weights: its channels-last output [B][L][s*C_out]
*is* the upsampled tensor [B][L*s][C_out].
"""
from __future__ import annotations


import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

LRELU_SLOPE = 0.1


class HifiganConfig:
    in_channels = 80
    upsample_initial_channel = 512
    upsample_factors = (8, 8, 2, 2)
    upsample_kernel_sizes = (16, 16, 4, 4)
    resblock_kernel_sizes = (3, 7, 11)
    resblock_dilation_sizes = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    inference_padding = 5
    conv_post_kernel = 7
    pad_reflect = True   # SpeechBrain's Conv1d default ("same" padding with padding_mode="reflect"); False = zeros
    halo = 32            # rows of materialised reflection either side of an activation (>= 5 * (11 - 1) / 2 = 25)


class HifiganConfigZeroPad(HifiganConfig):
    """The original HiFi-GAN's zero "same" padding (round-1 default; TMA out-of-bounds fill, no halo rows)."""
    pad_reflect = False


def init_weights(cfg=HifiganConfig, seed=0, std=0.01):
    """Seeded fp32 master weights in torch layout (Conv1d [Cout,Cin,k], ConvTranspose1d [Cin,Cout,k])."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: std * torch.randn(*s, generator=g)
    W = {}
    ch = cfg.upsample_initial_channel
    W["conv_pre.weight"], W["conv_pre.bias"] = rn(ch, cfg.in_channels, 7), rn(ch)
    for i, (s, k) in enumerate(zip(cfg.upsample_factors, cfg.upsample_kernel_sizes)):
        cout = ch // 2
        W[f"ups.{i}.weight"], W[f"ups.{i}.bias"] = rn(ch, cout, k), rn(cout)
        for j, (ks, dils) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            for d in range(len(dils)):
                r = f"resblocks.{i * len(cfg.resblock_kernel_sizes) + j}"
                W[f"{r}.convs1.{d}.weight"], W[f"{r}.convs1.{d}.bias"] = rn(cout, cout, ks), rn(cout)
                W[f"{r}.convs2.{d}.weight"], W[f"{r}.convs2.{d}.bias"] = rn(cout, cout, ks), rn(cout)
        ch = cout
    W["conv_post.weight"], W["conv_post.bias"] = rn(1, ch, cfg.conv_post_kernel), rn(1)
    return W


def fold_weight_norm(g, v, dim=0):
    """weight = g * v / ||v|| (norm over every dim but ``dim``), as torch.nn.utils.weight_norm stores it."""
    dims = [d for d in range(v.dim()) if d != dim]
    return g * v / torch.linalg.vector_norm(v, dim=dims, keepdim=True)


def load_speechbrain_state_dict(state_dict):
    """SpeechBrain ``HifiganGenerator.state_dict()`` (the ``generator.ckpt`` of tts-hifigan-libritts-16kHz) -> the flat
    weight dict this module consumes.  SpeechBrain wraps every layer (``conv_pre.conv.*``, ``ups.N.conv.*``,
    ``resblocks.N.convs1.M.conv.*``, ``conv_post.conv.*``) and applies ``torch.nn.utils.weight_norm``, which stores
    ``weight_g`` / ``weight_v`` (or ``parametrizations.weight.original0 / original1`` with the newer API); both are
    folded to plain weights here.  Raises KeyError on a checkpoint that lacks a layer."""
    sd = {k.replace(".conv.", "."): v.detach().float().cpu() for k, v in state_dict.items()}
    W = {}
    layers = sorted({k.rsplit(".", 1)[0].replace(".parametrizations.weight", "") for k in sd
                     if k.split(".")[0] in ("conv_pre", "ups", "resblocks", "conv_post")})
    for name in layers:
        if f"{name}.weight" in sd:
            w = sd[f"{name}.weight"]
        elif f"{name}.weight_g" in sd:
            w = fold_weight_norm(sd[f"{name}.weight_g"], sd[f"{name}.weight_v"])
        elif f"{name}.parametrizations.weight.original0" in sd:
            w = fold_weight_norm(sd[f"{name}.parametrizations.weight.original0"],
                                 sd[f"{name}.parametrizations.weight.original1"])
        else:
            continue
        W[f"{name}.weight"] = w
        W[f"{name}.bias"] = sd[f"{name}.bias"] if f"{name}.bias" in sd else torch.zeros(w.shape[1] if name.startswith("ups") else w.shape[0])
    for need in ("conv_pre.weight", "ups.0.weight", "resblocks.0.convs1.0.weight", "conv_post.weight"):
        if need not in W:
            raise KeyError(f"not a HifiganGenerator state_dict: {need} missing")
    return W


def _gemm_weight(w):
    """Conv1d weight [Cout, Cin, k] -> bf16 [Cout][Kpad], K index = tap*Cin + ci, Kpad multiple of 64."""
    cout, cin, k = w.shape
    m = w.permute(0, 2, 1).reshape(cout, k * cin)
    kpad = (k * cin + 63) // 64 * 64
    out = torch.zeros(cout, kpad)
    out[:, :k * cin] = m
    return out.to(torch.bfloat16).contiguous(), kpad


def _transposed_as_conv(w, bias, stride):
    """ConvTranspose1d weight [Cin, Cout, k] (padding (k-stride)//2) -> equivalent Conv1d weight
    [stride*Cout, Cin, taps] over the un-upsampled input: output phase r of position q reads
    in[q + o] * w[:, :, kk] for every kk with kk = r + p - o*stride."""
    cin, cout, k = w.shape
    p = (k - stride) // 2
    offs = [(r + p - kk) // stride for r in range(stride) for kk in range(k) if (r + p - kk) % stride == 0]
    reach = max(abs(o) for o in offs)
    taps = 2 * reach + 1
    wc = torch.zeros(stride * cout, cin, taps)
    for r in range(stride):
        for kk in range(k):
            if (r + p - kk) % stride == 0:
                o = (r + p - kk) // stride
                wc[r * cout:(r + 1) * cout, :, o + reach] = w[:, :, kk].t()
    return wc, bias.repeat(stride)


class _Layer:
    def __init__(self, w, bias, dil, dev):
        self.cout, self.cin, self.taps = w.shape
        self.dil = dil
        gw, self.kpad = _gemm_weight(w)
        self.w = gw.to(dev)
        # unpadded [Cout][taps*Cin] for the TMA pipeline (its tensor map walks K in blocks of 64 / 32 channels)
        self.tma_ok = (self.cin == 32 or self.cin % 64 == 0) and self.cout % 32 == 0
        self.w_tma = (w.permute(0, 2, 1).reshape(self.cout, -1).to(torch.bfloat16).contiguous().to(dev)
                      if self.tma_ok else None)
        self.bias = bias.float().to(dev).contiguous()


class HifiganGenerator:
    """``decode_batch(mel[B,80,T]) -> waveform [B,1,(T+10)*256]`` (SpeechBrain HIFIGAN.decode_batch).

    ``pipeline="tma"`` (default): TMA-fed warp-specialised persistent conv kernels; LeakyReLU is applied by the
    producing layer's epilogue.  With reflect padding (the default, see the module docstring) activations carry
    ``cfg.halo`` rows of materialised reflection either side (``adv_halo_fix_bf16``), so the same kernels - whose only
    padding is the TMA unit's zero fill - compute SpeechBrain's reflect "same" convs; the MRF average
    (``adv_avg_relayout_bf16``) switches to the zero halo the next transposed conv needs.  ``pipeline="gather"``:
    first-generation kernel (operands gathered with ordinary loads, activation and reflection on load)."""

    def __init__(self, weights=None, cfg=HifiganConfig, device=None, seed=0, pipeline=None, fuse="auto", epilogue="tma"):
        self.cfg = cfg
        # epilogue of the slab conv kernels: "tma" = output blocks staged in shared memory and written by the TMA engine,
        # "direct" = per-thread 16-byte stores, "auto" = each distinct layer shape is timed both ways on first use (outside
        # graph capture) and keeps the faster one.  Measured at 256 clips: the TMA form wins 17 - 37 % on the layers that
        # wait on L1 / HBM (128 channels and below, residual + two outputs) and loses 3 - 9 % on the tensor-bound 256-channel
        # and 11-tap layers, whose four epilogue warps have no slack (profiles/r02zh_vocoder_launches_b256.csv); whole
        # generator: "direct" 3 726, "auto" 3 957 (a layer timed alone, L2-warm, is not the layer in sequence), "tma" 4 046
        # clips/s - the default.
        self.epilogue = epilogue
        self._epi = {}
        # fused residual units on the narrow stages: "auto" = where measured faster, "always" = wherever the
        # kernel supports the shape, "never" = every conv is its own launch
        self.fuse = {True: "auto", False: "never"}.get(fuse, fuse)
        self.dev = device or ops._dev()
        self.pipeline = pipeline or "tma"
        if cfg.pad_reflect and any(cfg.halo % s for s in cfg.upsample_factors):
            raise ValueError("cfg.halo must be a multiple of every upsample factor")
        W = weights if weights is not None else init_weights(cfg, seed)
        self.master = W
        L = lambda name, dil=1: _Layer(W[name + ".weight"], W[name + ".bias"], dil, self.dev)
        self.conv_pre = L("conv_pre")
        self.ups, self.blocks = [], []
        nk = len(cfg.resblock_kernel_sizes)
        for i, s in enumerate(cfg.upsample_factors):
            wc, bc = _transposed_as_conv(W[f"ups.{i}.weight"], W[f"ups.{i}.bias"], s)
            self.ups.append((_Layer(wc, bc, 1, self.dev), s))
            stage = []
            for j, dils in enumerate(cfg.resblock_dilation_sizes):
                r = f"resblocks.{i * nk + j}"
                stage.append([(L(f"{r}.convs1.{d}", dil), L(f"{r}.convs2.{d}", 1)) for d, dil in enumerate(dils)])
            self.blocks.append(stage)
        wp = W["conv_post.weight"]
        self.post_w = wp[0].t().contiguous().float().to(self.dev)      # [taps][C]
        self.post_b = W["conv_post.bias"].float().to(self.dev)
        self.launches = 0

    def _buf(self, B, L, C):
        return torch.empty((B, L, C), dtype=torch.bfloat16, device=self.dev)

    # first-generation kernel: operands gathered with plain loads, optional LeakyReLU on load
    def _conv(self, x, layer, B, L, pre_slope=1.0, resid=None, scale=1.0, reflect=None, want_raw=True, act_slope=None):
        out = self._buf(B, L, layer.cout) if want_raw else None
        act = self._buf(B, L, layer.cout) if act_slope is not None else None
        check(lib().adv_conv1d_bf16(ptr(x), ptr(layer.w), ptr(layer.bias), ptr(resid), ptr(out), ptr(act), B, L,
                                    layer.cin, layer.taps, layer.dil, layer.cout, layer.kpad,
                                    int(self.cfg.pad_reflect if reflect is None else reflect), float(pre_slope),
                                    float(act_slope if act_slope is not None else 1.0), float(scale), stream_ptr()),
              "adv_conv1d_bf16")
        self.launches += 1
        return (out, act) if act_slope is not None else out

    # production kernel: TMA + warp specialisation; input is used as stored
    def _conv_tma(self, x, layer, B, L, resid=None, want_raw=True, act_slope=None):
        out = self._buf(B, L, layer.cout) if want_raw else None
        act = self._buf(B, L, layer.cout) if act_slope is not None else None

        def launch():
            check(lib().adv_conv1d_bf16_tma(ptr(x), ptr(layer.w_tma), ptr(layer.bias), ptr(resid), ptr(out), ptr(act), B, L,
                                            layer.cin, layer.taps, layer.dil, layer.cout,
                                            float(act_slope if act_slope is not None else 1.0), 1.0, stream_ptr()),
                  "adv_conv1d_bf16_tma")

        if self.epilogue == "auto":
            key = (layer.cin, layer.cout, layer.taps, layer.dil, resid is not None, want_raw, act_slope is not None, B, L)
            mode = self._epi.get(key)
            if mode is None and not torch.cuda.is_current_stream_capturing():
                best = None
                for m in (0, 1):            # time the layer both ways (same inputs, same outputs)
                    lib().adv_set_conv_epilogue(m)
                    launch()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(3):
                        launch()
                    b.record()
                    b.synchronize()
                    t = a.elapsed_time(b)
                    if best is None or t < best[0]:
                        best = (t, m)
                mode = self._epi[key] = best[1]
            lib().adv_set_conv_epilogue(1 if mode is None else mode)
        elif self.epilogue in ("tma", "direct", "tma_st"):   # "tma_st": TMA stores, residual by per-thread loads (A/B)
            lib().adv_set_conv_epilogue({"direct": 0, "tma": 1, "tma_st": 2}[self.epilogue])
        launch()
        self.launches += 1
        return out, act

    def _avg3(self, outs, slope=1.0):
        o = torch.empty_like(outs[0])
        check(lib().adv_avg3_bf16(ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), outs[0].numel(), float(slope), ptr(o),
                                  stream_ptr()), "adv_avg3_bf16")
        return o

    def _stages_gather(self, o, B, L):
        for (up, s), stage in zip(self.ups, self.blocks):
            # transposed convs never reflect: taps that fall outside the input contribute nothing
            o = self._conv(o, up, B, L, pre_slope=LRELU_SLOPE, reflect=False)   # [B][L][s*Cout] == [B][L*s][Cout]
            L, ch = L * s, up.cout // s
            o = o.view(B, L, ch)
            outs = []
            for branch in stage:
                xb = o
                for c1, c2 in branch:
                    xt = self._conv(xb, c1, B, L, pre_slope=LRELU_SLOPE)
                    xb = self._conv(xt, c2, B, L, pre_slope=LRELU_SLOPE, resid=xb)
                outs.append(xb)
            o = self._avg3(outs)
        return o, L

    # fused residual unit (conv1 -> LeakyReLU -> conv2 -> + x in one kernel): narrow stages whose two weight sets
    # fit in shared memory next to the slabs
    @staticmethod
    def _fusable(c1, c2):
        if c1.cin != c1.cout or c1.cin not in (32, 64) or c2.dil != 1 or c1.taps != c2.taps:
            return False
        row, rows = c1.cin * 2, 128 + (c1.taps - 1) * c1.dil
        r1k = lambda v: (v + 1023) // 1024 * 1024
        smem = 2 * c1.taps * c1.cin * row + 3 * r1k(rows * row) + r1k((128 + c1.taps - 1) * row) + 1280
        return rows <= 256 and smem <= 227 * 1024

    @staticmethod
    def _fused_wins(c1):
        """Measured on B200 (64 x 4 s clips): the fused unit's per-tile chain of hand-offs (TMA -> LeakyReLU pass ->
        conv1 -> epilogue -> conv2 -> epilogue) takes ~4.5 us per tile per CTA whatever the tap count, so it beats
        two separate conv launches only where 4 CTAs share an SM (32 channels, 3 taps: 237 vs 334 us per unit); with
        3 or fewer resident CTAs the two-launch path is faster (7 taps: 385 vs 338 us; 64 channels: 374-487 vs
        330-460 us)."""
        row, rows = c1.cin * 2, 128 + (c1.taps - 1) * c1.dil
        r1k = lambda v: (v + 1023) // 1024 * 1024
        smem = 2 * c1.taps * c1.cin * row + 3 * r1k(rows * row) + r1k((128 + c1.taps - 1) * row) + 1280
        return 4 * (smem + 1024) <= 228 * 1024

    def _resunit(self, x_raw, c1, c2, B, L):
        out = self._buf(B, L, c1.cout)
        check(lib().adv_resunit_bf16(ptr(x_raw), ptr(c1.w_tma), ptr(c1.bias), ptr(c2.w_tma), ptr(c2.bias), ptr(out), B, L,
                                     c1.cin, c1.taps, c1.dil, LRELU_SLOPE, stream_ptr()), "adv_resunit_bf16")
        self.launches += 1
        return out

    def _stages_tma(self, o, B, L):
        """``o`` = conv_pre output, already LeakyReLU'd.  Every tensor an (unfused) conv consumes was written
        activated by its producer; raw copies exist only where a residual, a fused unit or the MRF average needs
        them."""
        n_stage = len(self.ups)
        o_act = o
        for si, ((up, s), stage) in enumerate(zip(self.ups, self.blocks)):
            fused = [self.fuse != "never" and all(self._fusable(c1, c2) and (self.fuse == "always" or self._fused_wins(c1))
                                                  for c1, c2 in branch) for branch in stage]
            x_raw, x_act = self._conv_tma(o_act, up, B, L, act_slope=None if all(fused) else LRELU_SLOPE)
            L, ch = L * s, up.cout // s
            x_raw = x_raw.view(B, L, ch)
            x_act = None if x_act is None else x_act.view(B, L, ch)
            outs = []
            for branch, fz in zip(stage, fused):
                xb_raw, xb_act = x_raw, x_act
                for d, (c1, c2) in enumerate(branch):
                    if fz:
                        xb_raw = self._resunit(xb_raw, c1, c2, B, L)
                        continue
                    _, xt_act = self._conv_tma(xb_act, c1, B, L, want_raw=False, act_slope=LRELU_SLOPE)
                    last = d == len(branch) - 1
                    xb_raw, xb_act = self._conv_tma(xt_act, c2, B, L, resid=xb_raw,
                                                    act_slope=None if last else LRELU_SLOPE)
                outs.append(xb_raw)
            if si == n_stage - 1:
                o = self._avg3(outs)                       # raw: conv_post applies its own LeakyReLU(0.01)
            else:
                o_act = self._avg3(outs, LRELU_SLOPE)      # next stage's transposed conv reads it activated
        return o, L

    def _halo_fix(self, x, B, L, C, mode=1):
        check(lib().adv_halo_fix_bf16(ptr(x), B, L, C, self.cfg.halo, mode, stream_ptr()), "adv_halo_fix_bf16")
        self.launches += 1

    def _relayout(self, srcs, B, L, C, h_in, h_out, slope=1.0, mode=0):
        out = self._buf(B, L + 2 * h_out, C)
        a, b, c = (list(srcs) + [None, None])[:3]
        check(lib().adv_avg_relayout_bf16(ptr(a), ptr(b), ptr(c), B, L, C, h_in, h_out, mode, float(slope), ptr(out),
                                          stream_ptr()), "adv_avg_relayout_bf16")
        self.launches += 1
        return out

    def _stages_tma_reflect(self, o_act, B, L):
        """Reflect-padded generator on the TMA kernels.  ``o_act`` = LeakyReLU(conv_pre) [B][L][C], contiguous.  Inside a
        stage every tensor is [B][L + 2 H][C] (H = cfg.halo); a conv runs over all L + 2 H rows and the halo rows of
        what it wrote for the NEXT conv are replaced by the reflection of the interior.  The transposed conv wants zeros
        outside its input, so a stage's input carries a zero halo of H / stride rows - which the stride turns into
        exactly H output rows."""
        H = self.cfg.halo
        n_stage = len(self.ups)
        h = H // self.ups[0][1]
        x_in = self._relayout([o_act], B, L, self.conv_pre.cout, 0, h)
        o = None
        for si, ((up, s), stage) in enumerate(zip(self.ups, self.blocks)):
            x_raw, x_act = self._conv_tma(x_in, up, B, L + 2 * h, act_slope=LRELU_SLOPE)
            L, ch = L * s, up.cout // s
            Lp = L + 2 * H
            x_raw, x_act = x_raw.view(B, Lp, ch), x_act.view(B, Lp, ch)
            self._halo_fix(x_act, B, L, ch)
            outs = []
            for branch in stage:
                xb_raw, xb_act = x_raw, x_act
                for d, (c1, c2) in enumerate(branch):
                    _, xt_act = self._conv_tma(xb_act, c1, B, Lp, want_raw=False, act_slope=LRELU_SLOPE)
                    self._halo_fix(xt_act, B, L, ch)
                    last = d == len(branch) - 1
                    xb_raw, xb_act = self._conv_tma(xt_act, c2, B, Lp, resid=xb_raw, act_slope=None if last else LRELU_SLOPE)
                    if xb_act is not None:
                        self._halo_fix(xb_act, B, L, ch)
                outs.append(xb_raw)
            if si == n_stage - 1:
                o = self._relayout(outs, B, L, ch, H, 0)                        # raw mean, no halo: conv_post reflects itself
            else:
                h = H // self.ups[si + 1][1]
                x_in = self._relayout(outs, B, L, ch, H, h, slope=LRELU_SLOPE)  # activated mean with the next stage's zero halo
        return o, L

    @torch.no_grad()
    def forward_padded(self, mel):
        """mel [B, 80, T] fp32 -> waveform [B, (T + 2*pad) * prod(upsample_factors)] fp32."""
        cfg = self.cfg
        mel = mel.to(self.dev, torch.float32).contiguous()
        B, C, T = mel.shape
        pad = cfg.inference_padding
        L = T + 2 * pad
        x = self._buf(B, L, C)
        check(lib().adv_mel_to_channels_last(ptr(mel), B, C, T, pad, C, ptr(x), stream_ptr()), "adv_mel_to_channels_last")
        if self.pipeline == "tma":   # conv_pre (C_in = 80) runs on the gather kernel and hands over lrelu(o)
            _, o = self._conv(x, self.conv_pre, B, L, want_raw=False, act_slope=LRELU_SLOPE)
            o, L = self._stages_tma_reflect(o, B, L) if cfg.pad_reflect else self._stages_tma(o, B, L)
        else:
            o, L = self._stages_gather(self._conv(x, self.conv_pre, B, L), B, L)
        wav = torch.empty((B, L), dtype=torch.float32, device=self.dev)
        check(lib().adv_post_conv_tanh(ptr(o), ptr(self.post_w), ptr(self.post_b), B, L, o.shape[2],
                                       self.post_w.shape[0], 0.01, int(cfg.pad_reflect), ptr(wav), stream_ptr()),
              "adv_post_conv_tanh")
        return wav

    def decode_batch(self, spectrogram):
        """[B, 80, T] (or [80, T]) -> [B, 1, samples] like SpeechBrain's HIFIGAN.decode_batch."""
        if spectrogram.dim() == 2:
            spectrogram = spectrogram.unsqueeze(0)
        return self.forward_padded(spectrogram).unsqueeze(1)

    @staticmethod
    def flops_per_clip(T, cfg=HifiganConfig):
        """Dense MACs*2 of the generator for a T-frame mel."""
        L = T + 2 * cfg.inference_padding
        ch = cfg.upsample_initial_channel
        fl = 2 * L * ch * cfg.in_channels * 7
        for s, k in zip(cfg.upsample_factors, cfg.upsample_kernel_sizes):
            fl += 2 * L * ch * (ch // 2) * k
            L, ch = L * s, ch // 2
            for ks, dils in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
                fl += 2 * len(dils) * 2 * L * ch * ch * ks
        return fl + 2 * L * ch * cfg.conv_post_kernel


def align_shift(ref_wav, deg_wav, method="fft"):
    """Arg-max lag of the cross-correlation the reference computes with a direct conv1d (hifigan.py:113-126):
    device int32 tensor holding ``argmax(conv1d(pad(ref, (P, P)), deg)) - P`` with ``P = len(deg)``.

    ``method="fft"`` (default): overlap-save through our own transform kernels - zero-padded STFT (n_fft 1024,
    hop 512) of deg in 512-tap blocks and of the padded ref, one multiply-accumulate launch over (output frame, bin),
    one iSTFT, one arg-max: O(N log N), 6 launches.  ``method="direct"``: the O(N^2) fp32 accumulation in the
    reference's own order of operations (adv_xcorr_shift).  Both return the first maximum; they can only differ when
    two lags tie to within fp32 round-off of the correlation peak."""
    dev = ops._dev()
    ref = ref_wav.reshape(-1).to(dev, torch.float32).contiguous()
    deg = deg_wav.reshape(-1).to(dev, torch.float32).contiguous()
    shift = torch.empty(1, dtype=torch.int32, device=dev)
    if method == "direct":
        nb = lib().adv_xcorr_blocks(ref.numel(), deg.numel())
        ws_val = torch.empty(nb, dtype=torch.float32, device=dev)
        ws_idx = torch.empty(nb, dtype=torch.int32, device=dev)
        check(lib().adv_xcorr_shift(ptr(ref), ref.numel(), ptr(deg), deg.numel(), ptr(ws_val), ptr(ws_idx), ptr(shift),
                                    stream_ptr()), "adv_xcorr_shift")
        return shift
    if method != "fft":
        raise ValueError("method must be 'fft' or 'direct'")
    cc = xcorr_curve(ref, deg)
    check(lib().adv_argmax_first(ptr(cc), cc.numel(), deg.numel(), ptr(shift), stream_ptr()), "adv_argmax_first")
    return shift


def xcorr_curve(ref_wav, deg_wav):
    """``conv1d(pad(ref, (P, P)), deg)`` (hifigan.py:119-122), all ``len(ref) + len(deg) + 1`` lags, by overlap-save
    through the transform kernels: zero-padded STFTs (n_fft 1024, hop 512) of deg in 512-tap blocks and of the padded
    ref, adv_xcorr_fd_mac over (output frame, bin), one iSTFT whose 512-tap synthesis window keeps exactly the valid
    half of every circular correlation.  Returns a device float32 view [n_lags]."""
    from ._lib import STFT_ZERO_PAD, get_plan
    dev = ops._dev()
    ref = ref_wav.reshape(-1).to(dev, torch.float32).contiguous()
    deg = deg_wav.reshape(-1).to(dev, torch.float32).contiguous()
    NF, H, BINS = 1024, 512, 513
    n_ref, P = ref.numel(), deg.numel()
    n_lags = n_ref + P + 1
    nb, nq = (P + H - 1) // H, (n_lags + H - 1) // H
    # deg blocks: block b = deg[512 b : 512 b + 512] sits in frame b + 1 under the centred 512-tap window
    x = torch.zeros((1, H * nb + H), dtype=torch.float32, device=dev)
    x[0, 256:256 + P] = deg
    # ref padded by P zeros in front (the reference's F.pad); frame t holds y[512 (t - 1) : 512 (t + 1)]
    y = torch.zeros((1, H * (nq + nb + 1)), dtype=torch.float32, device=dev)
    y[0, P:P + n_ref] = ref

    def zstft(sig, win):
        n = sig.shape[1]
        T = 1 + n // H
        plan = get_plan(NF, H, win, None, T, n, 0)
        X = torch.empty((1, T, BINS), dtype=torch.complex64, device=dev)
        check(lib().adv_stft_ex(plan.handle, ptr(sig), sig.stride(0), 1, ptr(X), None, None, STFT_ZERO_PAD, stream_ptr()),
              "adv_stft_ex")
        return X, T

    D, _ = zstft(x, H)
    R, t_r = zstft(y, NF)
    Z = torch.empty((1, nq + 1, BINS), dtype=torch.complex64, device=dev)
    check(lib().adv_xcorr_fd_mac(ptr(D), nb, ptr(R), t_r, ptr(Z), nq, BINS, stream_ptr()), "adv_xcorr_fd_mac")
    out = ops.istft(Z.transpose(1, 2), NF, H, H, length=n_lags + 256)  # cc[j] = out[j + 256]
    return out[0, 256:]


def align_waveforms(ref_wav, deg_wav):
    """Same contract as the reference's ``align_waveforms`` (hifigan.py:113-136): returns ``(ref_aligned,
    deg_aligned)``, views of shape [1, 1, n] trimmed to the common length after shifting by the cross-correlation
    arg-max.  The O(N^2) correlation runs on the GPU; only the shift (one int) is read back, as the reference does
    with ``.item()``."""
    ref_wav = ref_wav.view(1, 1, -1)
    deg_wav = deg_wav.view(1, 1, -1)
    shift = int(align_shift(ref_wav, deg_wav).item())
    if shift > 0:
        ref_aligned = ref_wav[..., shift:]
        deg_aligned = deg_wav[..., : ref_aligned.shape[-1]]
    else:
        deg_aligned = deg_wav[..., -shift:]
        ref_aligned = ref_wav[..., : deg_aligned.shape[-1]]
    n = min(ref_aligned.shape[-1], deg_aligned.shape[-1])
    return ref_aligned[..., :n], deg_aligned[..., :n]


def band_swapped_waveforms(s_ref, s_voc, n_fft=1024, hop_length=256, win_length=1024, band_width=1000, f_max=8000):
    """hifigan.py:188-225 for one aligned pair: hann STFT of both signals, every 1 kHz band of the vocoded spectrum
    swapped into the real one, hann iSTFT without ``length`` (``hop * (T - 1)`` samples).  Returns [n_bands, samples].
    Two STFT launches, ONE band-swap launch, ONE batched iSTFT launch instead of the reference's 8 x (clone + masked
    copy + istft)."""
    window = torch.hann_window(win_length)
    wav = torch.stack([s_ref.reshape(-1), s_voc.reshape(-1)]).to(ops._dev(), torch.float32)
    X, _, _ = ops.stft(wav, n_fft, hop_length, win_length, window=window, want_mag=False, want_phase=False)
    combined = ops.band_swap_all(X[0:1], X[1:2], band_width, f_max)
    return ops.istft(combined, n_fft, hop_length, win_length, length=None, window=window)
