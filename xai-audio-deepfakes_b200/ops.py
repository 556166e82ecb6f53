"""Tensor-level wrappers over the C ABI (torch is only the owner of device memory and streams).

Every function here launches hand-written sm_100a kernels from libaddvisor_sm100.so on the
current CUDA stream.  Inputs that live on the host are moved with ``.to(device)`` exactly like
the reference does (audioprocessor.py:90,98); nothing is computed on the CPU.
"""
from __future__ import annotations


import torch

from . import _lib
from ._lib import check, get_plan, lib, ptr, stream_ptr

MODES = {"log1p": _lib.MASK_LOG1P, "linear": _lib.MASK_LINEAR}
# bins outside a mask smaller than the spectrum: "drop" removes them from BOTH outputs - what the reference's crop of
# magnitude and phase to the mask's extent amounts to (LMAC_metrics.py:136-139, loss_function.py:36-41); "keep_irr"
# zero-extends the mask, i.e. out-of-mask bins pass to the masked-out branch with gain 1
OUTSIDE = {"drop": _lib.MASK_DROP_OUTSIDE, "keep_irr": 0}


def _mode_flags(mode, outside):
    return MODES[mode] | OUTSIDE[outside]


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _on_input_device(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: plans, launches, the stream and the outputs
    then live where the data is (host inputs keep going to the current device, like the reference's ``.to(device)``)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _f32_rows(x, name):
    """float32 CUDA tensor [B, n] with unit stride on the last dim."""
    if x.dtype != torch.float32:
        x = x.float()
    x = x.to(_dev(), non_blocking=True)
    if x.dim() != 2:
        raise ValueError(f"{name} must be 2-D")
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def _mask3(mask):
    if mask.dim() == 4 and mask.shape[1] == 1:  # UNet output [B,1,F,T] (addvisor.py:82)
        mask = mask[:, 0]
    if mask.dim() != 3:
        raise ValueError("mask must be [B,F,T] (or [B,1,F,T])")
    return mask.to(_dev(), torch.float32, non_blocking=True).contiguous()


# ----------------------------------------------------------------------------------------------
@_on_input_device
def stft(wave, n_fft, hop, win_length, window=None, want_mag=True, want_phase=True):
    """wave [B,n] -> (X [B,F,T] complex64 with torch.stft's frame-major strides, mag, phase)."""
    wave = _f32_rows(wave, "waveform")
    B, n = wave.shape
    T, Fb = 1 + n // hop, n_fft // 2 + 1
    plan = get_plan(n_fft, hop, win_length, window, T, n, 0)  # forward-only plan
    X = torch.empty((B, T, Fb), dtype=torch.complex64, device=wave.device)
    mag = torch.empty((B, T, Fb), dtype=torch.float32, device=wave.device) if want_mag else None
    ph = torch.empty((B, T, Fb), dtype=torch.float32, device=wave.device) if want_phase else None
    check(lib().adv_stft(plan.handle, ptr(wave), wave.stride(0), B, ptr(X), ptr(mag), ptr(ph), stream_ptr()),
          "adv_stft")
    tr = lambda t: None if t is None else t.transpose(1, 2)
    return tr(X), tr(mag), tr(ph)


def _spec_strides(spec):
    """(tensor, sb, st, sf) element strides of a complex [B,F,T] tensor, no copy when possible."""
    if spec.dtype != torch.complex64:
        spec = spec.to(torch.complex64)
    spec = spec.to(_dev(), non_blocking=True).resolve_conj()   # (X.conj() is lazy: data_ptr() would be the unconjugated data)
    return spec, spec.stride(0), spec.stride(2), spec.stride(1)


@_on_input_device
def istft(spec, n_fft, hop, win_length, length=None, window=None, return_stats=False):
    """spec complex [B,F,T] (any strides) -> wave [B, length]; length=None => hop*(T-1)."""
    spec, sb, st, sf = _spec_strides(spec)
    B, Fb, T = spec.shape
    if Fb != n_fft // 2 + 1:  # torch.istft raises on a wrong bin count
        raise RuntimeError(f"istft: expected {n_fft // 2 + 1} frequency bins, got {Fb}")
    n_out = int(length) if length is not None else hop * (T - 1)
    plan = get_plan(n_fft, hop, win_length, window, T, 0, n_out)
    out = torch.empty((B, n_out), dtype=torch.float32, device=spec.device)
    stats = None
    if return_stats:
        stats = torch.empty((B, plan.tiles_istft(B), 2), dtype=torch.float64, device=spec.device)
    check(lib().adv_istft(plan.handle, ptr(spec), sb, st, sf, B, ptr(out), ptr(stats), stream_ptr()), "adv_istft")
    return (out, stats) if return_stats else out


@_on_input_device
def normalize_(x, stats=None, width=2, col=0, out=None):
    """(x - mean) / (std_unbiased + 1e-7) per row (classifier_embedder.py:59-63); ``stats`` are the
    per-tile partial sums an upstream kernel already produced, else one extra pass computes them."""
    x = _f32_rows(x, "waveform")
    B, n = x.shape
    if x.stride(0) != n:
        x = x.contiguous()
    if stats is None:
        parts = lib().adv_row_stats_parts(n)
        stats = torch.empty((B, parts, 2), dtype=torch.float64, device=x.device)
        check(lib().adv_row_stats(ptr(x), B, n, ptr(stats), stream_ptr()), "adv_row_stats")
        width, col = 2, 0
    parts = stats.shape[1]
    out = torch.empty_like(x) if out is None else out
    check(lib().adv_normalize(ptr(x), ptr(out), B, n, ptr(stats), parts, width, col, stream_ptr()), "adv_normalize")
    return out


@_on_input_device
def explain(wave, mask, n_fft, hop, win_length, length=None, mode="log1p", window=None, normalize=False,
            out=None, outside="drop"):
    """Fused wave + mask -> (masked-in wave, masked-out wave) [B, length] (LMAC_metrics.py:136-157).
    ``out=(rel, irr, stats)`` reuses caller-owned buffers (stats: float64 [B, tiles, 4]).
    A mask smaller than the [F, T] grid (the U-Net's 512 x 248 against 513 x 249) covers its top-left corner;
    ``outside`` says what happens to the other bins: ``"drop"`` (default, the reference's crop: gone from both
    outputs) or ``"keep_irr"`` (zero extension: they stay in the masked-out waveform)."""
    wave = _f32_rows(wave, "waveform")
    mask = _mask3(mask)
    B, n = wave.shape
    if mask.shape[0] != B:
        raise ValueError("mask batch does not match the waveforms")
    T = 1 + n // hop
    n_out = int(length) if length is not None else hop * (T - 1)
    plan = get_plan(n_fft, hop, win_length, window, T, n, n_out)
    if out is not None:
        rel, irr, stats = out
    else:
        rel = torch.empty((B, n_out), dtype=torch.float32, device=wave.device)
        irr = torch.empty_like(rel)
        stats = torch.empty((B, plan.tiles(B), 4), dtype=torch.float64, device=wave.device) if normalize else None
    check(lib().adv_explain(plan.handle, ptr(wave), wave.stride(0), ptr(mask), mask.shape[1], mask.shape[2],
                            _mode_flags(mode, outside), B, ptr(rel), ptr(irr), ptr(stats), stream_ptr()), "adv_explain")
    if normalize:
        check(lib().adv_normalize_pair(ptr(rel), ptr(irr), B, n_out, ptr(stats), stats.shape[1], stream_ptr()),
              "adv_normalize_pair")
    return rel, irr


@_on_input_device
def normalize_pair_(rel, irr, stats):
    """In-place zero_mean_unit_var_norm of both explain outputs from the per-tile sums the explain kernel wrote
    (stats float64 [B, tiles, 4]); one launch."""
    B, n_out = rel.shape
    check(lib().adv_normalize_pair(ptr(rel), ptr(irr), B, n_out, ptr(stats), stats.shape[1], stream_ptr()),
          "adv_normalize_pair")
    return rel, irr


def normalize_pair_lmac_(rel, irr, stats, p, theta, q, is_logit=False, workspace=None, accumulate=False):
    """``normalize_pair_`` and ``lmac`` (sums only) of the same batch in one launch - the metric reduction rides as an
    extra CTA of the normaliser's grid.  Up to 1 024 logits; returns the float64 [6] sums tensor of the workspace."""
    B, n_out = rel.shape
    n = p.numel()
    if n > 1024 or not (p.is_contiguous() and theta.is_contiguous() and q.is_contiguous()) or p.dtype != torch.float32:
        normalize_pair_(rel, irr, stats)
        return lmac(p, theta, q, is_logit=is_logit, want_scores=False, workspace=workspace, accumulate=accumulate)[1]
    ws = workspace if workspace is not None and workspace.n == n else LmacWorkspace(n, rel.device)
    flags = (1 if is_logit else 0) | (2 if accumulate else 0)
    check(lib().adv_normalize_pair_lmac(ptr(rel), ptr(irr), B, n_out, ptr(stats), stats.shape[1], ptr(p), ptr(theta), ptr(q),
                                        n, flags, None, ptr(ws.sums), stream_ptr()), "adv_normalize_pair_lmac")
    return ws.sums


def explain_tiles(n_fft, hop, win_length, n, batch, length=None, window=None):
    """Tiles per clip the explain kernel will use (second dim of its ``stats`` buffer)."""
    T = 1 + n // hop
    n_out = int(length) if length is not None else hop * (T - 1)
    return get_plan(n_fft, hop, win_length, window, T, n, n_out).tiles(batch)


@_on_input_device
def explain_spec(spec, mask, n_fft, hop, win_length, length=None, mode="log1p", window=None, normalize=False,
                 outside="drop"):
    """Same as :func:`explain` starting from an STFT (complex [B,F,T])."""
    spec, sb, st, sf = _spec_strides(spec)
    mask = _mask3(mask)
    B, Fb, T = spec.shape
    if Fb != n_fft // 2 + 1:
        raise RuntimeError(f"istft: expected {n_fft // 2 + 1} frequency bins, got {Fb}")
    n_out = int(length) if length is not None else hop * (T - 1)
    plan = get_plan(n_fft, hop, win_length, window, T, 0, n_out)
    rel = torch.empty((B, n_out), dtype=torch.float32, device=spec.device)
    irr = torch.empty_like(rel)
    stats = torch.empty((B, plan.tiles(B), 4), dtype=torch.float64, device=spec.device) if normalize else None
    check(lib().adv_explain_spec(plan.handle, ptr(spec), sb, st, sf, ptr(mask), mask.shape[1], mask.shape[2],
                                 _mode_flags(mode, outside), B, ptr(rel), ptr(irr), ptr(stats), stream_ptr()), "adv_explain_spec")
    if normalize:
        check(lib().adv_normalize_pair(ptr(rel), ptr(irr), B, n_out, ptr(stats), stats.shape[1], stream_ptr()),
              "adv_normalize_pair")
    return rel, irr


class _ExplainLinearFn(torch.autograd.Function):
    """Differentiable linear mask path of the training loss (loss_function.py:36-47):
        rel = istft(m * X),  irr = istft((1 - m) * X).
    Forward = the fused explain kernel.  Backward: both maps are linear in m with the same adjoint, so
        dL/dm = c_f * Re(conj(X) * STFT0(w; (g_rel - g_irr) * inv_env))
    - one element-wise product, ONE forward STFT with zero (not reflect) edge padding and the synthesis window,
    one fused conj-multiply + transpose.  ``spec`` gets no gradient (it is data on this path)."""

    @staticmethod
    def forward(ctx, mask, spec, n_fft, hop, win_length, length, window, outside):
        spec_d, sb, st, sf = _spec_strides(spec.detach())
        rel, irr = explain_spec(spec_d, mask.detach(), n_fft, hop, win_length, length=length, mode="linear", window=window,
                                outside=outside)
        ctx.save_for_backward(spec_d)
        ctx.geom = (n_fft, hop, win_length, length, window, tuple(mask.shape))
        return rel, irr

    @staticmethod
    def backward(ctx, g_rel, g_irr):
        (spec,) = ctx.saved_tensors
        n_fft, hop, win_length, length, window, mshape = ctx.geom
        B, Fb, T = spec.shape
        n_out = int(length) if length is not None else hop * (T - 1)
        if 1 + n_out // hop != T:
            raise NotImplementedError("explain backward needs length // hop + 1 == frames (the reference's geometry)")
        dev = spec.device
        plan_i = get_plan(n_fft, hop, win_length, window, T, 0, n_out)
        inv_env = torch.empty(n_out, dtype=torch.float32, device=dev)
        check(lib().adv_plan_inv_env(plan_i.handle, ptr(inv_env), stream_ptr()), "adv_plan_inv_env")
        zero = lambda g: torch.zeros((B, n_out), dtype=torch.float32, device=dev) if g is None else g.to(dev, torch.float32)
        d = ((zero(g_rel) - zero(g_irr)) * inv_env).contiguous()
        plan_f = get_plan(n_fft, hop, win_length, window, T, n_out, 0)
        A = torch.empty((B, T, Fb), dtype=torch.complex64, device=dev)
        check(lib().adv_stft_ex(plan_f.handle, ptr(d), d.stride(0), B, ptr(A), None, None, _lib.STFT_ZERO_PAD, stream_ptr()),
              "adv_stft_ex")
        m3 = mshape if len(mshape) == 3 else (mshape[0], mshape[2], mshape[3])
        gm = torch.empty(m3, dtype=torch.float32, device=dev)
        check(lib().adv_mask_grad_linear(ptr(spec), spec.stride(0), spec.stride(2), spec.stride(1), ptr(A), B, Fb, T,
                                         m3[1], m3[2], ptr(gm), stream_ptr()), "adv_mask_grad_linear")
        return gm.reshape(mshape), None, None, None, None, None, None, None


@_on_input_device
def explain_linear(spec, mask, n_fft, hop, win_length, length=None, window=None, outside="drop"):
    """Linear-mode explain from an STFT, differentiable w.r.t. ``mask`` (training callers: loss_function.py).
    Out-of-mask bins (``outside``, see :func:`explain`) are constants of the mask either way, so the gradient formula
    does not depend on the choice."""
    if torch.is_grad_enabled() and mask.requires_grad:
        return _ExplainLinearFn.apply(mask, spec, n_fft, hop, win_length, length, window, outside)
    return explain_spec(spec, mask, n_fft, hop, win_length, length=length, mode="linear", window=window, outside=outside)


@_on_input_device
def mask_apply(mag, phase, mask, mode="log1p"):
    """(|X|, angle X, mask) [B,F,T] -> (rel, irr) complex [B,F,T] exactly as the reference spells it
    (expm1(m*log1p(mag)) * exp(1j*phase)); outputs are frame-major like torch.stft's."""
    dev = _dev()
    B, Fb, T = mag.shape
    fm = lambda t: t.to(dev, torch.float32).transpose(1, 2).contiguous()  # no-op copy for our own outputs
    mag_fm, ph_fm = fm(mag), fm(phase)
    mask = _mask3(mask)
    rel = torch.empty((B, T, Fb), dtype=torch.complex64, device=dev)
    irr = torch.empty_like(rel)
    check(lib().adv_mask_apply(ptr(mag_fm), ptr(ph_fm), ptr(mask), B, T, Fb, mask.shape[1], mask.shape[2],
                               MODES[mode], ptr(rel), ptr(irr), stream_ptr()), "adv_mask_apply")
    return rel.transpose(1, 2), irr.transpose(1, 2)


class LmacWorkspace:
    """Device scratch for the metric reduction (block partials, ticket counter, result)."""

    def __init__(self, n, device):
        blocks = lib().adv_lmac_blocks(n)
        self.n = n
        self.partials = torch.empty((blocks, 5), dtype=torch.float64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.sums = torch.zeros(6, dtype=torch.float64, device=device)


@_on_input_device
def lmac(p, theta, q, is_logit=False, want_scores=True, workspace=None, accumulate=False):
    """Three [N] or [N,1] tensors -> (scores [N,7] or None, sums float64 [6]) on the device.
    Score columns: faithfulness, fidelity, AD, AI, AG, pc, oc; sums = the first five summed, then N."""
    dev = _dev()
    flat = lambda t: t.to(dev, torch.float32).reshape(-1).contiguous()
    p, theta, q = flat(p), flat(theta), flat(q)
    n = p.numel()
    if theta.numel() != n or q.numel() != n:
        raise ValueError("prediction tensors differ in length")
    ws = workspace if workspace is not None and workspace.n == n else LmacWorkspace(n, dev)
    scores = torch.empty((n, 7), dtype=torch.float32, device=dev) if want_scores else None
    flags = (1 if is_logit else 0) | (2 if accumulate else 0)
    check(lib().adv_lmac_reduce(ptr(p), ptr(theta), ptr(q), n, flags, ptr(scores), ptr(ws.sums),
                                ptr(ws.partials), ptr(ws.counter), stream_ptr()), "adv_lmac_reduce")
    return scores, ws.sums


@_on_input_device
def td_mask(wave, attribution, want_mask=True):
    """captum_saliency.py:136-143 per clip: returns (mask, wave*mask, wave*(1-mask))."""
    wave = _f32_rows(wave if wave.dim() == 2 else wave.unsqueeze(0), "wave")
    attr = _f32_rows(attribution if attribution.dim() == 2 else attribution.unsqueeze(0), "attribution")
    wave, attr = wave.contiguous(), attr.contiguous()
    B, n = wave.shape
    m = torch.empty_like(wave) if want_mask else None
    rel, irr = torch.empty_like(wave), torch.empty_like(wave)
    rowmax = torch.empty(B, dtype=torch.float32, device=wave.device)
    check(lib().adv_td_mask(ptr(wave), ptr(attr), B, n, ptr(m), ptr(rel), ptr(irr), ptr(rowmax), stream_ptr()),
          "adv_td_mask")
    return m, rel, irr


@_on_input_device
def mask_head(y1, weight, bias):
    """sigmoid(conv1x1(C->1)) on [B,C,H,W] -> [B,1,H,W] (addvisor.py:57-60,82)."""
    dev = _dev()
    y1 = y1.to(dev, torch.float32).contiguous()
    B, Cc, H, W = y1.shape
    w = weight.to(dev, torch.float32).reshape(-1).contiguous()
    b = bias.to(dev, torch.float32).reshape(-1).contiguous()
    if w.numel() != Cc:
        raise ValueError("mask head weight does not match the channel count")
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    check(lib().adv_mask_head(ptr(y1), ptr(w), ptr(b), B, Cc, H * W, ptr(out), stream_ptr()), "adv_mask_head")
    return out


@_on_input_device
def band_swap(spec_real, spec_voc, f_lo, f_hi):
    """Rows [f_lo, f_hi) of ``spec_voc`` replace those of ``spec_real`` (complex [B,F,T])."""
    dev = _dev()
    fm = lambda t: t.to(dev, torch.complex64).transpose(1, 2).contiguous()
    a, b = fm(spec_real), fm(spec_voc)
    B, T, Fb = a.shape
    out = torch.empty_like(a)
    check(lib().adv_band_swap(ptr(a), ptr(b), B, T, Fb, int(f_lo), int(f_hi), ptr(out), stream_ptr()), "adv_band_swap")
    return out.transpose(1, 2)


def band_rows(n_bins, start_hz, end_hz, f_top=8000.0):
    """Row range [lo, hi) selected by the reference's ``(freqs >= start) & (freqs < end)`` with
    ``freqs = linspace(0, f_top, n_bins)`` (hifigan.py:206-210, train_logReg_swapping.py:66-72)."""
    freqs = torch.linspace(0, f_top, n_bins)
    idx = torch.nonzero((freqs >= start_hz) & (freqs < end_hz)).flatten()
    return (int(idx[0]), int(idx[-1]) + 1) if idx.numel() else (0, 0)


@_on_input_device
def band_swap_all(spec_real, spec_voc, band_width=1000, f_max=8000, f_top=8000.0):
    """The whole fabrication loop of hifigan.py:206-214 / train_logReg_swapping.py:64-75 in one launch: for every
    ``band_width`` band below ``f_max`` the rows of ``spec_voc`` replace those of ``spec_real``.
    spec_* complex [B,F,T] -> complex [n_bands*B, F, T] (band-major), ready for one batched istft."""
    dev = _dev()
    fm = lambda t: t.to(dev, torch.complex64).transpose(1, 2).contiguous()
    a, b = fm(spec_real), fm(spec_voc)
    B, T, Fb = a.shape
    starts = list(range(0, f_max, band_width))
    edges = []
    for s0 in starts:
        lo, hi = band_rows(Fb, s0, s0 + band_width, f_top)
        if edges and edges[-1] != lo:
            raise NotImplementedError("bands must tile the frequency axis contiguously")
        if not edges:
            edges.append(lo)
        edges.append(hi)
    e = torch.tensor(edges, dtype=torch.int32, device=dev)
    out = torch.empty((len(starts) * B, T, Fb), dtype=torch.complex64, device=dev)
    check(lib().adv_band_swap_multi(ptr(a), ptr(b), B, T, Fb, ptr(e), len(starts), ptr(out), stream_ptr()),
          "adv_band_swap_multi")
    return out.transpose(1, 2)
