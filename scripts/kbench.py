"""Kernel micro-benchmark for the transform kernels on BASELINE cfg-2 shapes (64 x 64000, n_fft 512, hop 160).

Times each kernel from a CUDA graph of POOL launches over rotating buffers (> L2), CUDA events on the launch
stream, and checks the results against torch.stft / torch.istft on the GPU (quick sanity; the parity suite is
tests/).  Usage: python scripts/kbench.py [stft] [stft3] [istft] [explain] [--win hann] [--nfft 1024]
"""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("which", nargs="*", default=["stft", "stft3", "istft", "explain"])
ap.add_argument("--nfft", type=int, default=512)
ap.add_argument("--hop", type=int, default=160)
ap.add_argument("--win", default="rect")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--n", type=int, default=64000)
ap.add_argument("--pool", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--tag", default="")
ap.add_argument("--winlen", type=int, default=0)
ap.add_argument("--pdl", type=int, default=-1, help="0 / 1: programmatic dependent launch off / on (default: the library's)")
args = ap.parse_args()

pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ops = pkg.ops
if args.pdl >= 0:
    pkg._lib.lib().adv_set_pdl(args.pdl)
B, n, n_fft, hop = args.batch, args.n, args.nfft, args.hop
win_len = args.winlen or (n_fft if n_fft == 512 else 644)
window = None if args.win == "rect" else torch.hann_window(win_len)
F, T = n_fft // 2 + 1, 1 + n // hop
POOL = args.pool
g = torch.Generator(device="cuda").manual_seed(0)
wavs = [0.1 * torch.randn(B, n, generator=g, device="cuda") for _ in range(POOL)]
masks = [torch.rand(B, F, T, generator=g, device="cuda") for _ in range(POOL)]
kw = dict(n_fft=n_fft, hop=hop, win_length=win_len, window=window)
peak = 6537.0


def graph_time(fn, reps=args.reps):
    fn(0)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(POOL):
            fn(i)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (reps * POOL)


def ref_window():
    if window is None:
        w = torch.ones(win_len, device="cuda")
    else:
        w = window.cuda()
    return w


out = {"tag": args.tag, "cfg": [B, n, n_fft, hop, args.win]}
Xr = torch.stft(wavs[0], n_fft, hop_length=hop, win_length=win_len, window=ref_window(), return_complex=True)
if "stft" in args.which:
    X, _, _ = ops.stft(wavs[0], want_mag=False, want_phase=False, **kw)
    err = float((torch.view_as_real(X) - torch.view_as_real(Xr)).abs().max() / Xr.abs().max())
    t = graph_time(lambda i: ops.stft(wavs[i], want_mag=False, want_phase=False, **kw))
    by = 4 * n + 8 * F * T
    out["stft"] = {"us": t * 1e6, "GBps": by * B / t / 1e9, "frac": by * B / t / 1e9 / peak, "err": err}
if "stft3" in args.which:
    X, mag, ph = ops.stft(wavs[0], **kw)
    e_mag = float((mag - Xr.abs()).abs().max() / Xr.abs().max())
    d = (ph - Xr.angle()).abs()
    d = torch.minimum(d, 2 * torch.pi - d) * (Xr.abs() / Xr.abs().max())
    t = graph_time(lambda i: ops.stft(wavs[i], **kw))
    by = 4 * n + 16 * F * T
    out["stft3"] = {"us": t * 1e6, "GBps": by * B / t / 1e9, "frac": by * B / t / 1e9 / peak, "err_mag": e_mag,
                    "err_phase_w": float(d.max())}
if "istft" in args.which:
    specs = [ops.stft(w, want_mag=False, want_phase=False, **kw)[0] for w in wavs]
    y = ops.istft(specs[0], length=n, **kw)
    yr = torch.istft(Xr, n_fft, hop_length=hop, win_length=win_len, window=ref_window(), length=n)
    err = float((y - yr).abs().max() / yr.abs().max())
    t = graph_time(lambda i: ops.istft(specs[i], length=n, **kw))
    by = 4 * n + 8 * F * T
    out["istft"] = {"us": t * 1e6, "GBps": by * B / t / 1e9, "frac": by * B / t / 1e9 / peak, "err": err}
    del specs
if "explain" in args.which:
    rel, irr = ops.explain(wavs[0], masks[0], length=n, **kw)
    mag, phs = Xr.abs(), Xr.angle()
    lm = torch.log1p(mag)
    rr = torch.istft(torch.polar(torch.expm1(masks[0] * lm), phs), n_fft, hop_length=hop, win_length=win_len,
                     window=ref_window(), length=n)
    ir = torch.istft(torch.polar(torch.expm1((1 - masks[0]) * lm), phs), n_fft, hop_length=hop, win_length=win_len,
                     window=ref_window(), length=n)
    err = max(float((rel - rr).abs().max() / rr.abs().max()), float((irr - ir).abs().max() / ir.abs().max()))
    outs = [(torch.empty(B, n, device="cuda"), torch.empty(B, n, device="cuda"), None) for _ in range(POOL)]
    stats = [torch.empty((B, ops.get_plan(n_fft, hop, win_len, window, T, n, n).tiles(B), 4), dtype=torch.float64,
                         device="cuda") for _ in range(POOL)]
    t = graph_time(lambda i: ops.explain(wavs[i], masks[i], length=n, out=(outs[i][0], outs[i][1], stats[i]), **kw))
    by = 4 * n + 4 * F * T + 8 * n
    out["explain"] = {"us": t * 1e6, "GBps": by * B / t / 1e9, "frac": by * B / t / 1e9 / peak, "err": err}
print(json.dumps(out))
