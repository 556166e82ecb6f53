"""GPU parity of the n_fft 1024 streaming kernels (csrc/transform5_kernels.cu) against torch / the CPU oracle: the
reference's default geometry (1024 / hop 322 / 644-tap rectangular window, audioprocessor.py:23-31), hifigan.py's hann
1024 / hop 256 (188-225), other even hops and window lengths (1 - 3 predecessor strips), runs that begin and end inside
clips, clips shorter than one pass, strided spectrum rows, ``length`` shorter / longer / None, the per-clip statistics
and run-to-run bit-reproducibility.  Tolerance: max|y - ref| / max|ref| <= 1e-4 per clip (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
TOL = 1e-4
GEOMS = [(322, 644, "rect"),      # AudioProcessor() defaults: support = 2 hops
         (256, 1024, "hann"),     # hifigan.py:188-225: 4 hops
         (256, 1024, "rect"),
         (512, 1024, "hann"),     # 2 hops, 8 head rows
         (300, 900, "hann"),      # 3 hops, hop not a multiple of 4
         (128, 500, "rect"),      # support 500 -> 4 hops, wlo = 262
         (64, 256, "hann")]       # smallest hop


@pytest.fixture(scope="module")
def ops(pkg, built_lib):
    assert torch.cuda.is_available()
    return pkg.ops


def relerr(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def window(win, kind):
    return torch.ones(win) if kind == "rect" else torch.hann_window(win)


@pytest.mark.parametrize("hop,win,kind", GEOMS)
@pytest.mark.parametrize("B,n", [(1, 16000), (3, 6440), (37, 3220), (64, 80000), (150, 8000)])
def test_streaming_istft1024_matches_torch(ops, hop, win, kind, B, n):
    if B == 64 and (hop, win) not in ((322, 644), (256, 1024)):
        pytest.skip("full size on the two reference geometries")
    g = torch.Generator().manual_seed(B + n + hop)
    T = 1 + n // hop
    spec = torch.complex(torch.randn(B, 513, T, generator=g), torch.randn(B, 513, T, generator=g))
    spec[-1] *= 1e-3
    w = window(win, kind)
    wk = None if kind == "rect" else w
    lengths = [n - (n % 2), None, n - 38] if hop * (T - 1) % 2 == 0 else [n - (n % 2), n - 38]
    for length in lengths:
        want = torch.istft(spec, 1024, hop_length=hop, win_length=win, window=w, length=length)
        got = ops.istft(spec.cuda(), 1024, hop, win, length=length, window=wk)          # [B,F,T] contiguous: strided rows
        worst = max(relerr(got[b], want[b]) for b in range(B))
        assert got.shape == want.shape and worst < TOL, (length, worst)
    fm = spec.transpose(1, 2).contiguous().transpose(1, 2).cuda()                        # torch.stft's frame-major layout
    n2 = n - (n % 2)
    got2, stats = ops.istft(fm, 1024, hop, win, length=n2, window=wk, return_stats=True)
    want = torch.istft(spec, 1024, hop_length=hop, win_length=win, window=w, length=n2)
    assert relerr(got2, want) < TOL
    s = stats.sum(dim=1).cpu()
    np.testing.assert_allclose(s[:, 0], want.double().sum(dim=1), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(s[:, 1], (want.double() ** 2).sum(dim=1), rtol=1e-4)


def test_streaming_istft1024_takes_the_call(pkg, ops):
    """the plan of the reference-default geometry reports the streaming kernel's statistics layout (slots, not tiles)"""
    plan = ops.get_plan(1024, 322, 644, None, 249, 0, 80000)
    assert plan.tiles_istft(64) % 8 == 0 and plan.tiles_istft(64) >= 16


def test_streaming_istft1024_round_trip_and_determinism(ops):
    B, n, hop, win = 16, 80000, 322, 644
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(B, n, generator=g).cuda()
    X, _, _ = ops.stft(wav, 1024, hop, win, want_mag=False, want_phase=False)
    first = None
    for _ in range(20):
        y, st = ops.istft(X, 1024, hop, win, length=n, return_stats=True)
        if first is None:
            first = (y.clone(), st.clone())
            assert relerr(y, wav) < 1e-5
        else:
            assert torch.equal(y, first[0]) and torch.equal(st, first[1])


# ---------------------------------------------------------------------------------------------------------------------
# fused explain (explain5_kernel)
# ---------------------------------------------------------------------------------------------------------------------
def make(B, n, hop, seed, Fm=None, Tm=None):
    g = torch.Generator().manual_seed(seed)
    T = 1 + n // hop
    wav = 0.1 * torch.randn(B, n, generator=g)
    wav[-1] *= 1e-3                                 # a quiet clip: every bin takes the small-magnitude series
    if B > 2:
        wav[1, n // 3:] = 0.0                       # digital silence inside a clip
    mask = torch.rand(B, Fm or 513, Tm or T, generator=g)
    mask[0, :, : (Tm or T) // 4] = 0.0
    mask[0, :, (Tm or T) // 4: (Tm or T) // 2] = 1.0
    return wav, mask


EGEOMS = [(322, 644, "rect"), (512, 1024, "hann"), (300, 600, "rect"), (256, 700, "hann"), (128, 500, "rect")]


@pytest.mark.parametrize("hop,win,kind", EGEOMS)
@pytest.mark.parametrize("B,n", [(1, 16000), (3, 6440), (37, 3220), (64, 80000), (150, 8000)])
@pytest.mark.parametrize("mode", ["log1p", "linear"])
def test_streaming_explain1024_matches_torch(ops, hop, win, kind, B, n, mode):
    if B == 64 and ((hop, win) != (322, 644) or mode != "log1p"):
        pytest.skip("full size once, on the reference-default geometry")
    wav, mask = make(B, n, hop, B * n + hop)
    w = window(win, kind)
    wk = None if kind == "rect" else w
    X = torch.stft(wav, 1024, hop_length=hop, win_length=win, window=w, return_complex=True)
    mag, ph = X.abs(), X.angle()
    if mode == "log1p":
        lm = torch.log1p(mag)
        specs = [torch.polar(torch.expm1(m * lm), ph) for m in (mask, 1 - mask)]
    else:
        specs = [torch.polar(m * mag, ph) for m in (mask, 1 - mask)]
    want = [torch.istft(sp, 1024, hop_length=hop, win_length=win, window=w, length=n) for sp in specs]
    rel, irr = ops.explain(wav, mask, 1024, hop, win, length=n, mode=mode, window=wk)
    worst = max(max(relerr(rel[b], want[0][b]), relerr(irr[b], want[1][b])) for b in range(B))
    assert worst < TOL, worst
    reln, irrn = ops.explain(wav, mask, 1024, hop, win, length=n, mode=mode, window=wk, normalize=True)
    for got, ref in ((reln, want[0]), (irrn, want[1])):
        refn = (ref - ref.mean(-1, keepdim=True)) / (ref.std(-1, keepdim=True) + 1e-7)
        assert relerr(got, refn) < TOL


@pytest.mark.parametrize("outside", ["drop", "keep_irr"])
@pytest.mark.parametrize("Fm,Tm", [(512, 248), (513, 200), (400, 249), (512, 249)])
def test_streaming_explain1024_sub_size_masks(ops, outside, Fm, Tm):
    """the production case: a 512 x 248 U-Net mask against the 513 x 249 spectrum (LMAC_metrics.py:136-139)"""
    B, n, hop = 5, 80000, 322      # T = 249
    wav, mask = make(B, n, hop, 7 * Fm + Tm, Fm, Tm)
    cfg = dict(sampling_rate=16000, n_fft=1024, hop_length=hop, win_length=644, audio_length=5)
    rel_r, irr_r = R.explain(wav, mask, outside=outside, **cfg)
    rel, irr = ops.explain(wav, mask, 1024, hop, 644, length=n, outside=outside)
    assert relerr(rel, rel_r) < TOL and relerr(irr, irr_r) < TOL


def test_streaming_explain1024_statistics_and_determinism(ops):
    B, n, hop = 9, 40000, 322
    wav, mask = make(B, n, hop, 5)
    tiles = ops.explain_tiles(1024, hop, 644, n, B, length=n)
    outs = []
    for _ in range(3):
        rel = torch.empty(B, n, device="cuda")
        irr = torch.empty(B, n, device="cuda")
        stats = torch.full((B, tiles, 4), float("nan"), dtype=torch.float64, device="cuda")   # every slot must be written
        ops.explain(wav, mask, 1024, hop, 644, length=n, out=(rel, irr, stats))
        outs.append((rel.clone(), irr.clone(), stats.clone()))
    rel, irr, stats = outs[0]
    assert torch.isfinite(stats).all()
    s = stats.sum(dim=1).cpu()
    for col, x in ((0, rel), (2, irr)):
        xd = x.double().cpu()
        np.testing.assert_allclose(s[:, col], xd.sum(dim=1), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(s[:, col + 1], (xd ** 2).sum(dim=1), rtol=1e-5)
    for r2, i2, s2 in outs[1:]:
        assert torch.equal(rel, r2) and torch.equal(irr, i2) and torch.equal(stats, s2)


def test_streaming_explain1024_strided_rows(ops):
    B, n, hop = 3, 16000, 322
    wav, mask = make(B, n, hop, 11)
    cfg = dict(sampling_rate=n, n_fft=1024, hop_length=hop, win_length=644, audio_length=1)
    rel_r, irr_r = R.explain(wav, mask, **cfg)
    wide = torch.zeros(B, n + 13, device="cuda")
    wide[:, 3:3 + n] = wav.cuda()
    rel, irr = ops.explain(wide[:, 3:3 + n], mask, 1024, hop, 644, length=n)
    assert relerr(rel, rel_r) < TOL and relerr(irr, irr_r) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# one-frame-per-warp STFT (stft5_kernel): spectrum only on any window, magnitude / phase on non-rectangular windows
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hop,win,kind", GEOMS + [(322, 1024, "rect"), (250, 644, "hann")])
@pytest.mark.parametrize("B,n", [(1, 600), (2, 16000), (9, 8050), (64, 80000)])
def test_stft1024_matches_torch(ops, hop, win, kind, B, n):
    """clips barely longer than the reflect padding, frames that reach past both clip edges, ragged batches; the full
    compute_stft return (X, |X|, angle) as well as the spectrum alone"""
    if B == 64 and (hop, win) not in ((322, 644), (256, 1024)):
        pytest.skip("full size on the two reference geometries")
    g = torch.Generator().manual_seed(B * n + hop + win)
    wav = 0.1 * torch.randn(B, n, generator=g)
    w = window(win, kind)
    wk = None if kind == "rect" else w
    ref = torch.stft(wav, 1024, hop_length=hop, win_length=win, window=w, return_complex=True)
    X, _, _ = ops.stft(wav, 1024, hop, win, window=wk, want_mag=False, want_phase=False)
    assert X.shape == ref.shape
    scale = float(ref.abs().max())
    assert float((X.cpu() - ref).abs().max()) / scale < TOL
    X2, mag, ph = ops.stft(wav, 1024, hop, win, window=wk)
    assert float((X2.cpu() - ref).abs().max()) / scale < TOL
    assert float((mag.cpu() - ref.abs()).abs().max()) / scale < TOL
    d = (ph.cpu() - ref.angle()).abs()
    d = torch.minimum(d, 2 * torch.pi - d)
    strong = ref.abs() > 1e-3 * scale
    bound = torch.clamp(3e-7 * scale / ref.abs()[strong], min=1e-5)
    assert bool((d[strong] <= bound).all())
