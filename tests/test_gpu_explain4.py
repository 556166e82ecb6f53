"""GPU parity of the generation-4 streaming explain kernel (csrc/transform4_kernels.cu) against the CPU oracle:
the geometries it takes (n_fft 512, rectangular window, hop 128 / 160 / 256), runs that begin and end inside clips,
clips shorter than one pass, batches from one clip to many more units than CTAs, masks smaller than the spectrum with
both ``outside`` conventions, a ``length`` other than the input's, the per-clip statistics it hands the normaliser, and
run-to-run bit-reproducibility.  Tolerance: max|y - ref| / max|ref| <= 1e-4 per clip (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops(pkg, built_lib):
    assert torch.cuda.is_available()
    return pkg.ops


def relerr(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make(B, n, hop, seed, Fm=None, Tm=None):
    g = torch.Generator().manual_seed(seed)
    T = 1 + n // hop
    wav = 0.1 * torch.randn(B, n, generator=g)
    wav[-1] *= 1e-3                                 # a quiet clip: every bin takes the small-magnitude series
    if B > 2:
        wav[1, n // 3:] = 0.0                       # digital silence inside a clip
    mask = torch.rand(B, Fm or 257, Tm or T, generator=g)
    mask[0, :, : (Tm or T) // 4] = 0.0
    mask[0, :, (Tm or T) // 4: (Tm or T) // 2] = 1.0
    return wav, mask


@pytest.mark.parametrize("hop", [128, 160, 256])
@pytest.mark.parametrize("B,n", [(1, 16000),      # one clip: the grid shrinks to runs of >= 8 units
                                 (3, 6400),       # 21 units per clip at hop 160: runs straddle clip boundaries
                                 (37, 3200),      # clips shorter than one pass of 16 units: several clips per pass
                                 (64, 64000),     # BASELINE configs[1]
                                 (200, 8000)])    # many more units than 148 x 16
@pytest.mark.parametrize("mode", ["log1p", "linear"])
def test_streaming_explain_matches_oracle(ops, hop, B, n, mode):
    if B == 64 and (hop != 160 or mode != "log1p"):
        pytest.skip("full size once")
    wav, mask = make(B, n, hop, B * n + hop)
    cfg = dict(sampling_rate=n, n_fft=512, hop_length=hop, win_length=512, audio_length=1)
    rel_r, irr_r = R.explain(wav, mask, mode=mode, **cfg)
    rel, irr = ops.explain(wav, mask, 512, hop, 512, length=n, mode=mode)
    worst = max(max(relerr(rel[b], rel_r[b]), relerr(irr[b], irr_r[b])) for b in range(B))
    assert worst < TOL, worst
    # normalised outputs: exercises the per-clip statistics slots
    reln_r, irrn_r = R.explain(wav, mask, mode=mode, normalize=True, **cfg)
    reln, irrn = ops.explain(wav, mask, 512, hop, 512, length=n, mode=mode, normalize=True)
    assert relerr(reln, reln_r) < TOL and relerr(irrn, irrn_r) < TOL


@pytest.mark.parametrize("outside", ["drop", "keep_irr"])
@pytest.mark.parametrize("Fm,Tm", [(256, 100), (257, 90), (200, 101), (256, 101)])
def test_streaming_explain_sub_size_masks(ops, outside, Fm, Tm):
    B, n, hop = 5, 16000, 160      # T = 101
    wav, mask = make(B, n, hop, 7 * Fm + Tm, Fm, Tm)
    cfg = dict(sampling_rate=n, n_fft=512, hop_length=hop, win_length=512, audio_length=1)
    rel_r, irr_r = R.explain(wav, mask, outside=outside, **cfg)
    rel, irr = ops.explain(wav, mask, 512, hop, 512, length=n, outside=outside)
    assert relerr(rel, rel_r) < TOL and relerr(irr, irr_r) < TOL


@pytest.mark.parametrize("length", [15000, 16000, 16100, 15999])
def test_streaming_explain_lengths(ops, length):
    """``length`` shorter / longer than the input (torch.istft trims or zero-pads) and not a multiple of 32"""
    B, n, hop = 4, 16000, 160
    wav, mask = make(B, n, hop, length)
    T = 1 + n // hop
    X = torch.stft(wav, 512, hop_length=hop, win_length=512, window=torch.ones(512), return_complex=True)
    lm = torch.log1p(X.abs())
    ph = X.angle()
    want = [torch.istft(torch.polar(torch.expm1(m * lm), ph), 512, hop_length=hop, win_length=512, window=torch.ones(512),
                        length=length) for m in (mask, 1 - mask)]
    rel, irr = ops.explain(wav, mask, 512, hop, 512, length=length)
    assert rel.shape == (B, length)
    assert relerr(rel, want[0]) < TOL and relerr(irr, want[1]) < TOL


def test_streaming_explain_statistics_and_determinism(ops):
    B, n, hop = 9, 32000, 160
    wav, mask = make(B, n, hop, 5)
    tiles = ops.explain_tiles(512, hop, 512, n, B, length=n)
    outs = []
    for _ in range(3):
        rel = torch.empty(B, n, device="cuda")
        irr = torch.empty(B, n, device="cuda")
        stats = torch.full((B, tiles, 4), float("nan"), dtype=torch.float64, device="cuda")   # every slot must be written
        ops.explain(wav, mask, 512, hop, 512, length=n, out=(rel, irr, stats))
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


def test_streaming_explain_strided_rows_and_unaligned_outputs(ops):
    B, n, hop = 3, 16000, 160
    wav, mask = make(B, n, hop, 11)
    cfg = dict(sampling_rate=n, n_fft=512, hop_length=hop, win_length=512, audio_length=1)
    rel_r, irr_r = R.explain(wav, mask, **cfg)
    wide = torch.zeros(B, n + 13, device="cuda")
    wide[:, 3:3 + n] = wav.cuda()
    big = torch.zeros(2 * B * n + 8, device="cuda")
    rel = big[1:1 + B * n].view(B, n)
    irr = big[B * n + 5:B * n + 5 + B * n].view(B, n)
    ops.explain(wide[:, 3:3 + n], mask, 512, hop, 512, length=n, out=(rel, irr, None))
    assert relerr(rel, rel_r) < TOL and relerr(irr, irr_r) < TOL
    assert float(big[0]) == 0.0 and float(big[B * n + 1:B * n + 5].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# streaming iSTFT (istft4_kernel): same geometry domain, spectrum input with any strides
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hop", [128, 160, 256])
@pytest.mark.parametrize("B,n", [(1, 16000), (3, 6400), (37, 3200), (64, 64000), (200, 8000)])
def test_streaming_istft_matches_torch(ops, hop, B, n):
    if B == 64 and hop != 160:
        pytest.skip("full size once")
    g = torch.Generator().manual_seed(B + n + hop)
    T = 1 + n // hop
    spec = torch.complex(torch.randn(B, 257, T, generator=g), torch.randn(B, 257, T, generator=g))
    spec[-1] *= 1e-3
    w = torch.ones(512)
    for length in (n, None, n - 37):
        want = torch.istft(spec, 512, hop_length=hop, win_length=512, window=w, length=length)
        got = ops.istft(spec.cuda(), 512, hop, 512, length=length)                     # [B,F,T] contiguous: strided rows
        worst = max(relerr(got[b], want[b]) for b in range(B))
        assert got.shape == want.shape and worst < TOL, (length, worst)
    fm = spec.transpose(1, 2).contiguous().transpose(1, 2).cuda()                       # torch.stft's frame-major layout
    got2, stats = ops.istft(fm, 512, hop, 512, length=n, return_stats=True)
    want = torch.istft(spec, 512, hop_length=hop, win_length=512, window=w, length=n)
    assert relerr(got2, want) < TOL
    s = stats.sum(dim=1).cpu()
    np.testing.assert_allclose(s[:, 0], want.double().sum(dim=1), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(s[:, 1], (want.double() ** 2).sum(dim=1), rtol=1e-4)


def test_streaming_istft_round_trip_and_determinism(ops):
    B, n, hop = 16, 64000, 160
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(B, n, generator=g).cuda()
    X, _, _ = ops.stft(wav, 512, hop, 512, want_mag=False, want_phase=False)
    first = None
    for _ in range(20):
        y, st = ops.istft(X, 512, hop, 512, length=n, return_stats=True)
        if first is None:
            first = (y.clone(), st.clone())
            assert relerr(y, wav) < 1e-5
        else:
            assert torch.equal(y, first[0]) and torch.equal(st, first[1])

