"""GPU parity of the tcgen05 kernels: conv1d implicit GEMM, transposed conv, mel projection and the
full HiFi-GAN generator, against torch fp32 references / the oracle restatement (oracle/vocoder.py).

Tolerances: mel within 1e-4 (max|y-ref| / max|ref|, fp32 via 3xTF32); bf16 vocoder within 1e-2
relative L2 (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from oracle import ref_path as R
from oracle import vocoder as V

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf16r(t):
    return t.to(torch.bfloat16).float()


@pytest.fixture(scope="module")
def H(pkg, built_lib):
    assert torch.cuda.is_available()
    return pkg.hifigan


def run_conv(pkg, x_cl, w, b, dil, reflect=False, pre_slope=1.0, resid=None, scale=1.0):
    """x_cl [B,L,Cin] bf16 cuda; w [Cout,Cin,k] fp32 -> out [B,L,Cout] bf16 (through the C ABI)."""
    L_ = pkg._lib
    gw, kpad = pkg.hifigan._gemm_weight(w)
    gw, bias = gw.cuda(), b.float().cuda()
    B, L, Cin = x_cl.shape
    out = torch.empty(B, L, w.shape[0], dtype=torch.bfloat16, device="cuda")
    L_.check(L_.lib().adv_conv1d_bf16(L_.ptr(x_cl), L_.ptr(gw), L_.ptr(bias), L_.ptr(resid), L_.ptr(out), None, B, L, Cin,
                                      w.shape[2], dil, w.shape[0], kpad, int(reflect), float(pre_slope), 1.0,
                                      float(scale), L_.stream_ptr()), "adv_conv1d_bf16")
    return out


def run_conv_tma(pkg, x_cl, w, b, dil, resid=None, act_slope=0.1):
    """production TMA pipeline: returns (raw, activated) bf16 [B,L,Cout]"""
    L_ = pkg._lib
    gw = w.permute(0, 2, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous().cuda()
    bias = b.float().cuda()
    B, L, Cin = x_cl.shape
    raw = torch.empty(B, L, w.shape[0], dtype=torch.bfloat16, device="cuda")
    act = torch.empty_like(raw)
    L_.check(L_.lib().adv_conv1d_bf16_tma(L_.ptr(x_cl), L_.ptr(gw), L_.ptr(bias), L_.ptr(resid), L_.ptr(raw), L_.ptr(act),
                                          B, L, Cin, w.shape[2], dil, w.shape[0], float(act_slope), 1.0,
                                          L_.stream_ptr()), "adv_conv1d_bf16_tma")
    return raw, act


TMA_CASES = [
    # B, L, Cin, Cout, k, dil, resid
    (2, 300, 64, 64, 3, 1, False),
    (3, 257, 32, 32, 11, 5, True),        # 64-byte swizzle path (Cin = 32), ragged L
    (2, 1000, 128, 128, 7, 3, True),
    (1, 2088, 256, 256, 11, 5, True),     # MRF0 shape: 8 tiles x 44 K-blocks through a 4-stage ring
    (2, 90, 256, 2048, 3, 1, False),      # phase-stacked transposed conv (8 N tiles share each A tile)
    (5, 128, 64, 32, 3, 1, False),
    (40, 700, 64, 64, 7, 1, True),        # more tiles than CTAs: persistent loop + both TMEM buffers
    (2, 500, 64, 64, 11, 5, True),        # slab kernel: 178-row slab, 11 row-shifted descriptors (128-byte swizzle)
    (2, 500, 32, 32, 7, 3, True),         # slab kernel, 64-byte swizzle, odd row shifts
    (1, 300, 32, 32, 3, 1, False),
    (3, 1000, 64, 64, 3, 3, True),
    (2, 700, 256, 256, 7, 3, True),       # slab2 kernel: tile pairs, two channel groups, streamed weights
    (1, 513, 128, 128, 11, 5, True),      # slab2: 306-row slab in two TMA boxes, ragged last pair
    (3, 300, 512, 256, 3, 1, False),      # slab2: four channel groups, two N tiles
]


@pytest.mark.parametrize("B,L,Cin,Cout,k,dil,use_resid", TMA_CASES)
def test_conv1d_tma_matches_torch(pkg, H, B, L, Cin, Cout, k, dil, use_resid):
    g = torch.Generator().manual_seed(B * L + Cin + k + 7)
    x = bf16r(torch.randn(B, Cin, L, generator=g))
    w = bf16r(0.1 * torch.randn(Cout, Cin, k, generator=g))
    b = 0.1 * torch.randn(Cout, generator=g)
    resid = bf16r(torch.randn(B, Cout, L, generator=g)) if use_resid else None
    ref = V._conv(x, w, b, dil, False)
    if resid is not None:
        ref = ref + resid
    x_cl = x.transpose(1, 2).contiguous().to(torch.bfloat16).cuda()
    r_cl = resid.transpose(1, 2).contiguous().to(torch.bfloat16).cuda() if use_resid else None
    raw, act = run_conv_tma(pkg, x_cl, w, b, dil, r_cl)
    assert rel_l2(raw.float().cpu().transpose(1, 2), ref) < 4e-3
    assert rel_l2(act.float().cpu().transpose(1, 2), F.leaky_relu(ref, 0.1)) < 4e-3


RESUNIT_CASES = [
    # B, L, C, k, dil
    (2, 500, 64, 3, 1),
    (3, 257, 64, 7, 5),      # 158-row slab, ragged L, tiles of 122 outputs
    (2, 1000, 64, 3, 5),
    (2, 300, 32, 11, 5),     # 64-byte swizzle: intermediate written with the 2-bit XOR pattern
    (40, 700, 32, 7, 3),     # more tiles than resident CTAs: barrier phases over many iterations
    (1, 100, 32, 3, 1),      # single partial tile
    (2, 131, 64, 7, 1),
]


@pytest.mark.parametrize("B,L,C,k,dil", RESUNIT_CASES)
def test_fused_resunit_matches_torch(pkg, H, B, L, C, k, dil):
    """conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x in one kernel vs torch fp32 with the bf16 roundings of the
    unfused path modelled (activated input and intermediate are bf16)."""
    from importlib import import_module
    L_ = import_module("xai-audio-deepfakes_b200._lib")
    g = torch.Generator().manual_seed(B * L + C + k + dil)
    x = bf16r(torch.randn(B, C, L, generator=g))
    w1 = bf16r(0.1 * torch.randn(C, C, k, generator=g))
    w2 = bf16r(0.1 * torch.randn(C, C, k, generator=g))
    b1 = 0.1 * torch.randn(C, generator=g)
    b2 = 0.1 * torch.randn(C, generator=g)
    xt = bf16r(F.leaky_relu(V._conv(bf16r(F.leaky_relu(x, 0.1)), w1, b1, dil, False), 0.1))
    ref = V._conv(xt, w2, b2, 1, False) + x
    x_cl = x.transpose(1, 2).contiguous().to(torch.bfloat16).cuda()
    gw = lambda w: w.permute(0, 2, 1).reshape(C, -1).to(torch.bfloat16).contiguous().cuda()
    out = torch.empty_like(x_cl)
    w1g, w2g, b1g, b2g = gw(w1), gw(w2), b1.cuda(), b2.cuda()
    L_.check(L_.lib().adv_resunit_bf16(L_.ptr(x_cl), L_.ptr(w1g), L_.ptr(b1g), L_.ptr(w2g), L_.ptr(b2g), L_.ptr(out),
                                       B, L, C, k, dil, 0.1, L_.stream_ptr()), "adv_resunit_bf16")
    assert rel_l2(out.float().cpu().transpose(1, 2), ref) < 4e-3


def test_fused_resunit_rejects_oversized(pkg):
    from importlib import import_module
    L_ = import_module("xai-audio-deepfakes_b200._lib")
    x = torch.zeros(1, 200, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 11 * 64, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(64, device="cuda")
    rc = L_.lib().adv_resunit_bf16(L_.ptr(x), L_.ptr(w), L_.ptr(b), L_.ptr(w), L_.ptr(b), L_.ptr(x), 1, 200, 64, 11, 1,
                                   0.1, L_.stream_ptr())
    assert rc == L_.ADV_ERR_UNSUPPORTED   # two 11-tap weight sets of 64 channels exceed shared memory


CONV_CASES = [
    # B, L, Cin, Cout, k, dil, reflect, pre_slope, resid
    (2, 300, 64, 64, 3, 1, False, 1.0, False),
    (1, 1000, 80, 512, 7, 1, False, 1.0, False),     # conv_pre shape (Cin = 80: K padded 560 -> 576)
    (2, 257, 32, 32, 11, 5, False, 0.1, True),       # narrow MRF stage: one K-block spans two taps
    (3, 130, 128, 128, 7, 3, True, 0.1, True),       # reflect padding variant
    (1, 64, 512, 256, 3, 1, False, 0.1, False),
    (2, 90, 256, 2048, 3, 1, False, 0.1, False),     # phase-stacked transposed conv shape (N tiles of 256)
    (1, 40, 16, 16, 5, 2, False, 1.0, False),        # smallest N tile
]


@pytest.mark.parametrize("B,L,Cin,Cout,k,dil,reflect,slope,use_resid", CONV_CASES)
def test_conv1d_tc_matches_torch(pkg, H, B, L, Cin, Cout, k, dil, reflect, slope, use_resid):
    g = torch.Generator().manual_seed(B * L + Cin + k)
    x = bf16r(torch.randn(B, Cin, L, generator=g))
    w = bf16r(0.1 * torch.randn(Cout, Cin, k, generator=g))
    b = 0.1 * torch.randn(Cout, generator=g)
    resid = bf16r(torch.randn(B, Cout, L, generator=g)) if use_resid else None
    xin = F.leaky_relu(x, slope) if slope != 1.0 else x
    ref = V._conv(bf16r(xin), w, b, dil, reflect)
    if resid is not None:
        ref = ref + resid
    x_cl = x.transpose(1, 2).contiguous().to(torch.bfloat16).cuda()
    r_cl = resid.transpose(1, 2).contiguous().to(torch.bfloat16).cuda() if use_resid else None
    out = run_conv(pkg, x_cl, w, b, dil, reflect, slope, r_cl)
    got = out.float().cpu().transpose(1, 2)
    assert rel_l2(got, ref) < 4e-3      # bf16 output rounding (2^-9) dominates


def test_transposed_conv_as_conv(pkg, H):
    g = torch.Generator().manual_seed(3)
    for cin, cout, k, s, L in [(64, 32, 16, 8, 50), (32, 16, 4, 2, 333)]:
        x = bf16r(torch.randn(2, cin, L, generator=g))
        w = bf16r(0.1 * torch.randn(cin, cout, k, generator=g))
        b = 0.1 * torch.randn(cout, generator=g)
        ref = F.conv_transpose1d(bf16r(F.leaky_relu(x, 0.1)), w, b, stride=s, padding=(k - s) // 2)
        wc, bc = H._transposed_as_conv(w, b, s)
        out = run_conv(pkg, x.transpose(1, 2).contiguous().to(torch.bfloat16).cuda(), wc, bc, 1, False, 0.1)
        got = out.view(2, L * s, cout).float().cpu().transpose(1, 2)
        assert got.shape == ref.shape and rel_l2(got, ref) < 4e-3


@pytest.mark.parametrize("variant", ["torchaudio_default", "speechbrain_vocoder"])
def test_mel_matches_oracle(pkg, variant):
    g = torch.Generator().manual_seed(5)
    wav = 0.1 * torch.randn(3, 16000, generator=g)
    if variant == "torchaudio_default":     # audioprocessor.py:38-44
        ap = pkg.audioprocessor.AudioProcessor(audio_length=1)
        got = ap.mel_transform(wav)
        ref = R.mel_transform(wav)
        err = float((got.cpu() - ref).abs().max() / ref.abs().max())
        assert got.shape == ref.shape and err < 1e-4, err
    else:                                    # hifigan.py:163-178
        got, _ = pkg.mel.mel_spectogram(audio=wav, sample_rate=16000, hop_length=256, win_length=1024, n_mels=80,
                                        n_fft=1024, f_min=0.0, f_max=8000.0, power=1, normalized=False,
                                        min_max_energy_norm=True, norm="slaney", mel_scale="slaney", compression=True)
        ref = V.mel_spectogram(wav)
        assert got.shape == ref.shape
        # log domain: absolute error of log(x) == relative error of x
        assert float((got.cpu() - ref).abs().max()) < 2e-4


def test_mel_against_reference_golden(pkg):
    g = golden("mel_default_small.npz")
    sr, n_fft, hop, win, n_mels = (int(v) for v in g["params"])
    ap = pkg.audioprocessor.AudioProcessor(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win,
                                           n_mels=n_mels, audio_length=1)
    got = ap.mel_transform(torch.from_numpy(g["wav"])).cpu().numpy()
    assert got.shape == g["mel"].shape
    assert np.abs(got - g["mel"]).max() / np.abs(g["mel"]).max() < 1e-4


@pytest.mark.parametrize("variant", ["torchaudio_default", "speechbrain_vocoder"])
@pytest.mark.parametrize("B,n", [(1, 16000), (3, 64000), (5, 23456), (64, 8000)])
def test_mel_fused_single_launch(pkg, variant, B, n):
    """adv_mel_fused (STFT -> |X|^p -> band-compressed bf16x3 tcgen05 filterbank -> log in one launch) against torch's
    own stft + matmul in float64 and against the two-launch path; ragged tile fills (odd frame counts, clips that end
    inside a 64-slot tile, reflect-padded edge frames)."""
    g = torch.Generator().manual_seed(B * 1000 + n)
    wav = 0.1 * torch.randn(B, n, generator=g)
    if variant == "torchaudio_default":
        mt = pkg.mel.MelSpectrogram(16000, 1024, 322, 644, 80)
    else:
        mt = pkg.mel.MelSpectrogram(16000, 1024, 256, 1024, 80, 0.0, 8000.0, 1.0, "slaney", "slaney", log_compress=True)
    got = mt(wav.cuda())
    assert mt.last_path == "fused"
    mt.fused = False
    two = mt(wav.cuda())
    assert mt.last_path == "two-launch"
    win = torch.hann_window(mt.win_length, dtype=torch.float64)
    X = torch.stft(wav.double(), 1024, hop_length=mt.hop_length, win_length=mt.win_length, window=win, return_complex=True)
    ref = (X.abs() ** mt.power).transpose(1, 2) @ mt.fb.double()
    ref = ref.transpose(1, 2)
    if mt.log_compress:
        ref = torch.log(ref.clamp_min(mt.clip))
        assert float((got.cpu().double() - ref).abs().max()) < 2e-4      # log domain: absolute == relative of x
        assert float((got - two).abs().max()) < 2e-4
    else:
        assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-4
        assert float((got - two).abs().max() / two.abs().max()) < 1e-4
    assert got.shape == ref.shape


def test_mel_fused_refuses_dense_bank(pkg):
    """A dense [513 x 80] bank does not fit next to the 128 KB operand tile: ADV_ERR_UNSUPPORTED, the caller falls back
    to adv_stft + adv_mel_project and the result still matches."""
    g = torch.Generator().manual_seed(11)
    wav = 0.1 * torch.randn(2, 16000, generator=g)
    mt = pkg.mel.MelSpectrogram(16000, 1024, 256, 1024, 80)
    mt.fb = torch.rand(513, 80, generator=g)
    got = mt(wav.cuda())
    assert mt.last_path == "two-launch"
    X = torch.stft(wav.double(), 1024, hop_length=256, window=torch.hann_window(1024, dtype=torch.float64), return_complex=True)
    ref = ((X.abs() ** 2).transpose(1, 2) @ mt.fb.double()).transpose(1, 2)
    assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-4


def test_mel_fused_generic_power_and_partial_bank(pkg):
    """power 1.5 (out-of-line pow), 40 mels up to 4 kHz: the upper K chunks are empty and skipped."""
    g = torch.Generator().manual_seed(12)
    wav = 0.1 * torch.randn(4, 12000, generator=g)
    mt = pkg.mel.MelSpectrogram(16000, 1024, 200, 800, 40, 50.0, 4000.0, 1.5)
    got = mt(wav.cuda())
    assert mt.last_path == "fused"
    X = torch.stft(wav.double(), 1024, hop_length=200, win_length=800, window=torch.hann_window(800, dtype=torch.float64),
                   return_complex=True)
    ref = ((X.abs() ** 1.5).transpose(1, 2) @ mt.fb.double()).transpose(1, 2)
    assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-4


@pytest.mark.parametrize("reflect,pipeline,fuse", [(False, "tma", "always"), (False, "tma", "auto"), (False, "tma", "never"),
                                                   (False, "gather", "never"), (True, "gather", "never"),
                                                   (True, "tma", "auto")])
def test_hifigan_generator_matches_oracle(pkg, H, reflect, pipeline, fuse):
    """Whole generator (61 conv launches), seeded weights.  std 0.03 gives per-layer gains near 1 (the original
    N(0, 0.01^2) init lets biases dominate); larger scales saturate tanh and make the comparison chaotic."""
    cfg = type("Cfg", (H.HifiganConfig,), {"pad_reflect": reflect})
    W = H.init_weights(cfg, seed=1, std=0.03)
    gen = H.HifiganGenerator(W, cfg, pipeline=pipeline, fuse=fuse)
    g = torch.Generator().manual_seed(2)
    mel = -4 + 2 * torch.randn(2, 80, 9, generator=g)
    wav = gen.decode_batch(mel)
    assert wav.shape == (2, 1, (9 + 10) * 256)
    Wq = {k: (bf16r(v) if k.endswith("weight") and not k.startswith("conv_post") else v) for k, v in W.items()}
    ref = V.generator(mel, Wq, reflect=reflect)
    # vs exact fp32 math on the same (bf16-rounded) weights
    assert rel_l2(wav, ref) < 1e-2, rel_l2(wav, ref)
    # vs the oracle with bf16 storage of activations modelled: tighter
    refq = V.generator(mel, Wq, reflect=reflect, quantize=bf16r)
    assert rel_l2(wav, refq) < 1e-2, rel_l2(wav, refq)
    if reflect and pipeline == "tma":   # halo layout: every conv output that feeds a conv gets its halo rewritten
        assert gen.launches == 1 + 1 + 4 * (1 + 1 + 18 + 15 + 1)   # conv_pre, relayout, 4 x (up, fix, 18 convs, 15 fixes, average)
    elif fuse == "always":   # 64- and 32-channel stages: fused residual units (k 11 at 64 channels stays unfused)
        assert gen.launches == 1 + 2 * (1 + 18) + (1 + 3 + 3 + 6) + (1 + 9)
    elif fuse == "auto":   # fused only where four CTAs share an SM: the 3-tap branch of the 32-channel stage
        assert gen.launches == 1 + 3 * (1 + 18) + (1 + 3 + 6 + 6)
    else:
        assert gen.launches == 1 + 4 * (1 + 18)   # conv_pre + 4 x (upsample + 3 resblocks x 3 x 2 convs)



def test_hifigan_generator_full_length_reflect(pkg, H):
    """BASELINE configs[2] clip length (251 mel frames = 4 s) through the default (reflect-padded, SpeechBrain's Conv1d)
    generator on the TMA kernels, 8 clips, against the fp32 oracle and its bf16-storage variant; the gather pipeline
    (reflection on load, no halo rows) must agree with the halo pipeline."""
    cfg = H.HifiganConfig
    assert cfg.pad_reflect
    W = H.init_weights(cfg, seed=3, std=0.03)
    g = torch.Generator().manual_seed(4)
    mel = -4 + 2 * torch.randn(8, 80, 251, generator=g)
    gen = H.HifiganGenerator(W, cfg)
    wav = gen.decode_batch(mel)
    assert wav.shape == (8, 1, (251 + 10) * 256)
    Wq = {k: (bf16r(v) if k.endswith("weight") and not k.startswith("conv_post") else v) for k, v in W.items()}
    ref = V.generator(mel, Wq, reflect=True)
    assert rel_l2(wav, ref) < 1e-2, rel_l2(wav, ref)
    # the edges are where reflect and zero padding differ: hold the first / last 2048 samples to the same bar
    for sl in (slice(0, 2048), slice(-2048, None)):
        assert rel_l2(wav[..., sl], ref[..., sl]) < 1e-2
    zero = V.generator(mel[:1], Wq, reflect=False)
    assert rel_l2(zero[..., :2048], ref[:1, :, :2048]) > 5e-2    # (the test can tell the two paddings apart)
    wav_g = H.HifiganGenerator(W, cfg, pipeline="gather").decode_batch(mel[:2])
    assert rel_l2(wav[:2], wav_g.cpu()) < 1e-2


def test_hifigan_generator_batch_256_spot_check(pkg, H):
    """BASELINE configs[2] batch (256 x 4 s): clips of the full batch equal the same clips decoded in a batch of two
    (tile lists, persistent-grid sizing and halo handling do not depend on the batch), and one of them matches the
    oracle."""
    cfg = H.HifiganConfig
    W = H.init_weights(cfg, seed=5, std=0.03)
    g = torch.Generator().manual_seed(6)
    mel = -4 + 2 * torch.randn(256, 80, 251, generator=g)
    gen = H.HifiganGenerator(W, cfg)
    wav = gen.decode_batch(mel)
    pick = [0, 255]
    small = gen.decode_batch(mel[pick])
    assert torch.equal(wav[pick], small)
    Wq = {k: (bf16r(v) if k.endswith("weight") and not k.startswith("conv_post") else v) for k, v in W.items()}
    ref = V.generator(mel[255:256], Wq, reflect=True)
    assert rel_l2(wav[255:256], ref) < 1e-2


def test_load_speechbrain_state_dict(pkg, H):
    """weight_norm'd, ``.conv.``-wrapped names (both torch weight-norm spellings) fold back to the plain weights"""
    W = H.init_weights(seed=7, std=0.03)
    sd = {}
    for k, v in W.items():
        name, kind = k.rsplit(".", 1)
        wrapped = name + ".conv"
        if kind == "bias":
            sd[wrapped + ".bias"] = v
            continue
        dims = [d for d in range(v.dim()) if d != 0]
        gnorm = torch.linalg.vector_norm(v, dim=dims, keepdim=True)
        if name.startswith("resblocks"):
            sd[wrapped + ".parametrizations.weight.original0"] = gnorm
            sd[wrapped + ".parametrizations.weight.original1"] = 3.0 * v      # any positive rescaling of v folds away
        else:
            sd[wrapped + ".weight_g"] = gnorm
            sd[wrapped + ".weight_v"] = 0.5 * v
    back = H.load_speechbrain_state_dict(sd)
    assert set(back) == set(W)
    for k in W:
        assert torch.allclose(back[k], W[k], rtol=1e-5, atol=1e-7), k
    with pytest.raises(KeyError):
        H.load_speechbrain_state_dict({"foo.weight": torch.zeros(1)})


@pytest.mark.parametrize("epilogue", ["tma", "tma_st", "direct", "auto"])
def test_hifigan_epilogue_variants_agree_and_repeat(pkg, H, epilogue):
    """The slab conv kernels' epilogues (per-thread stores, TMA stores, TMA stores + TMA residual loads, per-layer choice)
    compute the same bf16 tensors - bit for bit - and every variant is reproducible run to run."""
    W = H.init_weights(H.HifiganConfig, seed=3, std=0.03)
    g = torch.Generator().manual_seed(4)
    mel = -4 + 2 * torch.randn(5, 80, 37, generator=g)
    base = H.HifiganGenerator(W, H.HifiganConfig, epilogue="direct").decode_batch(mel)
    gen = H.HifiganGenerator(W, H.HifiganConfig, epilogue=epilogue)
    first = gen.decode_batch(mel)
    assert torch.equal(first, base)
    for _ in range(5):
        assert torch.equal(gen.decode_batch(mel), first)
    pkg._lib.lib().adv_set_conv_epilogue(1)
