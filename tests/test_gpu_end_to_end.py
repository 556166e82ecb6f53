"""GPU end-to-end: run_addvisor_metrics (LMAC_metrics.py:117-172) with a seeded random-init SSL classifier and a
small mask network, against the same loop written with the oracle on CPU.  Checks per-clip probabilities
(abs <= 1e-3) and, when no threshold flips are possible, the five printed means (abs <= 1e-3 of their scale)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ref_path as R

pytestmark = pytest.mark.gpu


def tiny_wav2vec2(seed=0):
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    cfg = Wav2Vec2Config(hidden_size=192, num_hidden_layers=10, num_attention_heads=4, intermediate_size=384,
                         conv_dim=(64, 64, 64, 64, 64, 64, 64), feat_extract_norm="layer", do_stable_layer_norm=True,
                         num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4, mask_time_prob=0.0,
                         mask_feature_prob=0.0)
    torch.manual_seed(seed)
    return Wav2Vec2Model(cfg).eval()


class MagMask(nn.Module):
    """stand-in mask network on the magnitude spectrogram (train_addvisor.py:363 call shape)"""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.net = nn.Sequential(nn.Conv2d(1, 4, 3, padding=1), nn.Tanh(), nn.Conv2d(4, 1, 3, padding=1), nn.Sigmoid())

    def forward(self, x):
        return self.net(torch.log1p(x))


def test_run_addvisor_metrics_matches_oracle_loop(pkg, built_lib):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sr, n_fft, hop, win = 16000, 512, 160, 512
    cfg = dict(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win, audio_length=1)
    g = torch.Generator().manual_seed(0)
    N = 6
    waves = 0.1 * torch.randn(N, sr, generator=g) * torch.linspace(0.2, 1.5, N).unsqueeze(1)

    ssl_cpu = tiny_wav2vec2()
    clf = pkg.classifier_embedder.SimpleLogReg(0.8 * torch.randn(1, 192, generator=g).numpy(), [0.05])
    mask_cpu = MagMask().eval()

    # ---------------- oracle loop on CPU (LMAC_metrics.py:125-158 with the restated arithmetic) ----------------
    coef = torch.tensor(clf.coef_, dtype=torch.float32)
    icpt = torch.tensor(clf.intercept_, dtype=torch.float32)

    def feats_cpu(w):
        return ssl_cpu(R.zero_mean_unit_var_norm(w), output_hidden_states=True).hidden_states[9]

    p_ref, th_ref, q_ref = [], [], []
    with torch.no_grad():
        for i in range(0, N, 4):
            w = waves[i:i + 4]
            _, p = R.logreg(feats_cpu(w).mean(dim=1), coef, icpt)
            _, mag, _ = R.compute_stft(w, **cfg)
            mask = mask_cpu(mag.unsqueeze(1))[:, 0]
            rel, irr = R.explain(w, mask, **cfg)
            _, th = R.logreg(feats_cpu(rel).mean(dim=1), coef, icpt)
            _, q = R.logreg(feats_cpu(irr).mean(dim=1), coef, icpt)
            p_ref.append(p), th_ref.append(th), q_ref.append(q)
    p_ref, th_ref, q_ref = torch.cat(p_ref), torch.cat(th_ref), torch.cat(q_ref)
    want = R.lmac_means(p_ref, th_ref, q_ref).double().numpy()

    # ---------------- product path on the GPU ----------------
    ce = pkg.classifier_embedder
    ce.configure(wav2vec2=copy.deepcopy(ssl_cpu).cuda(), classifier=clf)
    ap = pkg.audioprocessor.AudioProcessor(**cfg)
    head = ce.TorchLogReg().cuda()
    mask_gpu = copy.deepcopy(mask_cpu).cuda()
    res = pkg.LMAC_metrics.run_addvisor_metrics(None, None, batch_size=4, waveforms=waves, model=mask_gpu,
                                                audio_processor=ap, torch_log_reg=head, mask_input="magnitude",
                                                verbose=False)
    assert res["count"] == N

    # per-clip probabilities through the same pieces
    with torch.no_grad():
        f = ap.extract_features(waves)
        _, p = head(f.mean(dim=1))
        _, mag, _ = ap.compute_stft(waves)
        rel, irr = ap.explain(waves, mask_gpu(mag.unsqueeze(1)), normalize=True)
        net = ce.get_wav2vec2()
        _, th = head(net(rel, output_hidden_states=True).hidden_states[9].mean(dim=1))
        _, q = head(net(irr, output_hidden_states=True).hidden_states[9].mean(dim=1))
    for got, ref in ((p, p_ref), (th, th_ref), (q, q_ref)):
        assert float((got.cpu() - ref).abs().max()) < 1e-3

    # metric plumbing, flip-free: the driver's means == the oracle's metric functions on the GPU path's own
    # probabilities (the driver recomputes exactly these, deterministically)
    got = np.array([res[k] for k in pkg.LMAC_metrics.METRIC_NAMES])
    own = R.lmac_means(p.cpu(), th.cpu(), q.cpu()).double().numpy()
    np.testing.assert_allclose(got, own, atol=1e-3)
    # and against the all-CPU loop whenever no decision can flip within the probability tolerance
    margin = min(float((p_ref - 0.5).abs().min()), float((th_ref - 0.5).abs().min()),
                 float((R.score_for_predicted_class(th_ref) - R.score_for_predicted_class(p_ref)).abs().min()))
    if margin > 2e-3:
        np.testing.assert_allclose(got, want, atol=0.2)          # AD / AG are percentages
        np.testing.assert_allclose(got[:2], want[:2], atol=2e-3)


def test_saliency_path(pkg, built_lib):
    """captum_saliency.py flow: input-x-gradient attribution (torch autograd) -> our time-domain mask kernels ->
    FF / fidelity from our metric kernel."""
    ce = pkg.classifier_embedder
    ssl = tiny_wav2vec2(1).cuda()
    g = torch.Generator().manual_seed(1)
    clf = ce.SimpleLogReg(0.8 * torch.randn(1, 192, generator=g).numpy(), [0.0])
    ce.configure(wav2vec2=ssl, classifier=clf)
    ap = pkg.audioprocessor.AudioProcessor(sampling_rate=16000, n_fft=512, hop_length=160, win_length=512, audio_length=1)
    model = pkg.captum_saliency.Wav2vec2LogReg(ap, ce.TorchLogReg().cuda())
    waves = [0.1 * torch.randn(16000, generator=g) for _ in range(3)]
    out = pkg.captum_saliency.compute_camptum_saliency_metrics(model, waves, verbose=False)
    assert out["count"] == 3 and 0.0 <= out["fidelity"] <= 1.0 and abs(out["faithfulness"]) <= 1.0
    # integrated gradients restatement: completeness (sum of attributions == F(x) - F(0)) on a smooth function.
    # (The reference's own model is scale-invariant - its normaliser divides by the clip's std - so along the
    # straight path from 0 its attributions sum to ~0 by construction; checked as well.)
    w = waves[0].reshape(1, -1).cuda()
    f = lambda x: (x ** 3).sum(dim=1) + (2.0 * x).sum(dim=1)
    ig = pkg.captum_saliency.integrated_gradients(f, w, n_steps=16)
    assert abs(float(ig.sum()) - float(f(w))) < 1e-4 * abs(float(f(w))) + 1e-5
    ig_model = pkg.captum_saliency.integrated_gradients(model, w, n_steps=8)
    assert abs(float(ig_model.sum())) < 1e-2


# ---------------------------------------------------------------------------------------------------
# training loss (loss_function.py:32-77): forward and backward of the linear mask path
# ---------------------------------------------------------------------------------------------------
def _torch_linear_path(mask, spec, n_fft, hop, win, window, n):
    """loss_function.py:36-47 in plain torch on the CPU: magnitude and phase cropped to the mask's extent (:38-45),
    both masked spectra zero-padded back to the full grid so that torch.istft accepts them."""
    B, Fb, T = spec.shape
    pad = (0, T - mask.shape[2], 0, Fb - mask.shape[1])
    m = torch.nn.functional.pad(mask, pad)
    inside = torch.nn.functional.pad(torch.ones_like(mask), pad)
    w = torch.ones(win) if window is None else window
    ist = lambda s: torch.istft(s, n_fft, hop_length=hop, win_length=win, window=w, length=n)
    return ist(m * spec), ist((inside - m) * spec)


@pytest.mark.parametrize("n_fft,hop,win,hann,n,Fm,Tm", [
    (512, 160, 512, False, 4800, 257, 31),     # benchmark geometry (rectangular window), full-size mask
    (512, 128, 400, True, 4096, 250, 30),      # hann window shorter than n_fft, mask smaller than the grid
    (1024, 322, 644, True, 9660, 512, 30),     # reference default geometry, U-Net-sized mask (512 x T-1)
])
def test_explain_linear_backward_matches_torch_autograd(pkg, built_lib, n_fft, hop, win, hann, n, Fm, Tm):
    ops = pkg.ops
    g = torch.Generator().manual_seed(n_fft + hop + Fm)
    B = 3
    wav = 0.1 * torch.randn(B, n, generator=g)
    window = torch.hann_window(win) if hann else None
    spec = torch.stft(wav, n_fft, hop_length=hop, win_length=win, window=torch.ones(win) if window is None else window,
                      return_complex=True)
    T = spec.shape[2]
    assert Tm <= T
    mask0 = torch.rand(B, Fm, Tm, generator=g)
    a, b = torch.randn(B, n, generator=g), torch.randn(B, n, generator=g)

    m_ref = mask0.clone().requires_grad_(True)
    rel_r, irr_r = _torch_linear_path(m_ref, spec, n_fft, hop, win, window, n)
    (rel_r * a).sum().add((irr_r * b).sum()).backward()

    m_gpu = mask0.clone().cuda().requires_grad_(True)
    rel, irr = ops.explain_linear(spec.cuda(), m_gpu, n_fft, hop, win, length=n, window=window)
    assert float((rel.cpu() - rel_r.detach()).abs().max() / rel_r.abs().max()) < 1e-4
    assert float((irr.cpu() - irr_r.detach()).abs().max() / irr_r.abs().max()) < 1e-4
    ((rel * a.cuda()).sum() + (irr * b.cuda()).sum()).backward()
    err = float((m_gpu.grad.cpu() - m_ref.grad).abs().max() / m_ref.grad.abs().max())
    assert err < 1e-4, err


def test_lmac_loss_forward_backward(pkg, built_lib):
    """LMACLoss.loss_function on the GPU path vs the reference's expressions in plain torch on the CPU, same seeded
    SSL model and logistic head: total loss, the three components and the gradient reaching the mask."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import torch.nn.functional as F
    sr, n_fft, hop, win = 16000, 512, 160, 512
    cfg = dict(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win, audio_length=1)
    g = torch.Generator().manual_seed(11)
    B = 2
    wav = 0.1 * torch.randn(B, sr, generator=g)
    ssl_cpu = tiny_wav2vec2(1)
    clf = pkg.classifier_embedder.SimpleLogReg(0.8 * torch.randn(1, 192, generator=g).numpy(), [0.05])
    class_pred = torch.tensor([[1.0], [0.0]])
    _, mag, phase = R.compute_stft(wav, **cfg)
    spec = torch.polar(mag, phase)
    mask0 = torch.rand(B, 1, 256, mag.shape[2] - 1, generator=g)

    # ---- reference expressions on the CPU (loss_function.py:32-62) ----
    coef = torch.tensor(clf.coef_, dtype=torch.float32)
    icpt = torch.tensor(clf.intercept_, dtype=torch.float32)
    m_ref = mask0.clone().requires_grad_(True)
    rel_r, irr_r = _torch_linear_path(m_ref.squeeze(1), spec, n_fft, hop, win, None, sr)

    def logits_cpu(w):
        x = (w - w.mean(dim=-1, keepdim=True)) / (w.std(dim=-1, keepdim=True) + 1e-7)
        feats = ssl_cpu(x, output_hidden_states=True).hidden_states[9].mean(dim=1)
        return feats @ coef.t() + icpt

    l_in = F.binary_cross_entropy_with_logits(logits_cpu(rel_r), class_pred)
    l_out = F.binary_cross_entropy_with_logits(logits_cpu(irr_r), 1 - class_pred)
    l1 = m_ref.abs().mean()
    w_ref = F.softplus(torch.tensor([3.0, 0.5, 3.0]))
    total_ref = (w_ref * torch.stack([l_in, l_out, l1])).sum()
    total_ref.backward()

    # ---- product path ----
    ce = pkg.classifier_embedder
    ce.configure(wav2vec2=copy.deepcopy(ssl_cpu).cuda(), classifier=clf)
    ap = pkg.audioprocessor.AudioProcessor(**cfg)
    loss_mod = importlib_loss(pkg).LMACLoss(audio_processor=ap, torch_logreg=ce.TorchLogReg(clf)).cuda()
    m_gpu = mask0.clone().cuda().requires_grad_(True)
    total, losses, w = loss_mod.loss_function(m_gpu, mag.cuda(), phase.cuda(), class_pred.cuda())
    total.backward()
    assert abs(float(total) - float(total_ref)) < 2e-3 * max(1.0, abs(float(total_ref)))
    assert float((losses.cpu() - torch.stack([l_in, l_out, l1]).detach()).abs().max()) < 2e-3
    gref = m_ref.grad
    err = float((m_gpu.grad.cpu() - gref).abs().max() / gref.abs().max())
    assert err < 5e-3, err


def importlib_loss(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".loss_function")


# ---------------------------------------------------------------------------------------------------
# drop-in import lines and BASELINE configs[0]
# ---------------------------------------------------------------------------------------------------
def test_dropin_reference_import_lines(pkg, built_lib):
    """The reference's own import lines (LMAC_metrics.py:4-6, loss_function.py:11-12, captum_saliency.py:1-2) with
    only ``dropin/`` added to sys.path, then a call through the names they bind."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch\n"
        "from audioprocessor import AudioProcessor\n"
        "from classifier_embedder import TorchLogReg, zero_mean_unit_var_norm\n"
        "from LMAC_metrics import compute_AD, compute_AI, compute_AG, compute_fidelity, compute_faithfulness\n"
        "from loss_function import LMACLoss\n"
        "from addvisor import UNet\n"
        "ap = AudioProcessor()\n"
        "x = 0.1 * torch.randn(2, 80000)\n"
        "X, mag, ph = ap.compute_stft(x)\n"
        "assert X.shape == (2, 513, 249) and X.is_cuda\n"
        "y = ap.compute_invert_stft(X)\n"
        "assert float((y.cpu() - x).abs().max()) < 1e-5\n"
        "p = torch.tensor([[.9],[.2],[.5],[.7]]); th = torch.tensor([[.95],[.4],[.6],[.3]])\n"
        "assert abs(compute_AD(th, p).cpu().tolist()[1] - 25.0) < 1e-4 and compute_fidelity(th, p, torch.Tensor([0.5])).shape == (4, 1)\n"
        "print('dropin ok')\n") % os.path.join(root, "dropin")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stdout + out.stderr


def test_cfg1_bundled_wavs_match_reference(pkg, built_lib):
    """BASELINE configs[0]: the 4 bundled wavs, reference-default geometry, seeded U-Net mask (zero-extended) and a full
    random mask, seeded XLS-R-shaped classifier + logistic head, against what the UNMODIFIED reference computed on the
    CPU (tests/golden/cfg1_wavs.npz, oracle/make_golden.py --only-cfg1): transforms <= 1e-4, per-clip |dp| <= 1e-3,
    the five means <= 1e-3 wherever no threshold can flip inside that tolerance."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import cfg1_wavs
    res = cfg1_wavs.run(pkg, timing_reps=3)
    assert res["stft_relerr"] < 1e-4
    assert res["unet_mask_maxabs_err"] < 1e-3      # cuDNN fp32 convolutions vs the CPU's
    for tag in ("unet", "full"):
        r = res[tag]
        assert r["rel_wave_relerr"] < 1e-4 and r["irr_wave_relerr"] < 1e-4, r
        assert r["max_abs_dp"] < 1e-3, r
        got = np.array([r["means"][k] for k in pkg.LMAC_metrics.METRIC_NAMES])
        want = np.array([r["reference_means"][k] for k in pkg.LMAC_metrics.METRIC_NAMES])
        if r["flip_margin_reference"] > 2 * r["max_abs_dp"]:     # no label / AI decision can differ
            np.testing.assert_allclose(got[:2], want[:2], atol=1e-3)          # FF, fidelity (fractions)
            np.testing.assert_allclose(got[2:], want[2:], atol=0.1 + 1e-3)    # AD / AI / AG are percentages: 1e-3 * 100


def test_fused_normalise_and_metric_launch_equals_the_two_launches(pkg, built_lib):
    """adv_normalize_pair_lmac (the metric reduction as an extra CTA of the normaliser's grid) against adv_normalize_pair +
    adv_lmac_reduce: bit-identical waveforms and sums, with and without accumulation; > 1 024 logits falls back."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(21)
    B, n = 37, 16000
    wav = 0.1 * torch.randn(B, n, generator=g)
    mask = torch.rand(B, 257, 101, generator=g)
    logits = 2 * torch.randn(3, B, generator=g).cuda()
    tiles = ops.explain_tiles(512, 160, 512, n, B, length=n)
    outs = []
    for fused in (False, True):
        rel = torch.empty(B, n, device="cuda")
        irr = torch.empty(B, n, device="cuda")
        stats = torch.empty(B, tiles, 4, dtype=torch.float64, device="cuda")
        ops.explain(wav, mask, 512, 160, 512, length=n, out=(rel, irr, stats))
        ws = ops.LmacWorkspace(B, rel.device)
        for rep in range(2):   # second pass accumulates
            if fused:
                if rep == 1:
                    ops.explain(wav, mask, 512, 160, 512, length=n, out=(rel, irr, stats))
                sums = ops.normalize_pair_lmac_(rel, irr, stats, logits[0], logits[1], logits[2], is_logit=True, workspace=ws,
                                                accumulate=rep == 1)
            else:
                if rep == 1:
                    ops.explain(wav, mask, 512, 160, 512, length=n, out=(rel, irr, stats))
                ops.normalize_pair_(rel, irr, stats)
                _, sums = ops.lmac(logits[0], logits[1], logits[2], is_logit=True, want_scores=False, workspace=ws,
                                   accumulate=rep == 1)
        outs.append((rel.clone(), irr.clone(), sums.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2]) and float(outs[0][2][5]) == 2 * B
    big = 2 * torch.randn(3, 1500, generator=g).cuda()   # more than one metric CTA: two launches behind the same call
    rel = torch.randn(4, 8000, device="cuda")
    irr = torch.randn(4, 8000, device="cuda")
    st = torch.stack([rel.double().sum(1), (rel.double() ** 2).sum(1), irr.double().sum(1), (irr.double() ** 2).sum(1)], 1).reshape(4, 1, 4)
    sums = ops.normalize_pair_lmac_(rel, irr, st.contiguous(), big[0], big[1], big[2], is_logit=True)
    assert float(sums[5]) == 1500
