"""Generate tests/golden/*.npz by EXECUTING the unmodified reference code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

How the reference is run without its weights/datasets:
  * ``accelerate`` and ``classifier_embedder`` are stubbed in ``sys.modules`` so
    that ``/root/reference/audioprocessor.py`` imports VERBATIM (its own
    ``compute_stft`` / ``compute_invert_stft`` / ``mel_transform`` then run on CPU);
  * functions/classes that live in modules with import-time side effects
    (LMAC_metrics.py, classifier_embedder.py, addvisor.py) are lifted with
    ``ast`` (the FunctionDef / ClassDef nodes are compiled unchanged);
  * the inline mask arithmetic (LMAC_metrics.py:136-144,151-154,
    loss_function.py:36-47, captum_saliency.py:136-143) sits inside functions
    that need a DataLoader / checkpoint; the statements on exactly those lines
    are extracted with ``ast`` and exec'd, unchanged, on prepared locals.

Nothing from the reference is copied into the repo: only the numeric outputs.
"""
from __future__ import annotations

import ast
import os
import sys
import types
import wave as wavmod
import warnings

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

warnings.filterwarnings("ignore")


sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_loader import lift, statements_on_lines  # noqa: E402
from ref_loader import import_reference_audioprocessor as _import_ref  # noqa: E402


def import_reference_audioprocessor():
    return _import_ref(REF)


def read_wav(path):
    with wavmod.open(path, "rb") as f:
        assert f.getsampwidth() == 2 and f.getnchannels() == 1
        raw = f.readframes(f.getnframes())
        return np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0, f.getframerate()


def c2n(t):
    return t.detach().cpu().numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    ap_mod, ce = import_reference_audioprocessor()
    AudioProcessor = ap_mod.AudioProcessor
    g = torch.Generator().manual_seed(1234)

    # ---- 1. STFT / iSTFT / normaliser fixtures, three geometries ------------------
    geoms = {
        # name: (sampling_rate, n_fft, hop, win, audio_length[s], input lengths)
        "cfg2_small": (2400, 512, 160, 512, 1, [2400, 1900, 2900]),     # BASELINE cfg-2 geometry
        "default_small": (4000, 1024, 322, 644, 1, [4000, 3000, 4600]),  # reference defaults
    }
    for name, (sr, n_fft, hop, win, al, lens) in geoms.items():
        ap = AudioProcessor(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win,
                            audio_length=al)
        out = {"params": np.array([sr, n_fft, hop, win, al])}
        for i, L in enumerate(lens):
            wav = 0.1 * torch.randn(2, L, generator=g)
            X, mag, ph = ap.compute_stft(wav)
            y = ap.compute_invert_stft(X)
            out[f"wav{i}"] = c2n(wav)
            if i == 0:  # spectra only for the exact-length case keeps the fixture small
                out["X0"], out["mag0"], out["phase0"] = c2n(X), c2n(mag), c2n(ph)
            out[f"istft{i}"] = c2n(y)
            out[f"norm{i}"] = c2n(ce.zero_mean_unit_var_norm(y))
        # 1-D input path (audioprocessor.py:83-90)
        wav1 = 0.1 * torch.randn(lens[0], generator=g)
        X1, m1, p1 = ap.compute_stft(wav1)
        out["wav1d"], out["X1d"], out["istft1d"] = c2n(wav1), c2n(X1), c2n(ap.compute_invert_stft(X1))
        np.savez_compressed(os.path.join(OUT, f"stft_{name}.npz"), **out)
        print(name, "T,F =", X.shape[-1], X.shape[-2])

    # ---- 2. inline mask arithmetic, executed from the reference's own lines -------
    lm_code, lm_lines = statements_on_lines(os.path.join(REF, "LMAC_metrics.py"),
                                            "run_addvisor_metrics", [(136, 144), (151, 154)])
    lin_code, lin_lines = statements_on_lines(os.path.join(REF, "loss_function.py"),
                                              "loss_function", [(36, 47)])
    print("LMAC_metrics lines", lm_lines, "loss_function lines", lin_lines)
    for name, (sr, n_fft, hop, win, al, lens) in geoms.items():
        ap = AudioProcessor(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win,
                            audio_length=al)
        wav = 0.1 * torch.randn(3, lens[0], generator=g)
        wav[2] *= 1e-3  # a quiet clip: exercises the small-|X| limit of expm1(m*log1p(a))
        _, magnitude, phase = ap.compute_stft(wav)
        mask = torch.rand(magnitude.shape, generator=g)
        mask[0, :, :3] = 0.0
        mask[0, :, 3:6] = 1.0
        env = dict(torch=torch, audio_processor=ap, device=torch.device("cpu"),
                   mask=mask, magnitude=magnitude.clone(), phase=phase.clone())
        exec(lm_code, env)
        lin_env = dict(torch=torch, audio_processor=ap, xhat=mask.unsqueeze(1),
                       X_stft_power=magnitude.clone(), X_stft_phase=phase.clone())
        exec(lin_code, lin_env)
        np.savez_compressed(
            os.path.join(OUT, f"explain_{name}.npz"),
            params=np.array([sr, n_fft, hop, win, al]),
            wav=c2n(wav), mask=c2n(mask),
            rel_spec=c2n(env["relevant_mask"]), irr_spec=c2n(env["irrelevant_mask"]),
            rel_wav=c2n(env["istft_waveforms"]), irr_wav=c2n(env["istft_irr_waveform"]),
            rel_norm=c2n(ce.zero_mean_unit_var_norm(env["istft_waveforms"])),
            irr_norm=c2n(ce.zero_mean_unit_var_norm(env["istft_irr_waveform"])),
            lin_rel_wav=c2n(lin_env["istft_relevant_mask"]),
            lin_irr_wav=c2n(lin_env["istft_irelevant_mask"]),
        )

    # ---- 3. LMAC metric functions (lifted unchanged) ------------------------------
    menv = dict(torch=torch, F=torch.nn.functional, device=torch.device("cpu"), eps=1e-10)
    lift(os.path.join(REF, "LMAC_metrics.py"),
         ["compute_fidelity", "get_score_for_predicted_class", "compute_faithfulness",
          "compute_AD", "compute_AI", "compute_AG"], menv)
    # known-answer vector of SURVEY.md section 8(a10) + seeded logits (incl. exact 0.5 ties)
    ka = [torch.tensor(v).view(-1, 1) for v in
          ([.9, .2, .5, .7], [.95, .4, .6, .3], [.1, .8, .5, .65])]
    logits = 2.0 * torch.randn(3, 257, 1, generator=g)
    logits[:, :4, :] = 0.0
    logits[0, 4, 0], logits[1, 4, 0] = 30.0, -30.0
    logits[0, 5, 0], logits[1, 5, 0] = -90.0, 90.0
    sets = {"ka": ka, "rand": [torch.sigmoid(l) for l in logits]}
    mout = {"logits": c2n(logits)}
    for tag, (p, th, q) in sets.items():
        mout[f"{tag}_p"], mout[f"{tag}_theta"], mout[f"{tag}_q"] = c2n(p), c2n(th), c2n(q)
        mout[f"{tag}_ff"] = c2n(menv["compute_faithfulness"](p, q))
        mout[f"{tag}_fid"] = c2n(menv["compute_fidelity"](th, p))
        mout[f"{tag}_ad"] = c2n(menv["compute_AD"](th, p))
        mout[f"{tag}_ai"] = c2n(menv["compute_AI"](th, p))
        mout[f"{tag}_ag"] = c2n(menv["compute_AG"](th, p))
    np.savez_compressed(os.path.join(OUT, "lmac_metrics.npz"), **mout)

    # ---- 4. TorchLogReg (lifted) ---------------------------------------------------
    coef = 0.05 * torch.randn(1, 1920, generator=g)
    icpt = torch.tensor([0.1])
    lenv = dict(torch=torch, nn=torch.nn,
                classifier=types.SimpleNamespace(coef_=c2n(coef), intercept_=c2n(icpt)))
    lift(os.path.join(REF, "classifier_embedder.py"), ["TorchLogReg"], lenv)
    feats = torch.randn(5, 1920, generator=g)
    lg, pr = lenv["TorchLogReg"]()(feats)
    np.savez_compressed(os.path.join(OUT, "logreg.npz"), coef=c2n(coef), intercept=c2n(icpt),
                        feats=c2n(feats), logits=c2n(lg), probs=c2n(pr))

    # ---- 5. UNet mask head (lifted ConvBlock/UNet; hook the head's input) ----------
    uenv = dict(torch=torch, nn=torch.nn)
    lift(os.path.join(REF, "addvisor.py"), ["ConvBlock", "UNet"], uenv)
    torch.manual_seed(0)
    net = uenv["UNet"]().eval()
    grabbed = {}
    net.mask_head.register_forward_hook(lambda m, i, o: grabbed.update(y1=i[0].detach()))
    with torch.no_grad():
        m = net(torch.rand(2, 1, 32, 12, generator=g))
    np.savez_compressed(os.path.join(OUT, "mask_head.npz"), y1=c2n(grabbed["y1"]), mask=c2n(m),
                        weight=c2n(net.mask_head[0].weight), bias=c2n(net.mask_head[0].bias))

    # ---- 6. time-domain (captum) mask lines ----------------------------------------
    td_code, td_lines = statements_on_lines(os.path.join(REF, "captum_saliency.py"),
                                            "compute_camptum_saliency_metrics", [(136, 143)])
    print("captum lines", td_lines)
    wave = 0.1 * torch.randn(1, 3000, generator=g)
    sal = torch.randn(1, 3000, generator=g) * wave
    tenv = dict(torch=torch, wave=wave, saliency_map=sal)
    exec(td_code, tenv)
    np.savez_compressed(os.path.join(OUT, "td_mask.npz"), wave=c2n(wave), attr=c2n(sal),
                        mask=c2n(tenv["mask"]), rel=c2n(tenv["wave_relevant"]),
                        irr=c2n(tenv["wave_irrelevant"]))

    # ---- 7. bundled wavs (1 s excerpts) through the default-geometry processor -----
    ap = AudioProcessor(sampling_rate=8000, audio_length=1)  # 8000-sample excerpts
    wout = {}
    for nm in ("fake_original", "real_original"):
        x, sr = read_wav(os.path.join(REF, "audio_samples", nm + ".wav"))
        assert sr == 16000
        seg = torch.from_numpy(x[16000:24000].copy())
        X, mag, ph = ap.compute_stft(seg)
        wout[nm + "_wav"] = c2n(seg)
        wout[nm + "_X"] = c2n(X)
        wout[nm + "_istft"] = c2n(ap.compute_invert_stft(X))
    np.savez_compressed(os.path.join(OUT, "wav_excerpts.npz"), **wout)

    # ---- 8. mel_transform (constructed but never called in the reference) ----------
    ap = AudioProcessor(sampling_rate=4000, audio_length=1)
    wav = 0.1 * torch.randn(2, 4000, generator=g)
    np.savez_compressed(os.path.join(OUT, "mel_default_small.npz"), wav=c2n(wav),
                        mel=c2n(ap.mel_transform(wav)), fb=c2n(ap.mel_transform.mel_scale.fb),
                        params=np.array([4000, 1024, 322, 644, 80]))
    golden_align()
    golden_cfg1()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes:", total)


def golden_align():
    """---- 9. align_waveforms (hifigan.py:113-136), the FunctionDef lifted unchanged with ``ast`` (the module body
    cannot be imported: it downloads the vocoder).  Three pairs: delayed / advanced / equal-length noisy copies."""
    import torch.nn.functional as F
    env = {"torch": torch, "F": F}
    lift(os.path.join(REF, "hifigan.py"), ["align_waveforms"], env)
    g = torch.Generator().manual_seed(77)
    out = {}
    for i, (n_ref, n_deg, delay) in enumerate([(1800, 1800, 23), (2000, 1700, -57), (1500, 1900, 0)]):
        base = torch.randn(n_ref + n_deg + 1024, generator=g)
        ref = base[512:512 + n_ref].clone()
        deg = (0.8 * base[512 + delay:512 + delay + n_deg] + 0.05 * torch.randn(n_deg, generator=g)).clone()
        ra, da = env["align_waveforms"](ref, deg)
        out[f"ref{i}"], out[f"deg{i}"] = c2n(ref), c2n(deg)
        out[f"ref_aligned{i}"], out[f"deg_aligned{i}"] = c2n(ra), c2n(da)
        out[f"delay{i}"] = np.array(delay)
    np.savez_compressed(os.path.join(OUT, "align.npz"), **out)


def golden_cfg1():
    """---- 10. BASELINE configs[0] / SURVEY 8(d) cfg-1: the 4 bundled wavs through the reference's own evaluation loop.

    ``AudioProcessor()`` defaults (5 s, n_fft 1024 / hop 322 / win 644); classifier = seeded random-init 9-layer
    XLS-R-2B-shaped ``Wav2Vec2Model`` (seed 0) behind the reference's ``extract_features`` + the lifted ``TorchLogReg``
    with ``coef ~ N(0, 0.05^2)``, ``intercept = 0`` (seed 0); masks: (a) the lifted reference ``UNet`` with seeded weights
    (``addvisor.seeded_init(net, 0)``: per-key generators, so the product's re-declared UNet gets the same weights) on
    ``mag[:, :512, :248]``, zero-extended to 513 x 249, and (b) a full-size ``sigmoid(N(0, 1.5^2))`` mask (seed 0).
    The loop body is the reference's: LMAC_metrics.py:130-131,136-157,160-162 exec'd unchanged; the five means come from
    the lifted ``compute_*``.  Stored: the PCM of the wavs, every 8th sample of the masked waveforms + float64 sums,
    a strided sample of the masks / spectrum, the per-clip probabilities and the five means, and the wall time of the
    reference CPU path in this container."""
    import importlib
    import time
    ap_mod, ce = import_reference_audioprocessor()
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    if root not in sys.path:
        sys.path.insert(0, root)
    pkg = importlib.import_module("xai-audio-deepfakes_b200")
    names = sorted(f[:-4] for f in os.listdir(os.path.join(REF, "audio_samples")) if f.endswith(".wav"))
    pcm = []
    for nm in names:
        with wavmod.open(os.path.join(REF, "audio_samples", nm + ".wav"), "rb") as f:
            assert f.getsampwidth() == 2 and f.getnchannels() == 1 and f.getframerate() == 16000
            pcm.append(np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").copy())
    assert all(len(x) == 80000 for x in pcm)
    pcm = np.stack(pcm)
    wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0)

    torch.set_num_threads(os.cpu_count() or 1)
    ssl = pkg.classifier_embedder.random_init_wav2vec2(seed=0)
    ap_mod.wav2vec2 = ssl                                   # the module global of audioprocessor.py:17-18
    g0 = torch.Generator().manual_seed(0)
    coef = 0.05 * torch.randn(1, 1920, generator=g0)
    icpt = torch.zeros(1)
    lenv = dict(torch=torch, nn=torch.nn, classifier=types.SimpleNamespace(coef_=c2n(coef), intercept_=c2n(icpt)))
    lift(os.path.join(REF, "classifier_embedder.py"), ["TorchLogReg"], lenv)
    torch_log_reg = lenv["TorchLogReg"]()
    uenv = dict(torch=torch, nn=torch.nn)
    lift(os.path.join(REF, "addvisor.py"), ["ConvBlock", "UNet"], uenv)
    net = pkg.addvisor.seeded_init(uenv["UNet"](), 0).eval()
    menv = dict(torch=torch, F=torch.nn.functional, device=torch.device("cpu"), eps=1e-10)
    lift(os.path.join(REF, "LMAC_metrics.py"),
         ["compute_fidelity", "get_score_for_predicted_class", "compute_faithfulness",
          "compute_AD", "compute_AI", "compute_AG"], menv)
    body, lines = statements_on_lines(os.path.join(REF, "LMAC_metrics.py"), "run_addvisor_metrics",
                                      [(130, 131), (136, 157)])
    tail, tl = statements_on_lines(os.path.join(REF, "LMAC_metrics.py"), "run_addvisor_metrics", [(160, 162)])
    print("cfg1: LMAC_metrics lines", lines, tl)

    ap = ap_mod.AudioProcessor()
    out = {"names": np.array(names), "pcm": pcm, "coef": c2n(coef), "intercept": c2n(icpt),
           "params": np.array([16000, 1024, 322, 644, 5])}
    with torch.no_grad():
        X, magnitude, phase = ap.compute_stft(wav)                    # collate_fn, LMAC_metrics.py:112
        t0 = time.perf_counter()
        features = ap.extract_features(wav)                           # collate_fn, LMAC_metrics.py:113
        t_feat = time.perf_counter() - t0
        t0 = time.perf_counter()
        m_unet = net(magnitude[:, :512, :248].unsqueeze(1))[:, 0]     # UNet call shape of addvisor.py:62-84
        t_unet = time.perf_counter() - t0
        masks = {"unet": torch.zeros_like(magnitude), "full": torch.sigmoid(1.5 * torch.randn(magnitude.shape, generator=g0))}
        masks["unet"][:, :512, :248] = m_unet
        out["X_s"] = c2n(X[:, ::8, ::8])
        out["mask_unet_s"] = c2n(m_unet[:, ::4, ::4])
        for tag, mask in masks.items():
            env = dict(menv)
            env.update(audio_processor=ap, torch_log_reg=torch_log_reg, mask=mask, features=features,
                       magnitude=magnitude.clone(), phase=phase.clone(), theta_out=[], predictions=[],
                       masked_predictions=[])
            t0 = time.perf_counter()
            exec(body, env)
            t_loop = time.perf_counter() - t0
            exec(tail, env)
            # transform-only time of the same lines (no SSL forward): what the B200 path replaces
            t0 = time.perf_counter()
            for _ in range(3):
                lm = torch.log1p(magnitude)
                ap.compute_invert_stft(torch.expm1(mask * lm) * torch.exp(1j * phase))
                ap.compute_invert_stft(torch.expm1((1 - mask) * lm) * torch.exp(1j * phase))
            t_tr = (time.perf_counter() - t0) / 3
            p, th, q = env["predictions"], env["theta_out"], env["masked_predictions"]
            out[f"{tag}_p"], out[f"{tag}_theta"], out[f"{tag}_q"] = c2n(p), c2n(th), c2n(q)
            out[f"{tag}_rel_s"] = c2n(env["istft_waveforms"][:, ::8])
            out[f"{tag}_irr_s"] = c2n(env["istft_irr_waveform"][:, ::8])
            out[f"{tag}_sums"] = np.stack([c2n(env[k].double().sum(dim=1)) for k in ("istft_waveforms", "istft_irr_waveform")] +
                                          [c2n((env[k].double() ** 2).sum(dim=1)) for k in ("istft_waveforms", "istft_irr_waveform")])
            means = [menv["compute_faithfulness"](p, q).mean().item(), menv["compute_fidelity"](th, p).float().mean().item(),
                     menv["compute_AD"](th, p).mean().item(), menv["compute_AI"](th, p).mean().item(),
                     menv["compute_AG"](th, p).mean().item()]
            out[f"{tag}_means"] = np.array(means)
            out[f"{tag}_seconds"] = np.array([t_loop, t_tr])
            pc = menv["get_score_for_predicted_class"](p.squeeze(1))
            oc = menv["get_score_for_predicted_class"](th.squeeze(1))
            margin = min(float((p - 0.5).abs().min()), float((th - 0.5).abs().min()), float((oc - pc).abs().min()))
            out[f"{tag}_margin"] = np.array(margin)
            print(f"cfg1[{tag}] p={p.flatten().tolist()} theta={th.flatten().tolist()} q={q.flatten().tolist()}")
            print(f"cfg1[{tag}] FF/Fid/AD/AI/AG = {means}  flip margin {margin:.2e}  loop {t_loop:.1f} s, transforms {t_tr*1e3:.1f} ms")
        out["seconds_features_unet"] = np.array([t_feat, t_unet])
    np.savez_compressed(os.path.join(OUT, "cfg1_wavs.npz"), **out)
    print("cfg1_wavs.npz bytes:", os.path.getsize(os.path.join(OUT, "cfg1_wavs.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--only-align":   # adds one fixture without touching the others
        os.makedirs(OUT, exist_ok=True)
        golden_align()
    elif len(sys.argv) > 1 and sys.argv[1] == "--only-cfg1":
        os.makedirs(OUT, exist_ok=True)
        golden_cfg1()
    else:
        main()
