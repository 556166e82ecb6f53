#!/usr/bin/env python
"""BASELINE configs[0]: ADDvisor explanation + LMAC metrics on the 4 bundled 16 kHz wavs.

The wavs, and what the UNMODIFIED reference computed from them on the CPU, travel as tests/golden/cfg1_wavs.npz
(oracle/make_golden.py --only-cfg1: reference ``AudioProcessor()`` defaults = 5 s / n_fft 1024 / hop 322 / win 644,
the reference's own loop body LMAC_metrics.py:130-157, lifted ``UNet`` / ``TorchLogReg`` / ``compute_*``).  This script
runs the same evaluation on the B200 path and prints one JSON line with both sets of numbers:

  * classifier: ``random_init_wav2vec2(seed=0)`` (9-layer XLS-R-2B shape; the reference's torch module, by the north
    star) + logistic regression ``coef ~ N(0, 0.05^2)``, ``intercept = 0`` taken from the fixture;
  * mask (a) "unet": ``addvisor.UNet`` with ``seeded_init(net, 0)`` on ``mag[:, :512, :248]``, zero-extended to 513 x 249
    (``outside="keep_irr"``); mask (b) "full": ``sigmoid(1.5 * randn)`` with the fixture's seed on the full grid;
  * FF / Fid / AD / AI / AG from ``LMAC_metrics.compute_*`` (the fused metric kernel).

    python scripts/cfg1_wavs.py            # needs a GPU; reads nothing outside the repo
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
FIXTURE = os.path.join(ROOT, "tests", "golden", "cfg1_wavs.npz")


def full_mask(shape):
    """mask (b): the generator state after the fixture's ``coef`` draw (oracle/make_golden.py:golden_cfg1)."""
    g0 = torch.Generator().manual_seed(0)
    torch.randn(1, 1920, generator=g0)
    return torch.sigmoid(1.5 * torch.randn(shape, generator=g0))


def run(pkg, fixture=FIXTURE, timing_reps=20):
    """Evaluate configs[0] on cuda:current; returns a dict with ours / reference numbers and the deviations."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = np.load(fixture)
    ce, M = pkg.classifier_embedder, pkg.LMAC_metrics
    wav = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0).cuda()                  # load_audio's scaling
    ce.configure(wav2vec2=ce.random_init_wav2vec2(seed=0).cuda(), classifier=ce.SimpleLogReg(g["coef"], g["intercept"]))
    head = ce.TorchLogReg().cuda()
    ap = pkg.audioprocessor.AudioProcessor()                                              # reference defaults
    net = pkg.addvisor.seeded_init(pkg.addvisor.UNet(), 0).eval().cuda()
    out = {"clips": [str(n) for n in g["names"]], "geometry": "5 s @16 kHz, n_fft 1024 / hop 322 / win 644"}
    with torch.no_grad():
        X, magnitude, phase = ap.compute_stft(wav)
        out["stft_relerr"] = float((torch.view_as_real(X[:, ::8, ::8]).cpu() - torch.view_as_real(torch.from_numpy(g["X_s"]))).abs().max()
                                   / torch.from_numpy(g["X_s"]).abs().max())
        feats = ap.extract_features(wav)
        _, p = head(torch.mean(feats, dim=1))
        m_unet = net(magnitude[:, :512, :248].unsqueeze(1))
        out["unet_mask_maxabs_err"] = float((m_unet[:, 0, ::4, ::4].cpu() - torch.from_numpy(g["mask_unet_s"])).abs().max())
        masks = {"unet": (m_unet, "keep_irr"), "full": (full_mask(tuple(magnitude.shape)).cuda(), "keep_irr")}
        ssl = ce.get_wav2vec2()
        for tag, (mask, outside) in masks.items():
            rel, irr = ap.explain(wav, mask, outside=outside)
            reln, irrn = ap.explain(wav, mask, normalize=True, outside=outside)
            _, th = head(torch.mean(ssl(reln, output_hidden_states=True).hidden_states[9], dim=1))
            _, q = head(torch.mean(ssl(irrn, output_hidden_states=True).hidden_states[9], dim=1))
            means = [float(M.compute_faithfulness(p, q).mean()), float(M.compute_fidelity(th, p).float().mean()),
                     float(M.compute_AD(th, p).mean()), float(M.compute_AI(th, p).mean()), float(M.compute_AG(th, p).mean())]
            ref_means = [float(v) for v in g[f"{tag}_means"]]
            dp = max(float((p.cpu() - torch.from_numpy(g[f"{tag}_p"])).abs().max()),
                     float((th.cpu() - torch.from_numpy(g[f"{tag}_theta"])).abs().max()),
                     float((q.cpu() - torch.from_numpy(g[f"{tag}_q"])).abs().max()))
            relerr = lambda a, b: float((a.cpu() - torch.from_numpy(b)).abs().max() / torch.from_numpy(b).abs().max())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(timing_reps):
                ap.explain(wav, mask, outside=outside)
            torch.cuda.synchronize()
            t_gpu = (time.perf_counter() - t0) / timing_reps
            out[tag] = {
                "means": dict(zip(M.METRIC_NAMES, means)), "reference_means": dict(zip(M.METRIC_NAMES, ref_means)),
                "max_abs_dp": dp, "flip_margin_reference": float(g[f"{tag}_margin"]),
                "rel_wave_relerr": relerr(rel[:, ::8], g[f"{tag}_rel_s"]), "irr_wave_relerr": relerr(irr[:, ::8], g[f"{tag}_irr_s"]),
                "wave_sums": [[float(v) for v in rel.double().sum(dim=1)], [float(v) for v in irr.double().sum(dim=1)]],
                "p": [float(v) for v in p.flatten()], "theta": [float(v) for v in th.flatten()], "q": [float(v) for v in q.flatten()],
                "transform_ms_b200_wall": t_gpu * 1e3,
                "reference_cpu_seconds_build_container": {"loop_with_ssl": float(g[f"{tag}_seconds"][0]),
                                                          "transforms_only": float(g[f"{tag}_seconds"][1])},
            }
    return out


def main():
    pkg = importlib.import_module("xai-audio-deepfakes_b200")
    pkg._lib.build()
    torch.cuda.set_device(0)
    print(json.dumps(run(pkg)))


if __name__ == "__main__":
    main()
