"""Gradient-saliency baseline pieces on our path (captum_saliency.py:84-100,112-212).

captum is not installed in this image, so the two attribution methods the reference names are
restated in a few lines of torch autograd (the SSL model and head are torch modules anyway):
``input_x_gradient`` (the active choice, captum_saliency.py:117) and ``integrated_gradients``
(:118, zero baseline, Gauss-Legendre steps).  What follows the attribution - |attr| / max, the two
time-domain products, the STFTs and the FF / fidelity sums - runs on our kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import LMAC_metrics as metrics
from . import ops


class Wav2vec2LogReg(nn.Module):
    """captum_saliency.py:84-100: waveform -> features -> mean over time -> logit."""

    def __init__(self, audioprocessor, logReg):
        super().__init__()
        self.ap = audioprocessor
        self.logReg = logReg

    def forward(self, waveform):
        features = self.ap.extract_features(waveform)
        if features.dim() == 2:
            features = features.unsqueeze(0)
        logits, _ = self.logReg(torch.mean(features, dim=1))
        return logits


def input_x_gradient(forward_fn, wave):
    x = wave.clone().detach().requires_grad_(True)
    out = forward_fn(x).sum()
    (grad,) = torch.autograd.grad(out, x)
    return (grad * x).detach()


def integrated_gradients(forward_fn, wave, n_steps=50):
    """captum's defaults: zero baseline, Gauss-Legendre quadrature on [0,1]."""
    nodes, weights = np.polynomial.legendre.leggauss(n_steps)
    alphas, weights = 0.5 * (nodes + 1.0), 0.5 * weights
    total = torch.zeros_like(wave)
    for a, w in zip(alphas, weights):
        x = (float(a) * wave).detach().requires_grad_(True)
        (grad,) = torch.autograd.grad(forward_fn(x).sum(), x)
        total += float(w) * grad
    return (wave * total).detach()


@torch.no_grad()
def saliency_masks(wave, attribution):
    """captum_saliency.py:136-143 -> (mask, wave*mask, wave*(1-mask)), one fused pass per clip."""
    return ops.td_mask(wave, attribution)


def compute_camptum_saliency_metrics(model, waves, method="input_x_gradient", n_steps=50, verbose=True):
    """captum_saliency.py:112-212 over an iterable of [n] waveforms (the reference walks a metadata
    file): attribution -> time-domain masks -> three classifier passes -> FF / fidelity."""
    p, th, q = [], [], []
    for wave in waves:
        wave = wave.reshape(1, -1).to(ops._dev())
        attr = (input_x_gradient(model, wave) if method == "input_x_gradient"
                else integrated_gradients(model, wave, n_steps))
        _, rel, irr = saliency_masks(wave, attr)
        with torch.no_grad():
            p.append(model(wave).reshape(-1))
            th.append(model(rel).reshape(-1))
            q.append(model(irr).reshape(-1))
    sums = metrics.lmac_sums(torch.cat(p), torch.cat(th), torch.cat(q), is_logit=True)
    out = metrics.finalize(sums)
    if verbose:
        print(f"faithfulness : {out['faithfulness']:.2f}")
        print(f"fidelity: {out['fidelity']:.2f}")
    return out
