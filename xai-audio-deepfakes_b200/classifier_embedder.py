"""Counterpart of the reference's ``classifier_embedder`` module (classifier_embedder.py:12-63).

The reference loads its weights at import time from paths that only exist on the authors'
machines (:12-16).  Here the frozen classifier is *registered* instead: call
``configure(wav2vec2=..., classifier=...)`` once with the reference's own torch ``Wav2Vec2Model``
and sklearn-style logistic regression (anything with ``coef_ [1,1920]`` / ``intercept_ [1]``).
Both stay plain torch modules, as the north star prescribes; only the normaliser in front of the
SSL model and the sigmoid/metric arithmetic behind the head are our CUDA kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops

wav2vec2 = None    # the reference's `wav2vec2` global (classifier_embedder.py:14-18)
classifier = None  # the reference's `classifier` global (classifier_embedder.py:12)
processor = None   # kept for name compatibility; unused downstream in the reference too


def configure(wav2vec2=None, classifier=None):
    """Register the frozen SSL model and the logistic-regression head."""
    g = globals()
    if wav2vec2 is not None:
        for prm in wav2vec2.parameters():  # classifier_embedder.py:17-18
            prm.requires_grad = False
        g["wav2vec2"] = wav2vec2.eval()
    if classifier is not None:
        g["classifier"] = classifier


def get_wav2vec2():
    if wav2vec2 is None:
        raise RuntimeError("no SSL model registered: call classifier_embedder.configure(wav2vec2=...)")
    return wav2vec2


class TorchLogReg(nn.Module):
    """classifier_embedder.py:21-38: ``Linear(1920, 1)`` holding the sklearn coefficients;
    ``forward`` returns ``(logits, sigmoid(logits))``."""

    def __init__(self, clf=None):
        super().__init__()
        clf = clf if clf is not None else classifier
        if clf is None:
            raise RuntimeError("no logistic regression registered: call configure(classifier=...)")
        coef = torch.as_tensor(np.asarray(clf.coef_), dtype=torch.float32)
        icpt = torch.as_tensor(np.asarray(clf.intercept_), dtype=torch.float32)
        self.linear = nn.Linear(coef.shape[1], coef.shape[0])
        self.linear.weight = nn.Parameter(coef, requires_grad=False)
        self.linear.bias = nn.Parameter(icpt, requires_grad=False)

    def forward(self, x):
        logits = self.linear(x)
        return logits, torch.sigmoid(logits)


class _NormalizeFn(torch.autograd.Function):
    """Forward = our row_stats + normalize kernels.  Backward (gradient-attribution / training callers only,
    outside the evaluation hot path) re-derives the gradient with torch ops from the saved input."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.normalize_(x.detach())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            xd = x.detach().to(g.device).requires_grad_(True)
            y = (xd - xd.mean(dim=-1, keepdim=True)) / (xd.std(dim=-1, keepdim=True) + 1e-7)
            (gx,) = torch.autograd.grad(y, xd, g)
        return gx.to(x.device)


def zero_mean_unit_var_norm(input_values):
    """classifier_embedder.py:59-63 over the last dim: (x - mean) / (std_unbiased + 1e-7).
    Runs the row_stats + normalize kernels; returns a CUDA tensor of the input's shape.  Differentiable
    (the reference keeps this op in torch precisely so that gradients reach the waveform)."""
    shape = input_values.shape
    x = input_values.reshape(-1, shape[-1])
    if torch.is_grad_enabled() and x.requires_grad:
        return _NormalizeFn.apply(x).reshape(shape)
    return ops.normalize_(x).reshape(shape)


class SimpleLogReg:
    """Stand-in for the joblib'd sklearn model when only coefficients are known."""

    def __init__(self, coef, intercept):
        self.coef_ = np.asarray(coef, dtype=np.float64).reshape(1, -1)
        self.intercept_ = np.asarray(intercept, dtype=np.float64).reshape(-1)


def random_init_wav2vec2(seed=0, num_layers=9):
    """Seeded random-init model with the truncated XLS-R-2B shape the reference loads
    (hidden 1920, 16 heads, FFN 7680, 7 conv layers of 512 channels, stable layer norm).
    No checkpoint exists offline; used by tests and the optional end-to-end benchmark."""
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    cfg = Wav2Vec2Config(hidden_size=1920, num_hidden_layers=num_layers, num_attention_heads=16,
                         intermediate_size=7680, feat_extract_norm="layer", do_stable_layer_norm=True,
                         conv_bias=True, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
                         mask_time_prob=0.0, mask_feature_prob=0.0)
    torch.manual_seed(seed)
    return Wav2Vec2Model(cfg).eval()
