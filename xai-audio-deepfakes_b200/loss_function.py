"""``LMACLoss`` with the reference's call surface (loss_function.py:19-77), forward AND backward on the B200 path.

    total, losses, w = LMACLoss().loss_function(xhat, X_stft_power, X_stft_phase, class_pred)

What runs where (reference lines in brackets):
  * m * |X| * e^{j phase}, (1 - m) * |X| * e^{j phase} and the two iSTFTs [36-47]: ONE fused explain kernel in linear
    mode; under autograd its backward is one zero-padded STFT + one fused conj-multiply (ops._ExplainLinearFn);
  * zero_mean_unit_var_norm [extract_features, audioprocessor.py:69-77]: our normaliser kernels, differentiable;
  * wav2vec2, mean over time, TorchLogReg [48-53]: the reference's own torch modules (north star);
  * BCE-with-logits on the two logits, L1 of the mask, softplus weights, optional total variation [54-75]: a handful of
    scalars / one reduction over the mask - torch ops, exactly the reference's expressions.
The reference builds ``audio_processor`` / ``torch_logreg`` as module globals (loss_function.py:14-15); here they are
constructor arguments with the same defaults.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .audioprocessor import AudioProcessor
from .classifier_embedder import TorchLogReg


class LMACLoss(nn.Module):
    def __init__(self, reg_w_tv=0.00, audio_processor=None, torch_logreg=None):
        super().__init__()
        self.reg_w_tv = reg_w_tv
        self.w_raw = nn.Parameter(torch.tensor([3.0, 0.5, 3.0], requires_grad=True))   # loss_function.py:24
        self.audio_processor = audio_processor if audio_processor is not None else AudioProcessor()
        self.torch_logreg = torch_logreg if torch_logreg is not None else TorchLogReg()

    @property
    def w(self):
        return F.softplus(self.w_raw)

    def masked_waveforms(self, xhat, X_stft_power, X_stft_phase):
        """loss_function.py:36-47: (relevant, irrelevant) waveforms [B, audio_length * sr], differentiable in xhat."""
        ap = self.audio_processor
        spec = torch.polar(X_stft_power.to(ops._dev(), torch.float32), X_stft_phase.to(ops._dev(), torch.float32))
        n = int(ap.audio_length * ap.sampling_rate)
        return ops.explain_linear(spec, xhat, ap.n_fft, ap.hop_length, ap.win_length, length=n)

    def loss_function(self, xhat, X_stft_power, X_stft_phase, class_pred):
        xhat = xhat.squeeze(1)
        rel_wave, irr_wave = self.masked_waveforms(xhat, X_stft_power, X_stft_phase)
        ap = self.audio_processor
        rel_feats = torch.mean(ap.extract_features(rel_wave).squeeze(0), dim=1)
        irr_feats = torch.mean(ap.extract_features(irr_wave).squeeze(0), dim=1)
        logreg = self.torch_logreg.to(rel_feats.device)
        rel_logits, _ = logreg(rel_feats)
        irr_logits, _ = logreg(irr_feats)
        class_pred = class_pred.to(rel_logits.device)
        l_in = F.binary_cross_entropy_with_logits(rel_logits, class_pred)
        l_out = F.binary_cross_entropy_with_logits(irr_logits, 1 - class_pred)
        reg_l1 = xhat.abs().mean()
        losses = torch.stack([l_in, l_out, reg_l1.to(l_in.device)])
        w = self.w.to(losses.device)
        total_loss = torch.sum(w * losses)
        if self.reg_w_tv > 0:   # computed and, as in the reference, not added to the returned total
            tv_h = torch.sum(torch.abs(xhat[:, :, :-1] - xhat[:, :, 1:]))
            tv_w = torch.sum(torch.abs(xhat[:, :-1, :] - xhat[:, 1:, :]))
            _ = reg_l1 + (tv_h + tv_w) * self.reg_w_tv
        return total_loss, losses, self.w
