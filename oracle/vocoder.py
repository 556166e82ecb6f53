"""torch-CPU fp32 restatement of the vocoder side of hifigan.py (TEST INFRASTRUCTURE ONLY).

**Parity unpinned.**  hifigan.py:92-93,106-110,163-180 calls two SpeechBrain objects
(``lobes.models.FastSpeech2.mel_spectogram`` and ``inference.vocoders.HIFIGAN`` from
``speechbrain/tts-hifigan-libritts-16kHz``).  SpeechBrain is an unpinned, un-vendored dependency of the
reference (no requirements file), it is not installed in this image and the checkpoint is unreachable, so
neither its code nor its outputs can be executed here.  What follows restates the published algorithm:

* mel: ``torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, win_length, hop_length, f_min, f_max,
  n_mels, power, normalized, norm, mel_scale)`` followed by ``log(clamp(mel, 1e-5))`` when
  ``compression`` - that is the body of SpeechBrain's ``mel_spectogram`` (torchaudio IS installed, so this
  half is the same library call SpeechBrain makes);
* generator: HiFi-GAN V1 as implemented by SpeechBrain's ``HifiganGenerator`` (hyper-parameters in
  ``HifiganConfig`` of the product module; corroborated by hop 256 = 8*8*2*2 at hifigan.py:166 and the
  "crop 1330 ~ 5*256" remark at hifigan.py:50).  One detail cannot be settled offline: SpeechBrain's Conv1d
  wrapper may apply *reflect* "same" padding where the original HiFi-GAN uses zeros; both are implemented
  (``pad_reflect``) and tested, zero padding is the default.

The anchor for parity is therefore the reference's own call sites plus this restatement with identical
(seeded) weights; the CUDA path must match it to 1e-2 relative L2 (bf16 activations, fp32 accumulation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1


def mel_spectogram(audio, sample_rate=16000, hop_length=256, win_length=1024, n_fft=1024, n_mels=80, f_min=0.0,
                   f_max=8000.0, power=1, normalized=False, norm="slaney", mel_scale="slaney", compression=True):
    """hifigan.py:163-178 argument set; returns the (optionally log-compressed) mel [.., n_mels, T]."""
    import torchaudio.transforms as T
    tr = T.MelSpectrogram(sample_rate=sample_rate, hop_length=hop_length, win_length=win_length, n_fft=n_fft,
                          n_mels=n_mels, f_min=f_min, f_max=f_max, power=power, normalized=normalized, norm=norm,
                          mel_scale=mel_scale)
    mel = tr(audio)
    return torch.log(torch.clamp(mel, min=1e-5)) if compression else mel


def _conv(x, w, b, dil, reflect):
    k = w.shape[-1]
    pad = (k - 1) // 2 * dil
    if reflect:
        return F.conv1d(F.pad(x, (pad, pad), mode="reflect"), w, b, dilation=dil)
    return F.conv1d(x, w, b, dilation=dil, padding=pad)


def generator(mel, W, upsample_factors=(8, 8, 2, 2), n_kernels=3, n_dils=3, dils=(1, 3, 5), inference_padding=5,
              reflect=False, quantize=None):
    """mel [B,80,T] fp32 -> waveform [B,1,(T+2*pad)*256].  ``quantize`` (e.g. bf16 round-trip) is applied to
    every stored activation so the oracle can also model the product's storage precision."""
    q = quantize if quantize is not None else (lambda t: t)
    o = F.pad(mel, (inference_padding, inference_padding), mode="replicate")
    o = q(_conv(q(o), W["conv_pre.weight"], W["conv_pre.bias"], 1, reflect))
    for i, s in enumerate(upsample_factors):
        w = W[f"ups.{i}.weight"]
        k = w.shape[-1]
        o = q(F.conv_transpose1d(F.leaky_relu(o, LRELU_SLOPE), w, W[f"ups.{i}.bias"], stride=s, padding=(k - s) // 2))
        z = None
        for j in range(n_kernels):
            r = f"resblocks.{i * n_kernels + j}"
            x = o
            for d in range(n_dils):
                xt = q(_conv(F.leaky_relu(x, LRELU_SLOPE), W[f"{r}.convs1.{d}.weight"], W[f"{r}.convs1.{d}.bias"],
                             dils[d], reflect))
                xt = _conv(F.leaky_relu(xt, LRELU_SLOPE), W[f"{r}.convs2.{d}.weight"], W[f"{r}.convs2.{d}.bias"], 1,
                           reflect)
                x = q(xt + x)
            z = x if z is None else z + x
        o = q(z / n_kernels)
    o = _conv(F.leaky_relu(o), W["conv_post.weight"], W["conv_post.bias"], 1, reflect)
    return torch.tanh(o)
