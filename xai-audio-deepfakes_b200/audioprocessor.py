"""Drop-in for the reference's ``audioprocessor.AudioProcessor`` (audioprocessor.py:22-131).

Same constructor keywords, same method names, argument meaning, return shapes and error texts;
the arithmetic runs in hand-written sm_100a kernels (libaddvisor_sm100.so) instead of
torch.stft / torch.istft / ATen elementwise ops.  Differences a caller can observe:

* results live on the current CUDA device (the reference uses ``Accelerator().device``);
* ``magnitude`` and ``phase`` come back with the same frame-major strides as ``X`` (the reference's
  ``.angle()`` happens to be contiguous); values are equal within fp32 round-off;
* extra methods ``explain`` / ``explain_from_stft`` expose the fused mask-apply + double iSTFT
  (+ normaliser) that ``run_addvisor_metrics`` needs, so masked spectra never reach HBM.

The SSL classifier stays the reference's own torch module: ``extract_features`` runs our normaliser
kernel and then calls whatever ``wav2vec2`` module ``classifier_embedder`` holds.
"""
from __future__ import annotations

import math
import wave as _wave

import numpy as np
import torch

from . import classifier_embedder as _ce
from . import ops


def _device():
    ops._lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def resample_filter(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """Polyphase sinc filter of ``torchaudio.transforms.Resample(orig_freq, new_freq)`` with its defaults
    (sinc_interp_hann) - the transform ``load_audio`` applies (audioprocessor.py:53-55) - built with the same torch
    expressions, dtypes included (float32 phase offsets, float64 tap positions), so the taps are bit-identical.
    Returns (h float32 [new][2*width + orig], range int32 [new][2] = non-zero tap span per phase, orig, new, width)
    with orig / new divided by their gcd.  A host-side constant like the mel filterbank."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    h = kernels.to(torch.float32)[:, 0, :].contiguous()
    nz = h != 0
    lo = torch.where(nz.any(1), nz.float().argmax(1), torch.zeros(new, dtype=torch.long))
    hi = torch.where(nz.any(1), h.shape[1] - nz.flip(1).float().argmax(1), torch.zeros(new, dtype=torch.long))
    rng = torch.stack([lo, hi], 1).to(torch.int32).contiguous()
    return h, rng, orig, new, width


def resample_rows(src, offsets, lengths, orig_freq, new_freq, n_out):
    """Clips ``src[offsets[b] : offsets[b] + lengths[b]]`` (CUDA int16 PCM or float32, mono) -> float32 [B, n_out]:
    decoded (int16 / 32768), resampled orig_freq -> new_freq like ``T.Resample``, zero-padded / cropped to n_out
    (audioprocessor.py:49-63 for a whole batch in one launch, ``adv_resample_rows``)."""
    lib_, L = ops._lib.lib(), ops._lib
    dev = src.device
    offs = torch.as_tensor(offsets, dtype=torch.int64).to(dev)
    lens = torch.as_tensor(lengths, dtype=torch.int32).to(dev)
    B = int(lens.numel())
    out = torch.empty((B, int(n_out)), dtype=torch.float32, device=dev)
    if int(orig_freq) == int(new_freq):
        h = rng = None
        orig = new = 1
        width = 0
    else:
        key = (int(orig_freq), int(new_freq), str(dev))
        if key not in _RESAMPLE_TABLES:
            h, rng, orig, new, width = resample_filter(orig_freq, new_freq)
            _RESAMPLE_TABLES[key] = (h.to(dev), rng.to(dev), orig, new, width)
        h, rng, orig, new, width = _RESAMPLE_TABLES[key]
    if src.dtype not in (torch.int16, torch.float32):
        raise TypeError("resample_rows expects int16 PCM or float32 samples")
    with torch.cuda.device(dev):
        L.check(lib_.adv_resample_rows(L.ptr(src), int(src.dtype == torch.int16), L.ptr(offs), L.ptr(lens), B, orig, new,
                                       width, L.ptr(h), L.ptr(rng), int(n_out), L.ptr(out), L.stream_ptr()),
                "adv_resample_rows")
    return out


_RESAMPLE_TABLES = {}


class AudioProcessor:
    def __init__(self, sampling_rate=16000, n_fft=1024, hop_length=322, win_length=644, n_mels=80,
                 audio_length=5):
        self.sampling_rate = sampling_rate
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.n_mels = n_mels
        self.audio_length = audio_length
        self._mel = None  # built on first use (audioprocessor.py:38-44 builds it eagerly, never calls it)

    # ------------------------------------------------------------------ audioprocessor.py:38-44
    def mel_transform(self, waveform):
        """torchaudio ``MelSpectrogram(sample_rate, n_fft, hop, win, n_mels)`` defaults: hann(win)
        window, power 2, HTK mel scale, no filterbank normalisation -> [., n_mels, T]."""
        from .mel import MelSpectrogram
        if self._mel is None:
            self._mel = MelSpectrogram(self.sampling_rate, self.n_fft, self.hop_length, self.win_length,
                                       self.n_mels)
        return self._mel(waveform)

    # ------------------------------------------------------------------ audioprocessor.py:49-63
    def load_audio(self, audio_path, target_sr=16000):
        """Host I/O (not on the GPU path): mono wav -> float tensor padded / cropped to
        ``audio_length * target_sr`` samples."""
        try:
            import torchaudio
            audio, sr = torchaudio.load(audio_path)
        except Exception:  # torchaudio.load needs torchcodec in this image; PCM16 wavs need neither
            with _wave.open(audio_path, "rb") as f:
                sr, ch, width = f.getframerate(), f.getnchannels(), f.getsampwidth()
                if width != 2:
                    raise
                pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").reshape(-1, ch).T
            audio = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
        if audio.ndim > 1:
            audio = audio.squeeze(0)
        if sr != target_sr:
            import torchaudio.transforms as T
            audio = T.Resample(orig_freq=sr, new_freq=target_sr)(audio)
        return self._fit(audio, int(self.audio_length * target_sr)), target_sr

    def load_audio_batch(self, audio_paths, target_sr=16000):
        """``load_audio`` for a list of files -> (float32 CUDA tensor [B, audio_length * target_sr], target_sr).
        Host side: the PCM16 payloads are read with the stdlib ``wave`` module into ONE pinned buffer and copied once;
        device side: decode, per-source-rate sinc resampling (T.Resample's filter) and pad / crop for all clips of a
        rate in one launch - instead of B x (torchaudio.load, Resample, F.pad) on the host and B copies.
        (batched wav I/O.)"""
        n_out = int(self.audio_length * target_sr)
        meta, chunks, pos = [], [], 0
        for path in audio_paths:
            with _wave.open(path, "rb") as f:
                sr, ch, width = f.getframerate(), f.getnchannels(), f.getsampwidth()
                if width != 2 or ch != 1:   # the reference's squeeze(0) only handles mono; 16-bit PCM is what it ships
                    raise ValueError(f"{path}: expected mono 16-bit PCM (got {ch} channels, {8 * width} bits)")
                pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
            meta.append((sr, pos, pcm.size))
            chunks.append(pcm)
            pos += pcm.size
        dev = _device()
        host = torch.empty(max(pos, 1), dtype=torch.int16).pin_memory()
        if pos:
            host[:pos] = torch.from_numpy(np.concatenate(chunks))
        src = host.to(dev, non_blocking=True)
        out = torch.empty((len(meta), n_out), dtype=torch.float32, device=dev)
        for sr in sorted({m[0] for m in meta}):
            rows = [i for i, m in enumerate(meta) if m[0] == sr]
            part = resample_rows(src, [meta[i][1] for i in rows], [meta[i][2] for i in rows], sr, target_sr, n_out)
            if len(rows) == len(meta):
                out = part
            else:
                out[torch.as_tensor(rows, device=dev)] = part
        return out, target_sr

    @staticmethod
    def _fit(x, length):
        cur = x.shape[-1]
        if cur < length:
            return torch.nn.functional.pad(x, (0, length - cur))
        return x[..., :length]

    # ------------------------------------------------------------------ audioprocessor.py:69-77
    def extract_features(self, waveforms):
        """normalise (our kernel) -> wav2vec2 (reference torch module) -> hidden_states[9].squeeze(0)."""
        audio = _ce.zero_mean_unit_var_norm(waveforms)
        net = _ce.get_wav2vec2()
        output = net(audio, output_hidden_states=True)
        return output.hidden_states[9].squeeze(0)

    # ------------------------------------------------------------------ audioprocessor.py:82-112
    def compute_stft(self, waveform):
        if waveform.dim() not in (1, 2):
            raise ValueError("waveform must be 1D (single) or 2D (batched waveforms)")
        single = waveform.dim() == 1
        wav = self._fit(waveform, int(self.audio_length * self.sampling_rate))
        X, mag, phase = ops.stft(wav.unsqueeze(0) if single else wav, self.n_fft, self.hop_length,
                                 self.win_length)
        if single:
            return X[0], mag[0], phase[0]
        return X, mag, phase

    # ------------------------------------------------------------------ audioprocessor.py:117-131
    def compute_invert_stft(self, spectrogram):
        if not torch.is_complex(spectrogram):
            raise ValueError("ISTFT expects complex input!")
        expected_length = self.audio_length * self.sampling_rate
        single = spectrogram.dim() == 2
        spec = spectrogram.unsqueeze(0) if single else spectrogram
        out = ops.istft(spec, self.n_fft, self.hop_length, self.win_length, length=expected_length)
        return out[0] if single else out

    # ------------------------------------------------------------------ fused path (ours)
    def explain(self, waveforms, mask, mode="log1p", normalize=False, outside="drop"):
        """waveforms [B,n] + mask [B,F',T'] -> (relevant, irrelevant) waveforms [B, audio_length*sr].

        One kernel does compute_stft -> ``expm1(mask*log1p(mag)) * exp(1j*phase)`` (and the 1-mask
        twin) -> compute_invert_stft twice (LMAC_metrics.py:136-157); ``mode="linear"`` is the
        training-loss variant (loss_function.py:36-47).  ``normalize=True`` also applies
        zero_mean_unit_var_norm, i.e. returns what extract_features feeds to wav2vec2."""
        if waveforms.dim() == 1:
            waveforms, mask = waveforms.unsqueeze(0), (mask.unsqueeze(0) if mask.dim() == 2 else mask)
        n = int(self.audio_length * self.sampling_rate)
        wav = self._fit(waveforms, n)
        return ops.explain(wav, mask, self.n_fft, self.hop_length, self.win_length, length=n, mode=mode,
                           normalize=normalize, outside=outside)

    def explain_from_stft(self, spectrogram, mask, mode="log1p", normalize=False, outside="drop"):
        """Same, starting from ``compute_stft``'s complex output (collate_fn already has it)."""
        if not torch.is_complex(spectrogram):
            raise ValueError("ISTFT expects complex input!")
        n = int(self.audio_length * self.sampling_rate)
        return ops.explain_spec(spectrogram, mask, self.n_fft, self.hop_length, self.win_length, length=n,
                                mode=mode, normalize=normalize, outside=outside)
