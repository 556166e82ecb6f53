"""Batched load_audio (wav I/O): adv_resample_rows - PCM16 decode + T.Resample's polyphase sinc
filter + pad / crop for a ragged batch in one launch - against the reference's per-file host path
(audioprocessor.py:49-63: torchaudio.load -> T.Resample -> F.pad / crop), i.e. torchaudio's own Resample on the CPU."""
import importlib
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T = pytest.importorskip("torchaudio.transforms")


def write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(pcm.astype("<i2").tobytes())


def host_reference(pcm, sr, target_sr, length):
    x = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
    if sr != target_sr:
        x = T.Resample(orig_freq=sr, new_freq=target_sr)(x)
    if x.shape[0] < length:
        x = torch.nn.functional.pad(x, (0, length - x.shape[0]))
    return x[:length]


def test_load_audio_batch_matches_host_path(pkg, built_lib, tmp_path):
    rs = np.random.RandomState(3)
    specs = [(16000, 80000), (16000, 31234), (44100, 100000), (44100, 300000), (8000, 20000), (22050, 50001), (48000, 7),
             (16000, 123456), (8000, 64000)]
    paths, want = [], []
    ap = pkg.audioprocessor.AudioProcessor(audio_length=5)
    for i, (sr, n) in enumerate(specs):
        pcm = (rs.randn(n) * 6000).clip(-32768, 32767).astype(np.int16)
        p = tmp_path / f"clip{i}_{sr}.wav"
        write_wav(p, pcm, sr)
        paths.append(str(p))
        want.append(host_reference(pcm, sr, 16000, 80000))
    got, sr = ap.load_audio_batch(paths)
    assert sr == 16000 and got.shape == (len(specs), 80000) and got.is_cuda
    want = torch.stack(want)
    assert float((got.cpu() - want).abs().max()) < 2e-6     # values in [-1, 1]; fp32 accumulation order only
    # same-rate clips are a pure decode + pad / crop: bit-exact
    same = [i for i, s in enumerate(specs) if s[0] == 16000]
    assert torch.equal(got.cpu()[same], want[same])
    # and the single-file method agrees with the batch
    one, _ = ap.load_audio(paths[2])
    assert float((one.cpu() - got[2].cpu()).abs().max()) < 2e-6


def test_resample_rows_float_input(pkg, built_lib):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 44100, generator=g)
    flat = x.reshape(-1).cuda()
    got = pkg.audioprocessor.resample_rows(flat, [0, 44100, 88200], [44100, 30000, 44100], 44100, 16000, 16000)
    for b, n in enumerate([44100, 30000, 44100]):
        ref = T.Resample(44100, 16000)(x[b, :n])
        ref = torch.nn.functional.pad(ref, (0, max(0, 16000 - ref.shape[0])))[:16000]
        assert float((got[b].cpu() - ref).abs().max()) < 1e-5
