"""Host-side band compression of the mel filterbank for adv_mel_fused (no GPU): the packed bf16 hi/lo tiles, un-swizzled,
must reproduce the bank to 16 mantissa bits, and the chunk table must cover every non-zero entry."""
import importlib

import numpy as np
import pytest
import torch

pkg = importlib.import_module("xai-audio-deepfakes_b200")


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("args", [dict(n_mels=80), dict(n_mels=80, f_max=8000.0, norm="slaney", mel_scale="slaney"),
                                  dict(n_mels=40, f_min=50.0, f_max=4000.0), dict(n_mels=128)])
def test_pack_band_tiles_reconstructs_bank(args):
    mel = pkg.mel
    fb = mel.melscale_fbanks(513, args.get("f_min", 0.0), args.get("f_max", 8000.0), args["n_mels"], 16000,
                             args.get("norm"), args.get("mel_scale", "htk")).numpy()
    raw, lo_base, table, nyq = mel.pack_band_tiles(fb, args["n_mels"])
    u16 = raw.view(np.uint16)
    assert raw.size == 2 * lo_base and lo_base % 1024 == 0
    rec = np.zeros((512, (args["n_mels"] + 15) // 16 * 16), dtype=np.float64)
    for c in range(8):
        n0, n, off = (int(v) for v in table[c])
        assert n % 8 == 0 and off % 1024 == 0 and n0 % 8 == 0
        for r in range(n):
            for k in range(64):
                pos = r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)
                hi = bf16_to_f32(u16[off // 2 + pos: off // 2 + pos + 1])[0]
                lo = bf16_to_f32(u16[(lo_base + off) // 2 + pos: (lo_base + off) // 2 + pos + 1])[0]
                rec[64 * c + k, n0 + r] = float(hi) + float(lo)
    want = fb[:512].astype(np.float64)
    assert np.abs(rec[:, :args["n_mels"]] - want).max() <= 2.0 ** -16 * np.abs(want).max()
    assert np.all(rec[:, args["n_mels"]:] == 0)
    assert np.array_equal(nyq[:args["n_mels"]], fb[512])
    # band compression: a triangular bank stays far below the dense 8 x 80 x 128 x 2 bytes
    assert raw.size <= 64 * 1024


def test_bf16_bits_round_to_nearest_even():
    x = torch.randn(4096).numpy()
    bits, vals = pkg.mel._bf16_bits(x)
    want = torch.from_numpy(x).to(torch.bfloat16)
    assert np.array_equal(vals, want.float().numpy())
    assert np.array_equal(bits, want.view(torch.int16).numpy().view(np.uint16))


def test_resample_filter_is_torchaudio_default():
    """audioprocessor.resample_filter restates T.Resample's kernel construction; where torchaudio's own (private) builder
    is importable the taps must be bit-identical, and every tap outside the per-phase range must be zero."""
    FF = pytest.importorskip("torchaudio.functional.functional")
    import math
    ap = pkg.audioprocessor
    for o, n in [(44100, 16000), (8000, 16000), (22050, 16000), (48000, 16000)]:
        h, rng, orig, new, width = ap.resample_filter(o, n)
        k, w = FF._get_sinc_resample_kernel(o, n, math.gcd(o, n))
        assert w == width and torch.equal(h, k[:, 0, :])
        for p in range(new):
            lo, hi = int(rng[p, 0]), int(rng[p, 1])
            assert not h[p, :lo].any() and not h[p, hi:].any()
