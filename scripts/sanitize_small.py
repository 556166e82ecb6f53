"""Small shapes through every new kernel family (for compute-sanitizer memcheck: one tool per gpurun call)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ops, H = pkg.ops, pkg.hifigan
g = torch.Generator().manual_seed(0)
for (n_fft, hop, win, n, B) in [(512, 160, 512, 8000, 3), (512, 100, 400, 5000, 2), (1024, 322, 644, 9660, 2)]:
    wav = 0.1 * torch.randn(B, n, generator=g)
    T, F = 1 + n // hop, n_fft // 2 + 1
    mask = torch.rand(B, F, T, generator=g)
    window = None if win == n_fft else torch.hann_window(win)
    X, mag, ph = ops.stft(wav, n_fft, hop, win, window=window)
    y = ops.istft(X, n_fft, hop, win, length=n, window=window)
    rel, irr = ops.explain(wav, mask, n_fft, hop, win, length=n, window=window, normalize=True)
    m = mask.cuda().requires_grad_(True)
    r2, i2 = ops.explain_linear(X, m, n_fft, hop, win, length=n, window=window)
    (r2.sum() + 2 * i2.sum()).backward()
    print(n_fft, hop, float((y.cpu() - wav).abs().max()), float(m.grad.abs().max()))
print("shift", int(H.align_shift(torch.randn(3000, generator=g), torch.randn(2500, generator=g)).item()))
print("bands", tuple(H.band_swapped_waveforms(torch.randn(5000, generator=g), torch.randn(5000, generator=g)).shape))
mel = pkg.audioprocessor.AudioProcessor(sampling_rate=8000, audio_length=1).mel_transform(torch.randn(2, 8000, generator=g))
print("mel", tuple(mel.shape))
W = H.init_weights(seed=1, std=0.03)
for fuse in ("always", "never"):
    wavv = H.HifiganGenerator(W, fuse=fuse).decode_batch(-4 + 2 * torch.randn(1, 80, 6, generator=g))
    print("voc", fuse, tuple(wavv.shape), float(wavv.abs().max()))
torch.cuda.synchronize()
print("done")
