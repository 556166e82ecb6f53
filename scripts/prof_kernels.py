"""Launch each transform kernel a few times on the BASELINE cfg-2 shapes (for ncu captures)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ops = pkg.ops
B, n, n_fft, hop, win = 64, 64000, 512, 160, 512
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(B, n, generator=g, device="cuda")
mask = torch.rand(B, 257, 401, generator=g, device="cuda")
for it in range(3):
    X, mag, ph = ops.stft(wav, n_fft, hop, win)
    X2, _, _ = ops.stft(wav, n_fft, hop, win, want_mag=False, want_phase=False)
    y = ops.istft(X, n_fft, hop, win, length=n)
    rel, irr = ops.explain(wav, mask, n_fft, hop, win, length=n, normalize=True)
    r2, i2 = ops.explain_spec(X, mask, n_fft, hop, win, length=n)
torch.cuda.synchronize()
print("ok", float((y - wav).abs().max()))
