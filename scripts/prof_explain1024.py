"""reference-default geometry (n_fft 1024 / hop 322 / win 644, 64 x 5 s clips): a few explain / istft calls for ncu captures"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(64, 80000, generator=g, device="cuda")
mask = torch.rand(64, 513, 249, generator=g, device="cuda")
X, _, _ = ops.stft(wav, 1024, 322, 644, want_mag=False, want_phase=False)
for _ in range(3):
    rel, irr = ops.explain(wav, mask, 1024, 322, 644, length=80000)
    y = ops.istft(X, 1024, 322, 644, length=80000)
torch.cuda.synchronize(); print("ok")
