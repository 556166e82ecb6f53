import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(64, 80000, generator=g, device="cuda")
for _ in range(3):
    X, mag, ph = ops.stft(wav, 1024, 322, 644)
    X2, _, _ = ops.stft(wav, 1024, 322, 644, want_mag=False, want_phase=False)
torch.cuda.synchronize(); print("ok")
