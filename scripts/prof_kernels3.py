"""Launch each generation-3 transform kernel on the BASELINE cfg-2 shapes and on the reference-default geometry
(for ncu captures: 3 iterations x 7 matching launches; capture the last iteration with -s 14 -c 7)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ops = pkg.ops
B = int(os.environ.get("PROF_B", "64"))
g = torch.Generator(device="cuda").manual_seed(0)
cfgs = [(64000, 512, 160, 512), (80000, 1024, 322, 644)]
data = []
for n, n_fft, hop, win in cfgs:
    wav = 0.1 * torch.randn(B, n, generator=g, device="cuda")
    mask = torch.rand(B, n_fft // 2 + 1, 1 + n // hop, generator=g, device="cuda")
    data.append((wav, mask))
for it in range(3):
    for (n, n_fft, hop, win), (wav, mask) in zip(cfgs, data):
        if n_fft == 512:
            X2, _, _ = ops.stft(wav, n_fft, hop, win, want_mag=False, want_phase=False)
        X, mag, ph = ops.stft(wav, n_fft, hop, win)
        y = ops.istft(X, n_fft, hop, win, length=n)
        rel, irr = ops.explain(wav, mask, n_fft, hop, win, length=n, normalize=True)
torch.cuda.synchronize()
print("ok", float((y - wav).abs().max()))
