"""Where does the streaming explain kernel differ from torch?  (debug helper)"""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
hop = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6400
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(B, n, generator=g, device="cuda")
mask = torch.rand(B, 257, 1 + n // hop, generator=g, device="cuda")
w = torch.ones(512, device="cuda")
X = torch.stft(wav, 512, hop_length=hop, win_length=512, window=w, return_complex=True)
lm, ph = torch.log1p(X.abs()), X.angle()
rr = torch.istft(torch.polar(torch.expm1(mask * lm), ph), 512, hop_length=hop, win_length=512, window=w, length=n)
ir = torch.istft(torch.polar(torch.expm1((1 - mask) * lm), ph), 512, hop_length=hop, win_length=512, window=w, length=n)
rel, irr = ops.explain(wav, mask, 512, hop, 512, length=n)
for name, got, want in (("rel", rel, rr), ("irr", irr, ir)):
  d = (got - want).abs() / want.abs().max()
  print("hop", hop, name, "max err", float(d.max()))
  for b in range(B):
    bad = torch.nonzero(d[b] > 1e-4).flatten()
    if bad.numel():
        rows = sorted(set((bad // 32).tolist()))
        print(" clip", b, "bad samples", bad.numel(), "first", int(bad[0]), "last", int(bad[-1]), "rows of 32:", rows[:40])
