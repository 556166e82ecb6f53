"""mel front-end (hifigan.py:163-178 geometry) on 64 x 4 s clips: a few calls for ncu / event timing"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
mel_mod = importlib.import_module("xai-audio-deepfakes_b200.mel")
g = torch.Generator(device="cuda").manual_seed(0)
wav = 0.1 * torch.randn(64, 64000, generator=g, device="cuda")
mt = mel_mod.MelSpectrogram(16000, 1024, 256, 1024, 80, 0.0, 8000.0, 1.0, "slaney", "slaney", log_compress=True)
for _ in range(3):
    out = mt(wav)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    out = mt(wav)
b.record()
torch.cuda.synchronize()
print("mel front-end: %.1f us per call, out %s" % (a.elapsed_time(b) * 1e3 / 20, tuple(out.shape)))
