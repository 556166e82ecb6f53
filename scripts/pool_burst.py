"""How the pooled graph's per-step time depends on the number of replays queued back to back (diagnostic)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
pipeline = importlib.import_module("xai-audio-deepfakes_b200.pipeline")
ap = pkg.audioprocessor.AudioProcessor(sampling_rate=16000, n_fft=512, hop_length=160, win_length=512, audio_length=4)
B, POOL = 64, 16
pp = pipeline.PipelinedPool(ap, B, POOL, explain_streams=2)
g = torch.Generator(device="cuda").manual_seed(1)
for p in pp.pipes:
    p.wav.copy_(0.1 * torch.randn(B, 64000, generator=g, device="cuda"))
    p.mask.copy_(torch.rand(B, 257, 401, generator=g, device="cuda"))
    p.logits.copy_(2.0 * torch.randn(3, B, generator=g, device="cuda"))
pp.capture()
out = {"trials": pp.trials}


def burst(n, sync_every=0):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        pp.graph.replay()
        if sync_every and (i + 1) % sync_every == 0:
            torch.cuda.synchronize()
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / (n * POOL), 1), round(t_host * 1e6 / n, 1)


for n in (6, 25, 100, 300, 300, 6, 100):
    out[f"burst_{n}_{len(out)}"] = burst(n)
for se in (8, 32):
    out[f"burst_300_sync{se}"] = burst(300, se)
print(json.dumps(out))
