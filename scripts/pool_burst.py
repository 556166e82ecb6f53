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
out = {"burst_6": round(pp.burst_us_per_step(), 1)}


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


import subprocess
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap", "--format=csv,noheader",
                        "-lms", "20"], stdout=subprocess.PIPE, text=True)
for n in (6, 25, 100, 300, 1000, 6, 100):
    out[f"burst_{n}_{len(out)}"] = burst(n)   # (us per step on the device, host us per replay)
for se in (8, 32):
    out[f"burst_300_sync{se}"] = burst(300, se)
smi.terminate()
lines = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
pw = sorted(float(l.split(",")[0].split()[0]) for l in lines)
ck = sorted(float(l.split(",")[1].split()[0]) for l in lines)
out["power_w"] = {"samples": len(pw), "median": pw[len(pw) // 2] if pw else None, "max": pw[-1] if pw else None}
out["sm_mhz"] = {"min": ck[0] if ck else None, "median": ck[len(ck) // 2] if ck else None}
out["power_cap_active_samples"] = sum("Active" in l and "Not Active" not in l for l in lines)
print(json.dumps(out))
