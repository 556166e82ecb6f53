"""mel front-end (hifigan.py:163-178 geometry; --default for audioprocessor.py:38-44) on B x 4 s clips: the fused single
launch against the two-launch path, timed from a CUDA graph of POOL calls over rotating inputs (CUDA events)."""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--n", type=int, default=64000)
ap.add_argument("--pool", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--default", action="store_true")
ap.add_argument("--tag", default="")
args = ap.parse_args()
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
mel_mod = importlib.import_module("xai-audio-deepfakes_b200.mel")
g = torch.Generator(device="cuda").manual_seed(0)
wavs = [0.1 * torch.randn(args.batch, args.n, generator=g, device="cuda") for _ in range(args.pool)]
if args.default:
    mt = mel_mod.MelSpectrogram(16000, 1024, 322, 644, 80)
else:
    mt = mel_mod.MelSpectrogram(16000, 1024, 256, 1024, 80, 0.0, 8000.0, 1.0, "slaney", "slaney", log_compress=True)


def graph_time(fn):
    fn(0)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(args.pool):
            fn(i)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (args.reps * args.pool)


out = {"tag": args.tag, "batch": args.batch, "n": args.n, "hop": mt.hop_length, "win": mt.win_length}
y1 = mt(wavs[0])
out["path"] = mt.last_path
out["fused_us"] = graph_time(lambda i: mt(wavs[i]))
mt.fused = False
y2 = mt(wavs[0])
out["two_launch_us"] = graph_time(lambda i: mt(wavs[i]))
out["max_abs_diff"] = float((y1 - y2).abs().max())
out["max_abs"] = float(y2.abs().max())
print(json.dumps(out))
