"""BASELINE.json configs[2]: HiFi-GAN resynthesis of 80-bin mels for B x 4 s clips (bf16 tcgen05 conv path).
Prints one JSON line: clips/s, achieved TFLOP/s (dense FLOPs of SURVEY.md 8(a12): 160.3 GFLOP/clip) and
the fraction of the measured bf16 peak."""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=251)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--pipeline", default="tma")
ap.add_argument("--fuse", default="auto")
ap.add_argument("--epilogue", default="tma", help="auto / tma / direct: epilogue of the slab conv kernels (HifiganGenerator)")
args = ap.parse_args()
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
H = pkg.hifigan
gen = H.HifiganGenerator(H.init_weights(seed=0, std=0.01), pipeline=args.pipeline, fuse=args.fuse, epilogue=args.epilogue)
g = torch.Generator(device="cuda").manual_seed(1234)
mel = -4 + 2 * torch.randn(args.batch, 80, args.frames, generator=g, device="cuda")
gen.decode_batch(mel)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.iters):
    wav = gen.decode_batch(mel)
b.record()
torch.cuda.synchronize()
t = a.elapsed_time(b) * 1e-3 / args.iters
flops = H.HifiganGenerator.flops_per_clip(args.frames) * args.batch
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1384.0}
peak = peaks["bf16_tflops_sustained"]
print(json.dumps({"metric": "vocoded clips/s", "value": args.batch / t, "unit": "clips/s", "batch": args.batch,
                  "frames": args.frames, "pipeline": args.pipeline, "fuse": args.fuse, "epilogue": args.epilogue, "epilogue_choices": {"tma": sum(gen._epi.values()), "direct": len(gen._epi) - sum(gen._epi.values())}, "launches": gen.launches // (args.iters + 1), "s_per_batch": t, "tflops": flops / t / 1e12,
                  "roofline": {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peak, "unit": "TFLOP/s",
                               "frac": flops / t / 1e12 / peak}, "out_shape": list(wav.shape)}))
