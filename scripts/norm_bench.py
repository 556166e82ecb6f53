import importlib, os, sys, json
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo") else ".")
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200"); pkg._lib.build(); ops = pkg.ops
g = torch.Generator(device="cuda").manual_seed(0)
POOL = 16
x = [0.1 * torch.randn(64, 64000, generator=g, device="cuda") + 0.01 for _ in range(POOL)]
def graph_timed(fn, reps=20):
    fn(0); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); keep = []
    with torch.cuda.graph(gr):
        for i in range(POOL): keep.append(fn(i))
    for _ in range(2): gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (reps * POOL)
Z = pkg.classifier_embedder.zero_mean_unit_var_norm
t = graph_timed(lambda i: Z(x[i]))
by = 64 * 64000 * 4 * 3
ref = (x[0] - x[0].mean(-1, keepdim=True)) / (x[0].std(-1, keepdim=True) + 1e-7)
print(json.dumps({"zero_mean_unit_var_norm_64x4s": {"us": t * 1e6, "GBps": by / t / 1e9, "frac": by / t / 1e9 / 6537.0, "err": float((Z(x[0]) - ref).abs().max())}}))
