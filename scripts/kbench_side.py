"""Side kernels of the path on BASELINE cfg-2 shapes: time per call (CUDA events, rotating inputs > L2 where the
tensors are small) and achieved GB/s against the algorithmic bytes of DESIGN.md section 4."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("xai-audio-deepfakes_b200")
pkg._lib.build()
ops = pkg.ops
H = pkg.hifigan
B, n, F, T = 64, 64000, 257, 401
g = torch.Generator(device="cuda").manual_seed(0)
POOL = 8


def timed(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


out = {}
wav = [0.1 * torch.randn(B, n, generator=g, device="cuda") for _ in range(POOL)]
mask = [torch.rand(B, F, T, generator=g, device="cuda") for _ in range(POOL)]
spec = [ops.stft(w, 512, 160, 512, want_mag=True, want_phase=True) for w in wav]

t = timed(lambda i: ops.mask_apply(spec[i % POOL][1], spec[i % POOL][2], mask[i % POOL]))
by = B * F * T * (4 * 3 + 16)
out["mask_apply"] = {"us": t * 1e6, "GBps": by / t / 1e9}

y1 = [torch.randn(16, 32, 256, 400, generator=g, device="cuda") for _ in range(2)]
wts, bias = torch.randn(32, device="cuda"), torch.zeros(1, device="cuda")
t = timed(lambda i: ops.mask_head(y1[i % 2], wts, bias))
by = 16 * 33 * 256 * 400 * 4
out["mask_head_b16"] = {"us": t * 1e6, "GBps": by / t / 1e9}

attr = [torch.randn(B, n, generator=g, device="cuda") for _ in range(POOL)]
t = timed(lambda i: ops.td_mask(wav[i % POOL], attr[i % POOL]))
by = B * n * 4 * (2 + 1 + 3)
out["td_mask"] = {"us": t * 1e6, "GBps": by / t / 1e9}

X = [s[0] for s in spec]
t = timed(lambda i: ops.band_swap(X[i % POOL], X[(i + 1) % POOL], 64, 128))
by = B * F * T * 8 * 3
out["band_swap"] = {"us": t * 1e6, "GBps": by / t / 1e9}

t = timed(lambda i: ops.normalize_(wav[i % POOL]))
by = B * n * 4 * 3
out["normalize(row_stats+scale)"] = {"us": t * 1e6, "GBps": by / t / 1e9}

lg = torch.randn(3, 100000, generator=g, device="cuda")
ws = ops.LmacWorkspace(100000, torch.device("cuda"))
t = timed(lambda i: ops.lmac(lg[0], lg[1], lg[2], is_logit=True, want_scores=False, workspace=ws))
out["lmac_100k"] = {"us": t * 1e6, "GBps": 100000 * 12 / t / 1e9}

ref, deg = torch.randn(64000, generator=g, device="cuda"), torch.randn(66816, generator=g, device="cuda")
t = timed(lambda i: H.align_shift(ref, deg, method="direct"), reps=5)
fma = 0.5 * (64000 + 66816 + 1) * 66816  # about half of the lag x tap rectangle meets non-zero samples
out["xcorr_4s_pair_direct"] = {"us": t * 1e6, "TFMA/s": fma / t / 1e12}
t = timed(lambda i: H.align_shift(ref, deg), reps=20)
out["xcorr_4s_pair_fft"] = {"us": t * 1e6, "launches": "2 fills + 2 copies (torch) + 2 stft + mac + istft + argmax",
                            "same_shift": int(H.align_shift(ref, deg)) == int(H.align_shift(ref, deg, method="direct"))}
print(json.dumps(out, indent=1))
