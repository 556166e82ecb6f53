"""Race hunt for the n_fft 1024 streaming kernels (explain5 / istft5 / stft5) and the fused mel front-end: many launches over a
pool of inputs, each result compared BITWISE with the first result for the same input (the kernels are deterministic by
construction), with a co-running memory-bound kernel on another stream perturbing the timing."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("xai-audio-deepfakes_b200"); ops = pkg.ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator(device="cuda").manual_seed(1)
side = torch.cuda.Stream()
junk = torch.empty(64 << 20, device="cuda")
bad_total = 0
GEO = [(322, 644, None, 64, 80000), (322, 644, None, 7, 30000), (256, 700, "hann", 37, 16000), (128, 500, None, 150, 8000)]
for hop, win, kind, B, n in GEO:
    T = 1 + n // hop
    window = torch.hann_window(win) if kind else None
    pool = [(0.1 * torch.randn(B, n, generator=g, device="cuda"), torch.rand(B, 513, T, generator=g, device="cuda")) for _ in range(4)]
    tiles = ops.explain_tiles(1024, hop, win, n, B, length=n)
    def run(k):
        out = (torch.empty(B, n, device="cuda"), torch.empty(B, n, device="cuda"), torch.empty(B, tiles, 4, dtype=torch.float64, device="cuda"))
        ops.explain(pool[k][0], pool[k][1], 1024, hop, win, length=n, window=window, out=out)
        return out
    ref = [run(k) for k in range(4)]
    specs = [ops.stft(w, 1024, hop, win, window=window, want_mag=False, want_phase=False)[0].clone() for w, _ in pool]
    iref = [ops.istft(x, 1024, hop, win, length=n, window=window, return_stats=True) for x in specs]
    iref = [(y.clone(), st.clone()) for y, st in iref]
    bad = [0, 0, 0]
    for it in range(iters):
        k = it % 4
        if it % 3 == 1:
            with torch.cuda.stream(side):
                junk.add_(1.0)
        out = run(k)
        if not (torch.equal(out[0], ref[k][0]) and torch.equal(out[1], ref[k][1]) and torch.equal(out[2], ref[k][2])):
            bad[0] += 1
        y, st = ops.istft(specs[k], 1024, hop, win, length=n, window=window, return_stats=True)
        if not (torch.equal(y, iref[k][0]) and torch.equal(st, iref[k][1])):
            bad[1] += 1
        X = ops.stft(pool[k][0], 1024, hop, win, window=window, want_mag=False, want_phase=False)[0]
        if not torch.equal(torch.view_as_real(X), torch.view_as_real(specs[k])):
            bad[2] += 1
    torch.cuda.synchronize()
    print("hop", hop, "win", win, kind or "rect", "B", B, "n", n, "iters", iters, "mismatches explain / istft / stft", bad)
    bad_total += sum(bad)
mel_mod = importlib.import_module("xai-audio-deepfakes_b200.mel")
mt = mel_mod.MelSpectrogram(16000, 1024, 256, 1024, 80, 0.0, 8000.0, 1.0, "slaney", "slaney", log_compress=True)
wavs = [0.1 * torch.randn(64, 64000, generator=g, device="cuda") for _ in range(4)]
mref = [mt(w).clone() for w in wavs]
bad = 0
for it in range(iters):
    if it % 3 == 1:
        with torch.cuda.stream(side):
            junk.add_(1.0)
    if not torch.equal(mt(wavs[it % 4]), mref[it % 4]):
        bad += 1
torch.cuda.synchronize()
print("mel fused iters", iters, "mismatches", bad)
bad_total += bad
print("TOTAL MISMATCHES", bad_total)
