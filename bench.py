#!/usr/bin/env python
"""Benchmark of the ADDvisor explanation-evaluation hot path (BASELINE.json metric:
"explained clips/s (4 s @ 16 kHz) STFT-mask-iSTFT + LMAC metrics").

Workload = BASELINE.json configs[1]: synthetic batch of 64 x 4 s clips @ 16 kHz, n_fft 512, hop 160,
rectangular win 512, mask on the [257, 401] grid.  One step = one batch through

    explain (STFT -> mask / 1-mask on log1p magnitude -> 2 x iSTFT)  -> normalise x2
    -> [SSL classifier: the reference's torch module, NOT part of the timed path; its logits are
        synthetic N(0, 2^2)] -> lmac_reduce (sigmoid + FF/Fid/AD/AI/AG sums)

value : clips/s with inputs resident in HBM (a pool of batches larger than L2 is rotated).
e2e   : same step through the public API with pinned HOST inputs (H2D inside the timed region) and
        the six metric sums read back to the host every step.
roofline : the dominant kernel (fused explain) timed alone with CUDA events on its launch stream.
cpu_baseline / --impl reference : the oracle port of the reference's torch-CPU path on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
For N > 1 launch with torch.distributed.run (one rank per GPU); ranks shard the clips, the only
collective is one all-reduce of the six metric sums at the end of the evaluation.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(sampling_rate=16000, n_fft=512, hop_length=160, win_length=512, audio_length=4)
BATCH = 64
N = CFG["sampling_rate"] * CFG["audio_length"]
T = 1 + N // CFG["hop_length"]
F = CFG["n_fft"] // 2 + 1
# algorithmic bytes per clip (SURVEY.md 8d / DESIGN.md): fused explain reads the wave (4N) and the mask
# (4FT) and writes two waves (2*4N); stft = 4N + 8FT; istft = 8FT + 4N
BYTES_EXPLAIN = 4 * N + 4 * F * T + 2 * 4 * N
BYTES_STFT = 4 * N + 8 * F * T
BYTES_ISTFT = 8 * F * T + 4 * N
METRIC = "explained clips/s (4s@16kHz) STFT-mask-iSTFT+LMAC metrics"
POOL = 16  # rotating input/output sets: 16 x 75.6 MB = 1.2 GB >> 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_path_step(R, wav, mask, logits):
    """The reference's CPU path for one batch (oracle port, torch CPU fp32)."""
    rel, irr = R.explain(wav, mask, mode="log1p", normalize=True, **CFG)
    pr = torch.sigmoid(logits).unsqueeze(-1)
    return rel, irr, R.lmac_sums(pr[0], pr[1], pr[2])


def synth_host(seed, batch=BATCH):
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(batch, N, generator=g)
    mask = torch.rand(batch, F, T, generator=g)
    logits = 2.0 * torch.randn(3, batch, generator=g)
    return wav, mask, logits


def workload_config(n_gpus):
    """The ``config`` object both arms print (identical keys and values: it names the workload, nothing run-specific)."""
    return {"workload": "configs[1]: 64 x 4 s clips @16 kHz per GPU-step, n_fft 512 hop 160 win 512 (rectangular), "
                        "STFT -> log1p mask / 1-mask -> 2 x iSTFT -> normalise x2 -> LMAC sums on given classifier "
                        "logits (SSL model = the reference's torch module, not timed)",
            "batch_per_gpu": BATCH, "clip_seconds": CFG["audio_length"], "sample_rate": CFG["sampling_rate"],
            "n_fft": CFG["n_fft"], "hop": CFG["hop_length"], "win": CFG["win_length"], "mask": "log1p",
            "parallelism": f"dp{n_gpus}",
            "l2": f"inputs larger than L2: {POOL} rotating buffer sets (1.2 GB) > 126 MB"}


def reference_stepper():
    """(step(wav, mask, logits), kind, note): the UNMODIFIED reference's own code from baseline/_ref (a verbatim,
    git-ignored copy made by __graft_entry__.build() where /root/reference exists; it travels to the GPU box) driven
    through oracle/ref_loader.ReferencePath; the oracle port only if that copy is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(ref_dir, "audioprocessor.py")) and os.path.exists(os.path.join(ref_dir, "LMAC_metrics.py")):
        try:
            from oracle.ref_loader import ReferencePath
            rp = ReferencePath(ref_dir, **CFG)
            return rp.step, "reference", ("baseline/_ref: AudioProcessor.compute_stft / compute_invert_stft, "
                                          f"LMAC_metrics.py lines {rp.lines[0]}-{rp.lines[-1]} and compute_* run verbatim")
        except Exception as e:  # fall through to the port, say why
            note = f"baseline/_ref failed to load ({e!r}); "
    else:
        note = "baseline/_ref absent; "
    from oracle import ref_path as R
    return (lambda w, m, l: cpu_path_step(R, w, m, l)), "port", note + "oracle/ref_path.py (same torch-CPU calls, bit-identical)"


def run_reference(args, out=sys.stdout):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    step, kind, note = reference_stepper()
    wav, mask, logits = synth_host(1234)
    steps, warm = max(1, min(args.steps, 40)), max(1, min(args.warmup, 40))
    for _ in range(warm):
        step(wav, mask, logits)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(wav, mask, logits)
    dt = (time.perf_counter() - t0) / steps
    val = BATCH / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} steps x {BATCH} clips, torch {torch.__version__} CPU fp32; {note}"},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def time_loop(fn, iters):
    """CUDA-event time of `iters` calls on the current stream, synchronised on both sides (seconds)."""
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version line at
    init whatever NCCL_DEBUG says), so file descriptor 1 is pointed at stderr for the rest of the process and the
    JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=4000)
    ap_.add_argument("--warmup", type=int, default=48)
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--schedule", default="pooled", choices=["pooled", "streams"],
                     help="pooled: one CUDA graph over the pool of batches with the normaliser / metric reduction of "
                          "batch j on a high-priority stream next to explain(j+1); streams: one graph per batch, "
                          "replayed round-robin on --streams streams")
    ap_.add_argument("--streams", type=int, default=2,
                     help="independent batches are replayed round-robin on this many CUDA streams so that the tail "
                          "wave of one step's kernels overlaps the head of the next step's")
    args = ap_.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # keeps stdout to the one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = importlib.import_module("xai-audio-deepfakes_b200")
    pkg._lib.build()
    # one process per GPU: keep each rank (and the pinned host buffers of the end-to-end leg) on the cores next to its
    # GPU.  Single-process runs stay unbound: the CPU baseline wants every host core.
    affinity = pkg.distributed.bind_host_to_gpu(local) if world > 1 and os.environ.get("ADV_NO_BIND") is None else None
    ops = pkg.ops
    from importlib import import_module
    pipeline = import_module("xai-audio-deepfakes_b200.pipeline")
    ap = pkg.audioprocessor.AudioProcessor(**CFG)
    steps, warm = max(1, args.steps), max(3, args.warmup)   # (>= 3 warm-up steps: the timing rules' floor)

    # ---- device-resident pool (each rank owns its shard of clips: weak scaling, B clips per rank-step)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    pooled = args.schedule == "pooled"
    if pooled:
        # one CUDA graph over the whole pool: explain(j) on alternating streams, normalise(j) + lmac(j) behind it on
        # a high-priority stream so that they run next to explain(j+1) (pipeline.PipelinedPool)
        pp = pipeline.PipelinedPool(ap, BATCH, POOL, explain_streams=args.streams)
        pool = pp.pipes
    else:
        pool = [pipeline.ExplainPipeline(ap, BATCH, use_graph=True, accumulate=True) for _ in range(POOL)]
    for p in pool:
        p.wav.copy_(0.1 * torch.randn(BATCH, N, generator=gen, device="cuda"))
        p.mask.copy_(torch.rand(BATCH, F, T, generator=gen, device="cuda"))
        p.logits.copy_(2.0 * torch.randn(3, BATCH, generator=gen, device="cuda"))
    ns = max(1, min(args.streams, POOL))
    streams = [torch.cuda.Stream() for _ in range(ns)]
    burst_us = None
    if pooled:
        pp.capture()
        burst_us = round(pp.burst_us_per_step(), 1)  # 96 steps from a cool GPU, before the sustained timed region

    def run_steps(k):
        """k steps on the current stream: whole pools replay the pooled graph, the remainder replays a tail graph over the
        first k % POOL buffer sets (captured on first use, i.e. during warm-up) - every step takes the same pipelined
        schedule whatever --steps is.  2 launches of ours per step (explain; normaliser + metric CTA); metric sums accumulate inside the metric CTA."""
        if pooled:
            if k <= 4 * POOL:      # a short run is ONE graph of exactly k steps (no second launch, no un-pipelined seam)
                pp.replay_tail(k)
                return
            for _ in range(k // POOL):
                pp.replay()
            if k % POOL:
                pp.replay_tail(k % POOL)
            return
        for i in range(k):
            j = i % POOL
            with torch.cuda.stream(streams[j % ns]):  # buffer set j always runs on stream j % ns: no cross-stream hazards
                pool[j].step()

    def fork(ev):   # side streams start after `ev` (recorded on the main stream)
        for st in streams:
            st.wait_event(ev)

    def join():     # main stream continues after everything queued on the side streams
        main = torch.cuda.current_stream()
        for st in streams:
            e = torch.cuda.Event()
            e.record(st)
            main.wait_event(e)

    sampler = ClockSampler(local) if rank == 0 else None
    # pre-heat (NOT the warm-up): ~0.7 s of the same steps so that the clock sampler sees the GPU under load and the
    # clocks have ramped; reported separately as preheat_s / preheat_steps
    t_w = time.perf_counter()
    preheat_steps = 0
    while time.perf_counter() - t_w < 0.7:
        run_steps(POOL)
        preheat_steps += POOL
        torch.cuda.synchronize()
    preheat_s = time.perf_counter() - t_w
    # warm-up: exactly --warmup steps of exactly what the timed region runs (same split into pooled / tail replays), then
    # the epilogue once (stack + sum, and the NCCL all-reduce): CUDA loads kernels lazily on first use and NCCL sets its
    # channels up on the first collective - tens of milliseconds that do not belong to the steps
    if pooled:   # capture the graphs the timed region will replay (capture synchronises the device)
        if steps <= 4 * POOL:
            pp.replay_tail(steps)
        elif steps % POOL:
            pp.replay_tail(steps % POOL)
    run_steps(warm)
    warm_total = torch.stack([p.sums for p in pool]).sum(dim=0)
    if world > 1:
        dist.all_reduce(warm_total, op=dist.ReduceOp.SUM)
    del warm_total
    torch.cuda.synchronize()
    for p in pool:
        p.sums.zero_()
        p.launches = 0
    if pooled:
        pp.launches = 0

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fork(a)
    run_steps(steps)
    join()
    total = torch.stack([p.sums for p in pool]).sum(dim=0)
    if world > 1:  # the one real exchange: six float64 sums, once per evaluation
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    b.record()
    torch.cuda.synchronize()
    elapsed = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed)
    launches = sum(p.launches for p in pool) + (pp.launches if pooled else 0)
    value = world * BATCH * steps / elapsed
    metrics = pkg.LMAC_metrics.finalize(total)

    # ---- dominant kernel alone (CUDA events on its launch stream), pool rotated the same way
    peak, peak_src = peaks()

    def graph_time(fn, reps):
        """avg seconds per launch of fn(i), i rotating over the pool, replayed from one CUDA graph"""
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(POOL):
                fn(i)
        for _ in range(3):
            g.replay()
        return time_loop(lambda i: g.replay(), reps) / (reps * POOL)

    kw = dict(n_fft=ap.n_fft, hop=ap.hop_length, win_length=ap.win_length)
    reps = max(4, min(steps // POOL, 100))
    t_k = graph_time(lambda i: ops.explain(pool[i].wav, pool[i].mask, length=N,
                                           out=(pool[i].rel, pool[i].irr, pool[i].stats), **kw), reps)
    ach = BYTES_EXPLAIN * BATCH / t_k / 1e9

    # ---- API-boundary kernels the north star judges on HBM GB/s: stft (X only / X+mag+phase) and istft
    specs = [ops.stft(p.wav, want_mag=False, want_phase=False, **kw)[0] for p in pool]
    t_stft = graph_time(lambda i: ops.stft(pool[i].wav, want_mag=False, want_phase=False, **kw), reps)
    t_stft3 = graph_time(lambda i: ops.stft(pool[i].wav, **kw), reps)
    t_istft = graph_time(lambda i: ops.istft(specs[i], length=N, **kw), reps)
    del specs
    # the same two kernels at 256 clips per launch (BASELINE configs[2]'s batch): fixed launch / ramp / tail costs
    # amortised, 4 rotating sets (0.26 GB of waveforms, 0.84 GB of spectra) > L2
    big = None
    try:
        BB = 256
        wb = [0.1 * torch.randn(BB, N, generator=gen, device="cuda") for _ in range(4)]
        sb_ = [ops.stft(w_, want_mag=False, want_phase=False, **kw)[0] for w_ in wb]

        def big_time(fn):   # the 4 sets replayed from one CUDA graph (eager Python calls are not always faster than these kernels)
            fn(0)
            torch.cuda.synchronize()
            g4 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g4):
                for i in range(4):
                    fn(i)
            for _ in range(3):
                g4.replay()
            return time_loop(lambda i: g4.replay(), 10) / 40

        tb_s = big_time(lambda i: ops.stft(wb[i], want_mag=False, want_phase=False, **kw))
        tb_3 = big_time(lambda i: ops.stft(wb[i], **kw))
        tb_i = big_time(lambda i: ops.istft(sb_[i], length=N, **kw))
        big = {"clips_per_launch": BB,
               "stft_X": {"us": tb_s * 1e6, "frac": BYTES_STFT * BB / tb_s / 1e9 / peak},
               "stft_X_mag_phase": {"us": tb_3 * 1e6, "frac": (BYTES_STFT + 8 * F * T) * BB / tb_3 / 1e9 / peak},
               "istft": {"us": tb_i * 1e6, "frac": BYTES_ISTFT * BB / tb_i / 1e9 / peak}}
        del wb, sb_
    except Exception as e:
        big = {"error": repr(e)[:200]}
    # ... and at 1 024 clips per launch (the chunk size of BASELINE configs[3]): where the kernels settle once a launch is
    # long against its fixed costs (2 rotating sets: 0.5 GB of waveforms, 1.7 GB of spectra)
    huge = None
    try:
        BH = 1024
        wh = [0.1 * torch.randn(BH, N, generator=gen, device="cuda") for _ in range(2)]
        mh = [torch.rand(BH, F, T, generator=gen, device="cuda") for _ in range(2)]
        sh = [ops.stft(w_, want_mag=False, want_phase=False, **kw)[0] for w_ in wh]
        oh = [(torch.empty(BH, N, device="cuda"), torch.empty(BH, N, device="cuda"),
               torch.empty((BH, ops.get_plan(ap.n_fft, ap.hop_length, ap.win_length, None, T, N, N).tiles(BH), 4),
                           dtype=torch.float64, device="cuda")) for _ in range(2)]

        def huge_time(fn):
            fn(0)
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                for i in range(2):
                    fn(i)
            for _ in range(2):
                g2.replay()
            return time_loop(lambda i: g2.replay(), 6) / 12

        th_s = huge_time(lambda i: ops.stft(wh[i], want_mag=False, want_phase=False, **kw))
        th_3 = huge_time(lambda i: ops.stft(wh[i], **kw))
        th_i = huge_time(lambda i: ops.istft(sh[i], length=N, **kw))
        th_e = huge_time(lambda i: ops.explain(wh[i], mh[i], length=N, out=oh[i], **kw))
        huge = {"clips_per_launch": BH,
                "stft_X": {"us": th_s * 1e6, "frac": BYTES_STFT * BH / th_s / 1e9 / peak},
                "stft_X_mag_phase": {"us": th_3 * 1e6, "frac": (BYTES_STFT + 8 * F * T) * BH / th_3 / 1e9 / peak},
                "istft": {"us": th_i * 1e6, "frac": BYTES_ISTFT * BH / th_i / 1e9 / peak},
                "explain": {"us": th_e * 1e6, "frac": BYTES_EXPLAIN * BH / th_e / 1e9 / peak, "clips_per_s": BH / th_e}}
        del wh, mh, sh, oh
    except Exception as e:
        huge = {"error": repr(e)[:200]}
    # the reference's DEFAULT geometry (AudioProcessor(): n_fft 1024, hop 322, win 644, 5 s clips) through the same
    # entry points: streaming n_fft 1024 kernels for explain / istft (transform5_kernels.cu), generation-2 STFT
    refdef = None
    try:
        kw1 = dict(n_fft=1024, hop=322, win_length=644)
        n1, F1, T1 = 80000, 513, 1 + 80000 // 322
        w1 = [0.1 * torch.randn(BATCH, n1, generator=gen, device="cuda") for _ in range(8)]
        m1 = [torch.rand(BATCH, F1, T1, generator=gen, device="cuda") for _ in range(8)]
        s1 = [ops.stft(w_, want_mag=False, want_phase=False, **kw1)[0] for w_ in w1]

        def t8(fn):   # 8 rotating sets (1.3 GB of spectra) replayed from one CUDA graph, like the kernels above: eager calls
            fn(0)     # from Python cost 25 - 35 us each, more than the faster of these kernels
            torch.cuda.synchronize()
            g8 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g8):
                for i in range(8):
                    fn(i)
            for _ in range(3):
                g8.replay()
            return time_loop(lambda i: g8.replay(), 10) / 80

        t1e = t8(lambda i: ops.explain(w1[i], m1[i], length=n1, **kw1))
        t1s = t8(lambda i: ops.stft(w1[i], **kw1))
        t1x = t8(lambda i: ops.stft(w1[i], want_mag=False, want_phase=False, **kw1))
        t1i = t8(lambda i: ops.istft(s1[i], length=n1, **kw1))
        by_e, by_s, by_i = 4 * n1 + 4 * F1 * T1 + 8 * n1, 4 * n1 + 16 * F1 * T1, 8 * F1 * T1 + 4 * n1
        refdef = {"geometry": "64 x 5 s clips, n_fft 1024 / hop 322 / win 644",
                  "explain": {"us": t1e * 1e6, "frac": by_e * BATCH / t1e / 1e9 / peak, "clips_per_s": BATCH / t1e},
                  "stft_X_mag_phase": {"us": t1s * 1e6, "frac": by_s * BATCH / t1s / 1e9 / peak},
                  "stft_X": {"us": t1x * 1e6, "frac": (4 * n1 + 8 * F1 * T1) * BATCH / t1x / 1e9 / peak},
                  "istft": {"us": t1i * 1e6, "frac": by_i * BATCH / t1i / 1e9 / peak}}
        del w1, m1, s1
    except Exception as e:
        refdef = {"error": repr(e)[:200]}
    # mel front-end of the vocoder path (hifigan.py:163-178 geometry: n_fft 1024, hop 256, hann 1024, 80 mels), 64 clips
    # per call: ONE launch (STFT -> |X| -> band-compressed bf16x3 tcgen05 filterbank -> log, spectrum stays on the SM);
    # the two-launch path (adv_stft + 3xTF32 adv_mel_project) is timed next to it
    mel_k = None
    try:
        mel_mod = import_module("xai-audio-deepfakes_b200.mel")
        mt = mel_mod.MelSpectrogram(16000, 1024, 256, 1024, 80, 0.0, 8000.0, 1.0, "slaney", "slaney", log_compress=True)
        t_mel = graph_time(lambda i: mt(pool[i].wav), reps)
        path = mt.last_path
        mt.fused = False
        t_mel2 = graph_time(lambda i: mt(pool[i].wav), reps)
        Tm_ = 1 + N // 256
        mel_bytes = 4 * N + 4 * 80 * Tm_          # wave in, mel out (the spectrum is an intermediate)
        mel_k = {"us": t_mel * 1e6, "GBps": mel_bytes * BATCH / t_mel / 1e9,
                 "frac": mel_bytes * BATCH / t_mel / 1e9 / peak, "path": path, "launches": 1 if path == "fused" else 2,
                 "two_launch_us": t_mel2 * 1e6,
                 "note": "n_fft 1024, hop 256, hann, 80 slaney mels, power 1, log; bytes = wave in + mel out; FFT-issue "
                         "bound (2 x 1024-point real FFTs per frame pair), not HBM bound"}
    except Exception as e:
        mel_k = {"error": repr(e)[:200]}

    # ---- end to end through the public API with pinned host inputs
    # A single-process run binds itself to the CPUs next to the GPU for these legs only (the pinned buffers are first-touched
    # there; boxes whose H2D peak measured 45 - 50 instead of 55 GB/s had them on the other socket) and gives the cores back
    # before the CPU baseline.  Ranks of a multi-process run are bound from the start.
    saved_aff, e2e_aff = None, affinity
    if world == 1 and os.environ.get("ADV_NO_BIND") is None:
        try:
            saved_aff = os.sched_getaffinity(0)
            e2e_aff = pkg.distributed.bind_host_to_gpu(local)
        except Exception:
            saved_aff = None
    hp = pipeline.HostFedPipeline(ap, BATCH, use_graph=True)
    host_sets = []
    for k in range(4):
        w, m, l = synth_host(99 + 10 * rank + k)
        host_sets.append((w.pin_memory(), m.pin_memory(), l.pin_memory()))
    e2e_steps = max(8, min(steps, 200))
    for i in range(4):
        hp.step_host(*host_sets[i % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_e = time_loop(lambda i: hp.step_host(*host_sets[i % 4]), e2e_steps)
    te = torch.tensor([t_e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * BATCH * e2e_steps / float(te)
    # ---- second end-to-end leg at the reference's own device boundary (LMAC_metrics.py:106,132): only waveforms (and
    # logits) cross PCIe, the mask comes from the U-Net's mask head on the device (device-resident decoder activation
    # y1 [B,32,256,400] = 0.84 GB, read every step); plus the measured pinned-H2D peak that both legs are held against
    e2e_wave, pcie = None, None
    try:
        big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        big_d.copy_(big_h, non_blocking=True)
        t_p = time_loop(lambda i: big_d.copy_(big_h, non_blocking=True), 6) / 6
        h2d_peak = big_h.numel() / t_p / 1e9
        del big_h, big_d
        y1 = torch.randn(BATCH, 32, F - 1, T - 1, generator=gen, device="cuda")
        hw_ = 0.3 * torch.randn(32, generator=gen, device="cuda")
        hb_ = torch.zeros(1, device="cuda")
        wp = pipeline.WaveFedPipeline(ap, BATCH, y1, hw_, hb_)
        wsets = [(h[0], h[2]) for h in host_sets]
        for i in range(4):
            wp.step_host(*wsets[i % 4])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_w2 = time_loop(lambda i: wp.step_host(*wsets[i % 4]), e2e_steps)
        tw2 = torch.tensor([t_w2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw2, op=dist.ReduceOp.MAX)
        e2e_wave = {"value": world * BATCH * e2e_steps / float(tw2), "unit": "clips/s", "h2d_bytes_per_step": wp.h2d_bytes,
                    "d2h_bytes_per_step": wp.d2h_bytes, "steps": e2e_steps,
                    "pcie_frac": wp.h2d_bytes * e2e_steps / float(tw2) / 1e9 / h2d_peak,
                    "note": "waveforms + logits H2D only; mask = sigmoid(1x1 conv) of a device-resident U-Net decoder "
                            "activation [B,32,256,400] (adv_mask_head, 0.84 GB read per step), outside='drop'"}
        pcie = {"h2d_peak_GBps": h2d_peak, "e2e_frac": hp.h2d_bytes * e2e_steps / float(te) / 1e9 / h2d_peak,
                "note": "one pinned 256 MiB cudaMemcpyAsync H2D, CUDA events; e2e_frac = the headline e2e leg's H2D bytes / time / this peak"}
        del y1, wp
    except Exception as e:
        e2e_wave = {"error": repr(e)[:200]}
    if saved_aff is not None:
        try:
            os.sched_setaffinity(0, saved_aff)
        except Exception:
            pass
    # ---- vocoder side path (BASELINE configs[2] geometry, smaller batch): HiFi-GAN on the tcgen05 conv kernels
    voc = None
    try:
        H = pkg.hifigan
        gen_v = H.HifiganGenerator(H.init_weights(seed=0, std=0.01))
        vb, vt = 256, 251   # BASELINE configs[2]: 256 x 4 s clips
        mel = -4 + 2 * torch.randn(vb, 80, vt, generator=gen, device="cuda")
        gen_v.decode_batch(mel)
        n_launch = gen_v.launches
        voc_reps = 10
        t_v = time_loop(lambda i: gen_v.decode_batch(mel), voc_reps) / voc_reps
        fl = H.HifiganGenerator.flops_per_clip(vt) * vb
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
        voc = {"workload": f"HiFi-GAN V1 generator, {vb} x 4 s clips (80 x {vt} mel -> 66816 samples), bf16, "
                           f"{'reflect' if H.HifiganConfig.pad_reflect else 'zero'} 'same' padding",
               "clips_per_s": vb / t_v, "tflops": fl / t_v / 1e12, "bound": "tensor", "peak": pk,
               "frac": fl / t_v / 1e12 / pk, "launches_per_batch": n_launch, "reps": voc_reps}
        del gen_v, mel
    except Exception as e:  # the vocoder is a side path: never fail the headline bench on it
        voc = {"error": repr(e)[:200]}
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on this box's host cores: bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_step, kind, note = reference_stepper()
        w, m, l = synth_host(1234)
        cpu_step(w, m, l)
        reps_c, t0 = 0, time.perf_counter()
        while reps_c < 3 or (time.perf_counter() - t0 < 12.0 and reps_c < 400):  # ~12 s of CPU work
            cpu_step(w, m, l)
            reps_c += 1
        dtc = (time.perf_counter() - t0) / reps_c
        cpu = {"value": BATCH / dtc, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{reps_c} x {BATCH} clips of the same workload, torch {torch.__version__} CPU fp32, "
                         f"os.cpu_count()={os.cpu_count()}; {note}"}

    traffic, issue = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("explain_kernel_bytes_per_launch")
        # what actually bounds the fused kernel: instruction issue (one warp-instruction per scheduler per cycle).
        # Instruction count per launch from the committed ncu capture, time measured live, clock = the sampled one.
        wi = tj.get("explain_kernel_warp_inst_per_launch")
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        if wi:
            n_sm = torch.cuda.get_device_properties(local).multi_processor_count
            peak_i = n_sm * 4 * mhz * 1e6
            issue = {"warp_inst_per_launch": wi, "achieved_ginst_s": wi / t_k / 1e9, "peak_ginst_s": peak_i / 1e9,
                     "frac": wi / t_k / peak_i, "note": "smsp__inst_executed.sum (ncu, profiles/traffic.json) / live "
                     "kernel time vs SMs x 4 schedulers x sampled SM clock"}

    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": elapsed / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "run": {"cuda_graph": True, "schedule": args.schedule, "streams": ns, "burst_us_per_step": burst_us,
                "preheat_s": round(preheat_s, 3), "preheat_steps": preheat_steps,
                "host_affinity": (f"{len(affinity)} cpus local to the GPU (NVML)" if affinity else "unbound"),
                "e2e_host_affinity": (f"{len(e2e_aff)} cpus local to the GPU (NVML)" if e2e_aff else "unbound")},
        "e2e": {"value": e2e_val, "unit": "clips/s", "h2d_bytes_per_step": hp.h2d_bytes,
                "d2h_bytes_per_step": hp.d2h_bytes, "steps": e2e_steps},
        "e2e_wave_only": e2e_wave, "pcie": pcie,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "explain4_kernel<log1p, hop 160> (fused STFT + mask / 1-mask + 2 x iSTFT; streaming warps, no CTA barrier)",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_launch": BYTES_EXPLAIN * BATCH, "us_per_launch": t_k * 1e6,
                     "issue": issue},
        "kernels": {
            "stft_X": {"GBps": BYTES_STFT * BATCH / t_stft / 1e9, "frac": BYTES_STFT * BATCH / t_stft / 1e9 / peak,
                       "us": t_stft * 1e6},
            "stft_X_mag_phase": {"GBps": (BYTES_STFT + 8 * F * T) * BATCH / t_stft3 / 1e9,
                                 "frac": (BYTES_STFT + 8 * F * T) * BATCH / t_stft3 / 1e9 / peak, "us": t_stft3 * 1e6},
            "istft": {"GBps": BYTES_ISTFT * BATCH / t_istft / 1e9, "frac": BYTES_ISTFT * BATCH / t_istft / 1e9 / peak,
                      "us": t_istft * 1e6},
            "mel_frontend": mel_k,
            "batch256": big,
            "batch1024": huge,
            "reference_default_geometry": refdef,
        },
        "vocoder": voc,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "lmac_means": metrics,
    }
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
