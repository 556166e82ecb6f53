"""Steady-state evaluation pipeline: the per-batch GPU work of run_addvisor_metrics that we own,
on preallocated buffers and (optionally) replayed as one CUDA graph.

    wave [B,n] + mask [B,F,T] --explain--> rel, irr --normalise x2--> (SSL classifier, not ours)
    logits p / theta / q [B] --lmac_reduce--> six float64 sums

Launches per step: explain, then normalize (both waves) with the metric reduction riding as one extra CTA of the same
grid = 2 kernels, no allocation, no host sync.
``step_host`` is the end-to-end entry: inputs arrive in pinned host memory, are copied to the device
on a side stream (double-buffered against the previous step's compute), and the six sums come back
to pinned host memory.
"""
from __future__ import annotations

import torch

from . import ops

KERNELS_PER_STEP = 2


class ExplainPipeline:
    def __init__(self, audio_processor, batch, mode="log1p", use_graph=True, device=None, accumulate=False):
        ap = audio_processor
        self.accumulate = accumulate  # metric sums become running totals over the steps
        self.ap, self.batch, self.mode = ap, batch, mode
        self.dev = device or ops._dev()
        self.n = int(ap.audio_length * ap.sampling_rate)
        self.T, self.F = 1 + self.n // ap.hop_length, ap.n_fft // 2 + 1
        tiles = ops.explain_tiles(ap.n_fft, ap.hop_length, ap.win_length, self.n, batch, length=self.n)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.wav = torch.zeros((batch, self.n), **f32)
        self.mask = torch.zeros((batch, self.F, self.T), **f32)
        self.logits = torch.zeros((3, batch), **f32)
        self.rel = torch.empty((batch, self.n), **f32)
        self.irr = torch.empty((batch, self.n), **f32)
        self.stats = torch.empty((batch, tiles, 4), dtype=torch.float64, device=self.dev)
        self.ws = ops.LmacWorkspace(batch, self.dev)
        self.sums = self.ws.sums
        self.graph = None
        self.launches = 0
        if use_graph:
            self._capture()

    # the launches of one step, on the current stream
    def _enqueue(self):
        self.enqueue_explain()
        self.enqueue_post()

    # the same step in two parts, for callers that overlap part 2 of batch i with part 1 of batch i+1
    def enqueue_explain(self):
        ap = self.ap
        ops.explain(self.wav, self.mask, ap.n_fft, ap.hop_length, ap.win_length, length=self.n, mode=self.mode,
                    normalize=False, out=(self.rel, self.irr, self.stats))

    def enqueue_post(self):
        ops.normalize_pair_lmac_(self.rel, self.irr, self.stats, self.logits[0], self.logits[1], self.logits[2], is_logit=True,
                                 workspace=self.ws, accumulate=self.accumulate)

    def _capture(self):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # warm up outside capture (plans, function attributes)
                self._enqueue()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()
        self.sums.zero_()

    def step(self):
        """Run one batch on whatever is in self.wav / self.mask / self.logits; results in
        self.rel / self.irr (normalised) and self.sums.  Asynchronous."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        self.launches += KERNELS_PER_STEP
        return self.sums


class PipelinedPool:
    """Software-pipelined evaluation over a pool of batches, captured as ONE CUDA graph:

        explain(j)                      on one of ``explain_streams`` normal-priority streams (round robin, so the
                                        almost empty last round of batch j overlaps the first round of batch j+1)
        normalize(j) -> lmac_reduce(j)  on a HIGH-priority stream, after explain(j)

    The fused explain kernel is instruction-issue bound and its register footprint is capped (96 per thread) so
    that two 256-thread CTAs of the memory-bound normaliser fit next to a resident explain CTA: the normaliser of
    batch j streams through L2 / HBM while explain(j+1) owns the issue slots.  One replay = ``len(pool)`` steps."""

    def __init__(self, audio_processor, batch, sets, mode="log1p", explain_streams=2, accumulate=True, pdl=False):
        # programmatic dependent launch stays OFF for this schedule: early-scheduled CTAs of the next explain kernel would
        # sit in griddepcontrol.wait on the shared memory the normaliser CTAs of the high-priority stream are meant to
        # use (measured: 774 k -> 613 k clips/s); single-stream chains keep the library default (on)
        self.pdl = bool(pdl)
        self.pipes = [ExplainPipeline(audio_processor, batch, mode, use_graph=False, accumulate=accumulate)
                      for _ in range(sets)]
        dev = self.pipes[0].dev
        self.es = [torch.cuda.Stream(device=dev) for _ in range(max(1, explain_streams))]
        self.ns = torch.cuda.Stream(device=dev, priority=-1)
        self.cap = torch.cuda.Stream(device=dev)
        self.graph = None
        self.tails = {}   # k -> graph over the first k buffer sets (the remainder of a run that is not a multiple of the pool)
        self.launches = 0

    def _enqueue_all(self, root, count=None):
        """``count`` steps (default: one per buffer set); step i runs on buffer set i % len(pipes) - a set that is used
        again waits for the normaliser / metric reduction of its previous step."""
        from ._lib import lib
        prev = lib().adv_set_pdl(int(self.pdl))
        try:
            self._enqueue_all_inner(root, count)
        finally:
            lib().adv_set_pdl(prev)

    def _enqueue_all_inner(self, root, count=None):
        n = len(self.pipes)
        count = n if count is None else count
        start = torch.cuda.Event()
        start.record(root)
        for s in self.es + [self.ns]:
            s.wait_event(start)
        post_done = {}
        for i in range(count):
            p = self.pipes[i % n]
            e = self.es[i % len(self.es)]
            with torch.cuda.stream(e):
                if i >= n:
                    e.wait_event(post_done[i - n])
                p.enqueue_explain()
                done = torch.cuda.Event()
                done.record(e)
            with torch.cuda.stream(self.ns):
                self.ns.wait_event(done)
                p.enqueue_post()
                if i + n < count:
                    post_done[i] = torch.cuda.Event()
                    post_done[i].record(self.ns)
        for s in self.es + [self.ns]:
            ev = torch.cuda.Event()
            ev.record(s)
            root.wait_event(ev)

    def capture(self):
        torch.cuda.synchronize()
        with torch.cuda.stream(self.cap):   # warm up outside capture (plans, function attributes)
            self._enqueue_all(self.cap)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.cap):
            self._enqueue_all(self.cap)
        for p in self.pipes:
            p.sums.zero_()

    def burst_us_per_step(self, replays=6):
        """Per-step time of a short burst of replays (CUDA events), microseconds: 77 - 78 us on B200, and the same over
        16 000 steps (``scripts/pool_burst.py``).  ``bench.py`` prints it next to the sustained figure as a cross-check
        that nothing but the steps sits in its timed region."""
        if self.graph is None:
            self.capture()
        for _ in range(2):
            self.graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            self.graph.replay()
        b.record()
        torch.cuda.synchronize()
        for p in self.pipes:
            p.sums.zero_()
        return a.elapsed_time(b) * 1e3 / (replays * len(self.pipes))

    def replay_tail(self, k):
        """The same pipelined schedule over exactly ``k`` steps (buffer set i % pool for step i), from its own graph -
        captured on first use (call it once during warm-up: capture synchronises the device).  Used for the remainder of
        a run that is not a multiple of the pool and, for short runs, for the whole run (one graph launch)."""
        if k <= 0:
            raise ValueError("step count must be positive")
        g = self.tails.get(k)
        if g is None:
            saved = [p.sums.clone() for p in self.pipes]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.cap):
                self._enqueue_all(self.cap, k)
            self.tails[k] = g
            for p, sv in zip(self.pipes, saved):   # (capture does not execute, the sums are untouched; belt and braces)
                p.sums.copy_(sv)
        g.replay()
        self.launches += KERNELS_PER_STEP * k

    def replay(self):
        """``len(self.pipes)`` steps on the current stream (asynchronous)."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        self.launches += KERNELS_PER_STEP * len(self.pipes)


class HostFedPipeline:
    """End-to-end stepping with host-resident inputs: two ExplainPipelines alternate so that the
    H2D copy of step i+1 (copy stream) overlaps the kernels of step i (compute stream)."""

    def __init__(self, audio_processor, batch, mode="log1p", use_graph=True):
        self.pipes = [ExplainPipeline(audio_processor, batch, mode, use_graph) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream()
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.host_sums = [torch.empty(6, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.i = 0
        p = self.pipes[0]
        self.h2d_bytes = 4 * (p.wav.numel() + p.mask.numel() + p.logits.numel())
        self.d2h_bytes = 8 * 6
        for e in self.done:
            e.record()

    def step_host(self, wav_pinned, mask_pinned, logits_pinned):
        k = self.i & 1
        p = self.pipes[k]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])      # buffers of slot k are free again
            p.wav.copy_(wav_pinned, non_blocking=True)
            p.mask.copy_(mask_pinned, non_blocking=True)
            p.logits.copy_(logits_pinned, non_blocking=True)
            self.copied[k].record()
        main.wait_event(self.copied[k])
        p.step()
        self.host_sums[k].copy_(p.sums, non_blocking=True)
        self.done[k].record(main)
        self.i += 1
        return self.host_sums[k]

    @property
    def launches(self):
        return sum(p.launches for p in self.pipes)


class WaveFedPipeline:
    """End-to-end stepping at the reference's own device boundary (LMAC_metrics.py:106,132): only the WAVEFORMS (and the
    classifier logits) arrive from the host; the mask is produced on the device - here by the U-Net's mask head
    (``adv_mask_head``: 1x1 conv 32 -> 1 + sigmoid, addvisor.py:57-60,82) from a device-resident decoder activation
    ``y1`` [B, 32, F', T'] standing in for the (reference, torch) U-Net body - and covers the top-left [F', T'] corner
    of the spectrum with the reference's crop semantics (``outside="drop"``).  Two buffer sets alternate so that the
    H2D copy of step i+1 overlaps the kernels of step i."""

    def __init__(self, audio_processor, batch, y1, head_weight, head_bias, mode="log1p"):
        ap = audio_processor
        self.ap, self.batch, self.mode = ap, batch, mode
        self.dev = ops._dev()
        self.n = int(ap.audio_length * ap.sampling_rate)
        self.y1, self.hw, self.hb = y1, head_weight, head_bias
        tiles = ops.explain_tiles(ap.n_fft, ap.hop_length, ap.win_length, self.n, batch, length=self.n)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.sets = []
        for _ in range(2):
            d = dict(wav=torch.zeros((batch, self.n), **f32), logits=torch.zeros((3, batch), **f32),
                     rel=torch.empty((batch, self.n), **f32), irr=torch.empty((batch, self.n), **f32),
                     stats=torch.empty((batch, tiles, 4), dtype=torch.float64, device=self.dev),
                     ws=ops.LmacWorkspace(batch, self.dev), copied=torch.cuda.Event(), done=torch.cuda.Event(),
                     host_sums=torch.empty(6, dtype=torch.float64).pin_memory())
            d["done"].record()
            self.sets.append(d)
        self.copy_stream = torch.cuda.Stream()
        self.i = 0
        self.launches = 0
        self.h2d_bytes = 4 * (batch * self.n + 3 * batch)
        self.d2h_bytes = 8 * 6

    def step_host(self, wav_pinned, logits_pinned):
        ap, d = self.ap, self.sets[self.i & 1]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(d["done"])
            d["wav"].copy_(wav_pinned, non_blocking=True)
            d["logits"].copy_(logits_pinned, non_blocking=True)
            d["copied"].record()
        main.wait_event(d["copied"])
        mask = ops.mask_head(self.y1, self.hw, self.hb)
        ops.explain(d["wav"], mask, ap.n_fft, ap.hop_length, ap.win_length, length=self.n, mode=self.mode, normalize=True,
                    out=(d["rel"], d["irr"], d["stats"]), outside="drop")
        ops.lmac(d["logits"][0], d["logits"][1], d["logits"][2], is_logit=True, want_scores=False, workspace=d["ws"])
        d["host_sums"].copy_(d["ws"].sums, non_blocking=True)
        d["done"].record(main)
        self.launches += KERNELS_PER_STEP + 1
        self.i += 1
        return d["host_sums"]

