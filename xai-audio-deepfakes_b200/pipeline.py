"""Steady-state evaluation pipeline: the per-batch GPU work of run_addvisor_metrics that we own,
on preallocated buffers and (optionally) replayed as one CUDA graph.

    wave [B,n] + mask [B,F,T] --explain--> rel, irr --normalise x2--> (SSL classifier, not ours)
    logits p / theta / q [B] --lmac_reduce--> six float64 sums

Launches per step: explain, normalize (both waves), lmac = 3 kernels, no allocation, no host sync.
``step_host`` is the end-to-end entry: inputs arrive in pinned host memory, are copied to the device
on a side stream (double-buffered against the previous step's compute), and the six sums come back
to pinned host memory.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

KERNELS_PER_STEP = 3


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

    # the three launches of one step, on the current stream
    def _enqueue(self):
        ap = self.ap
        ops.explain(self.wav, self.mask, ap.n_fft, ap.hop_length, ap.win_length, length=self.n, mode=self.mode,
                    normalize=True, out=(self.rel, self.irr, self.stats))
        ops.lmac(self.logits[0], self.logits[1], self.logits[2], is_logit=True, want_scores=False,
                 workspace=self.ws, accumulate=self.accumulate)

    # the same step in two parts, for callers that overlap part 2 of batch i with part 1 of batch i+1
    def enqueue_explain(self):
        ap = self.ap
        ops.explain(self.wav, self.mask, ap.n_fft, ap.hop_length, ap.win_length, length=self.n, mode=self.mode,
                    normalize=False, out=(self.rel, self.irr, self.stats))

    def enqueue_post(self):
        ops.normalize_pair_(self.rel, self.irr, self.stats)
        ops.lmac(self.logits[0], self.logits[1], self.logits[2], is_logit=True, want_scores=False,
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

    def __init__(self, audio_processor, batch, sets, mode="log1p", explain_streams=2, accumulate=True):
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
        start = torch.cuda.Event()
        start.record(root)
        for s in self.es + [self.ns]:
            s.wait_event(start)
        for j, p in enumerate(self.pipes[:count]):
            e = self.es[j % len(self.es)]
            with torch.cuda.stream(e):
                p.enqueue_explain()
                done = torch.cuda.Event()
                done.record(e)
            with torch.cuda.stream(self.ns):
                self.ns.wait_event(done)
                p.enqueue_post()
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
        """The same pipelined schedule over the first ``k`` buffer sets only (``k`` steps), from its own graph - captured
        on first use (call it once during warm-up: capture synchronises the device)."""
        if not 0 < k < len(self.pipes):
            raise ValueError("tail length must be in (0, pool size)")
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
