"""Batch-sharded data parallelism for the evaluation path.

Clips are independent, so each rank runs the whole pipeline on a contiguous shard; the only
exchange is ONE all-reduce (SUM) of six float64 partial sums {FF, fidelity, AD, AI, AG, count}
per evaluation - 48 bytes over NCCL/NVLink, enqueued on the compute stream right after the
metric kernel.  No data-path collective exists or is needed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if initialized() else 1


def shard_bounds(n: int, r: int | None = None, w: int | None = None):
    """Contiguous shard [lo, hi) of ``n`` units owned by rank ``r`` of ``w`` (ceil-sized shards)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per = (n + w - 1) // w
    lo = min(n, r * per)
    return lo, min(n, lo + per)


def bind_host_to_gpu(local_rank: int):
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``local_rank`` (same NUMA node / PCIe root),
    so that the pinned host buffers it allocates afterwards are first-touched next to the GPU that reads them.  With
    one process per GPU on an 8-GPU box this is what keeps the host-fed path (``pipeline.HostFedPipeline``) from
    funnelling every H2D copy through one socket's memory.  Returns the CPU list, or None when NVML / the affinity
    call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = local_rank
        if vis:  # NVML enumerates physical devices: map the local rank through the visibility list when it is numeric
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def allreduce_sums(sums: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of the six metric partial sums (no-op on a single process)."""
    if initialized() and world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def chunk_bounds(n_chunks: int, r: int | None = None, w: int | None = None):
    """Shard of whole chunks owned by rank ``r`` (same rule as :func:`shard_bounds`, in units of chunks)."""
    return shard_bounds(n_chunks, r, w)


@torch.no_grad()
def evaluate_synthetic_sharded(audio_processor, n_clips: int, chunk: int = 1024, seed: int = 1234, mode: str = "log1p",
                               rank_: int | None = None, world_: int | None = None, reduce: bool = True):
    """BASELINE configs[3] / SURVEY 8(d) cfg-4: ``n_clips`` synthetic clips, batch-sharded over the ranks in whole
    chunks of ``chunk`` clips.  Chunk ``c`` is generated ON THE DEVICE from ``seed + c`` (global chunk index), so every
    world size sees identical data; each rank runs explain -> normalise x2 -> LMAC sums on its chunks and the only
    exchange is the all-reduce of the partial sums.

    Returns a float64 device tensor [10]: the six LMAC sums {FF, fidelity, AD, AI, AG, count} followed by a checksum
    of the transform outputs before normalisation {sum rel, sum rel^2, sum irr, sum irr^2} (from the per-tile
    partials the fused kernel writes), so that world-size invariance covers the kernels and not only the metrics.
    The classifier is not on this path (reference torch module): its three logit vectors are synthetic, N(0, 2^2).
    ``rank_`` / ``world_`` override the process group (single-process tests of the sharding)."""
    from . import ops
    ap = audio_processor
    dev = ops._dev()
    n = int(ap.audio_length * ap.sampling_rate)
    F, T = ap.n_fft // 2 + 1, 1 + n // ap.hop_length
    n_chunks = (n_clips + chunk - 1) // chunk
    lo, hi = chunk_bounds(n_chunks, rank_, world_)
    total = torch.zeros(10, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    ws = {}
    for c in range(lo, hi):
        size = min(chunk, n_clips - c * chunk)
        gen.manual_seed(seed + c)
        wav = 0.1 * torch.randn(size, n, generator=gen, device=dev)
        mask = torch.rand(size, F, T, generator=gen, device=dev)
        logits = 2.0 * torch.randn(3, size, generator=gen, device=dev)
        tiles = ops.explain_tiles(ap.n_fft, ap.hop_length, ap.win_length, n, size, length=n)
        rel = torch.empty((size, n), dtype=torch.float32, device=dev)
        irr = torch.empty_like(rel)
        stats = torch.empty((size, tiles, 4), dtype=torch.float64, device=dev)
        ops.explain(wav, mask, ap.n_fft, ap.hop_length, ap.win_length, length=n, mode=mode, normalize=True,
                    out=(rel, irr, stats))
        if size not in ws:
            ws[size] = ops.LmacWorkspace(size, dev)
        _, sums = ops.lmac(logits[0], logits[1], logits[2], is_logit=True, want_scores=False, workspace=ws[size])
        total[:6] += sums
        total[6:] += stats.sum(dim=(0, 1))
    if reduce:
        allreduce_sums(total)
    return total
