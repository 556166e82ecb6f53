"""CPU, world_size 2 over gloo: shard bounds + the six-sum all-reduce give the single-process
metrics (the per-rank partial sums come from the oracle here; on GPUs they come from adv_lmac_reduce)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("xai-audio-deepfakes_b200")
    from oracle import ref_path as R
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    probs = torch.sigmoid(2 * torch.randn(3, n, 1, generator=g))
    lo, hi = pkg.distributed.shard_bounds(n)
    if hi > lo:
        sums = R.lmac_sums(probs[0, lo:hi], probs[1, lo:hi], probs[2, lo:hi])
    else:
        sums = torch.zeros(6, dtype=torch.float64)
    sums = pkg.distributed.allreduce_sums(sums)
    res = pkg.LMAC_metrics.finalize(sums)
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.array([res[k] for k in pkg.LMAC_metrics.METRIC_NAMES] + [res["count"], lo, hi]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [257, 3, 1])
def test_two_rank_allreduce_matches_single_process(tmp_path, n):
    from oracle import ref_path as R
    port = 29500 + (os.getpid() + n) % 2000
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    g = torch.Generator().manual_seed(5)
    probs = torch.sigmoid(2 * torch.randn(3, n, 1, generator=g))
    want = R.lmac_means(probs[0], probs[1], probs[2]).double().numpy()
    np.testing.assert_allclose(r0[:5], want, atol=1e-5)
    np.testing.assert_array_equal(r0[:6], r1[:6])          # every rank holds the same result
    assert r0[5] == n
    assert r0[6] == 0 and r0[7] == r1[6] and r1[7] == n    # contiguous, disjoint, complete shards


def test_shard_bounds_cover(pkg):
    for n in (0, 1, 7, 64, 100000):
        for w in (1, 2, 4, 8):
            spans = [pkg.distributed.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_chunk_bounds_cover_whole_chunks(pkg):
    """configs[3] shards whole chunks: every chunk is owned by exactly one rank, also with more ranks than chunks"""
    D = pkg.distributed
    for n_chunks in (0, 1, 5, 98):
        for w in (1, 2, 4, 8):
            owned = []
            for r in range(w):
                lo, hi = D.chunk_bounds(n_chunks, r, w)
                owned += list(range(lo, hi))
            assert owned == list(range(n_chunks))
