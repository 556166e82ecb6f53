"""BASELINE configs[3]: the batch-sharded evaluation gives the same sums for every world size,
and those sums are the oracle's.  World sizes are simulated in one process through the rank / world overrides of
``distributed.evaluate_synthetic_sharded`` (the gloo test covers the all-reduce itself on CPU)."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu

CFG = dict(sampling_rate=16000, n_fft=512, hop_length=160, win_length=512, audio_length=1)


def test_world_size_invariance_and_oracle(pkg, built_lib):
    ap = pkg.audioprocessor.AudioProcessor(**CFG)
    D = pkg.distributed
    n_clips, chunk = 150, 32  # 5 chunks, the last one ragged (22 clips)
    one = D.evaluate_synthetic_sharded(ap, n_clips, chunk, reduce=False).cpu().numpy()
    assert one[5] == n_clips
    for world in (2, 4, 8):  # 8 ranks > 5 chunks: some shards are empty
        parts = [D.evaluate_synthetic_sharded(ap, n_clips, chunk, rank_=r, world_=world, reduce=False).cpu().numpy()
                 for r in range(world)]
        np.testing.assert_allclose(np.sum(parts, axis=0), one, rtol=1e-12, atol=1e-9)
        assert sum(p[5] for p in parts) == n_clips

    # oracle on the same data (regenerated on the device from the same seeds, then moved to the host)
    n, F, T = 16000, 257, 101
    want = np.zeros(10)
    gen = torch.Generator(device="cuda")
    for c in range((n_clips + chunk - 1) // chunk):
        size = min(chunk, n_clips - c * chunk)
        gen.manual_seed(1234 + c)
        wav = (0.1 * torch.randn(size, n, generator=gen, device="cuda")).cpu()
        mask = torch.rand(size, F, T, generator=gen, device="cuda").cpu()
        logits = (2.0 * torch.randn(3, size, generator=gen, device="cuda")).cpu()
        rel, irr = R.explain(wav, mask, normalize=False, **CFG)
        pr = torch.sigmoid(logits).unsqueeze(-1)
        want[:6] += R.lmac_sums(pr[0], pr[1], pr[2]).numpy()
        want[6:] += [rel.double().sum(), (rel.double() ** 2).sum(), irr.double().sum(), (irr.double() ** 2).sum()]
    np.testing.assert_allclose(one[:6], want[:6], rtol=1e-5, atol=1e-3)          # metric sums (<= 1e-3 absolute)
    np.testing.assert_allclose(one[[7, 9]], want[[7, 9]], rtol=1e-4)             # energies of both outputs
    scale = np.sqrt(want[[7, 9]] * n_clips * n)                                  # |sum| <= sqrt(N * energy)
    assert np.all(np.abs(one[[6, 8]] - want[[6, 8]]) <= 1e-4 * scale)
