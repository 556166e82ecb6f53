"""Drop-in for the metric functions and evaluation driver of the reference's ``LMAC_metrics.py``.

``compute_fidelity / compute_faithfulness / compute_AD / compute_AI / compute_AG`` keep the
reference's names, argument order and output shapes (LMAC_metrics.py:31-73); each is a column of
the one fused sm_100a kernel ``adv_lmac_reduce``.  ``run_addvisor_metrics`` is the reference's
evaluation loop (:117-172) re-plumbed so that, per batch, the GPU work we own is three launches
(fused explain, two normalisers) plus one metric reduction at the end, and - when
``torch.distributed`` is initialised - one all-reduce of six float64 partial sums.
"""
from __future__ import annotations

import torch

from . import classifier_embedder as _ce
from . import distributed as dist_utils
from . import ops

eps = 1e-10  # LMAC_metrics.py:28 (lives inside the kernel as 1e-10f)


def _col(p, theta, q, col):
    scores, _ = ops.lmac(p, theta, q)
    return scores[:, col]


@torch.no_grad()
def compute_fidelity(theta_out, predictions, threshold=0.5):
    """LMAC_metrics.py:31-38 -> float [N,1]: 1 where masked-in and original labels agree.  ``threshold`` may be a
    float or a tensor (the reference's default is ``torch.Tensor([0.5])``).  The fused metric kernel bakes in 0.5;
    any other threshold is the same three comparisons spelled with device ops (no host round trip)."""
    if torch.is_tensor(threshold) and threshold.numel() != 1:
        thr = threshold.to(ops._dev(), torch.float32)
    else:
        thr = float(threshold.reshape(-1)[0]) if torch.is_tensor(threshold) else float(threshold)
        if thr == 0.5:
            return _col(predictions, theta_out, predictions, 1).reshape(-1, 1)
    dev = ops._dev()
    p, th = predictions.to(dev, torch.float32), theta_out.to(dev, torch.float32)
    return ((p > thr).long() == (th > thr).long()).float().reshape(-1, 1)


def get_score_for_predicted_class(p):
    """LMAC_metrics.py:43-45: p where p > 0.5 else 1 - p (same shape as ``p``)."""
    scores, _ = ops.lmac(p, p, p)
    return scores[:, 5].reshape(p.shape)


@torch.no_grad()
def compute_faithfulness(predictions, predictions_masked):
    """LMAC_metrics.py:48-52 -> [N]."""
    return _col(predictions, predictions, predictions_masked, 0)


@torch.no_grad()
def compute_AD(theta_out, predictions):
    """LMAC_metrics.py:55-59 -> [N]."""
    return _col(predictions, theta_out, predictions, 2)


@torch.no_grad()
def compute_AI(theta_out, predictions):
    """LMAC_metrics.py:62-66 -> [N]."""
    return _col(predictions, theta_out, predictions, 3)


@torch.no_grad()
def compute_AG(theta_out, predictions):
    """LMAC_metrics.py:69-73 -> [N]."""
    return _col(predictions, theta_out, predictions, 4)


@torch.no_grad()
def lmac_sums(predictions, theta_out, masked_predictions, is_logit=False, workspace=None):
    """float64 device tensor [6]: sums of FF, fidelity, AD, AI, AG over the samples and the count."""
    _, sums = ops.lmac(predictions, theta_out, masked_predictions, is_logit=is_logit, want_scores=False,
                       workspace=workspace)
    return sums


METRIC_NAMES = ("faithfulness", "fidelity", "average drop", "average increase", "average gain")


def finalize(sums, verbose=False):
    """Six (all-reduced) sums -> dict of the five means, printed like LMAC_metrics.py:164-172."""
    s = sums.detach().to("cpu", torch.float64)
    n = float(s[5])
    out = {k: (float(s[i]) / n if n > 0 else float("nan")) for i, k in enumerate(METRIC_NAMES)}
    out["count"] = int(n)
    if verbose:
        print(f"faithfulness : {out['faithfulness']:.2f}")
        print(f"fidelity: {out['fidelity']:.2f}")
        print(f"average drop : {out['average drop']:.2f}")
        print(f"average increase: {out['average increase']:.2f}")
        print(f"average gain : {out['average gain']:.2f}")
    return out


@torch.no_grad()
def run_addvisor_metrics(dir_path1=None, dir_path2=None, batch_size=4, *, waveforms=None, model=None,
                         audio_processor=None, torch_log_reg=None, mask_input="features", mode="log1p",
                         verbose=True):
    """The LMAC_metrics.py:117-172 loop.

    The reference ignores ``dir_path1/2`` and reads a hard-wired metadata file; here the clips are
    passed in: ``waveforms`` is a [N, n] tensor or an iterable of [n] / [b, n] tensors (host or
    device).  ``model`` is the mask network (``mask = model(features)`` as at :132, or
    ``model(magnitude.unsqueeze(1))`` when ``mask_input="magnitude"``, train_addvisor.py:363);
    ``torch_log_reg`` the TorchLogReg head; ``audio_processor`` an AudioProcessor.
    Under torch.distributed each rank evaluates its contiguous shard and the six metric sums are
    all-reduced (SUM) once at the end."""
    if waveforms is None or model is None or audio_processor is None or torch_log_reg is None:
        raise ValueError("run_addvisor_metrics needs waveforms=, model=, audio_processor=, torch_log_reg=")
    ap = audio_processor
    if torch.is_tensor(waveforms):
        lo, hi = dist_utils.shard_bounds(waveforms.shape[0])
        batches = [waveforms[i:min(i + batch_size, hi)] for i in range(lo, hi, batch_size)]
    else:
        batches = list(waveforms)
        lo, hi = dist_utils.shard_bounds(len(batches))
        batches = [b if b.dim() == 2 else b.unsqueeze(0) for b in batches[lo:hi]]

    p_all, th_all, q_all = [], [], []
    for wav in batches:
        feats = ap.extract_features(wav)
        if feats.dim() == 2:
            feats = feats.unsqueeze(0)
        logit_p, _ = torch_log_reg(torch.mean(feats, dim=1))                      # :130
        if mask_input == "magnitude":
            _, magnitude, _ = ap.compute_stft(wav)
            mask = model(magnitude.unsqueeze(1))
        else:
            mask = model(feats)                                                   # :132
        rel, irr = ap.explain(wav, mask, mode=mode, normalize=True)               # :136-153 fused
        net = _ce.get_wav2vec2()  # rel / irr are already normalised: skip extract_features' normaliser
        f_rel = net(rel, output_hidden_states=True).hidden_states[9]
        f_irr = net(irr, output_hidden_states=True).hidden_states[9]
        logit_th, _ = torch_log_reg(torch.mean(f_rel, dim=1))                     # :146
        logit_q, _ = torch_log_reg(torch.mean(f_irr, dim=1))                      # :156
        p_all.append(logit_p.reshape(-1))
        th_all.append(logit_th.reshape(-1))
        q_all.append(logit_q.reshape(-1))

    if p_all:
        sums = lmac_sums(torch.cat(p_all), torch.cat(th_all), torch.cat(q_all), is_logit=True)  # :160-172
    else:
        sums = torch.zeros(6, dtype=torch.float64, device=ops._dev())
    sums = dist_utils.allreduce_sums(sums)
    return finalize(sums, verbose=verbose and dist_utils.rank() == 0)
