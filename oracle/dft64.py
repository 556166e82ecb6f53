"""First-principles float64 definitions (numpy) of the transforms on the path.

TEST INFRASTRUCTURE ONLY.  No FFT library: the DFT is a dense matrix product,
reflect padding and overlap-add are explicit index arithmetic.  Small sizes only.
These definitions pin oracle/ref_path.py (and through it the CUDA kernels)
independently of torch's FFT back-ends.

Semantics restated (torch.stft / torch.istft as called at
audioprocessor.py:102-108 and :123-129):
  * window: ``win_length`` taps (ones when None) centred in ``n_fft`` with
    left pad (n_fft - win_length)//2;
  * centre=True: reflect-pad n_fft/2 each side, frame t starts at padded index t*hop,
    T = 1 + N//hop, F = n_fft//2 + 1, no normalisation, one-sided;
  * inverse: y[p] = sum_t w[p - t*hop] * irfft(X_t)[p - t*hop] / sum_t w^2[p - t*hop],
    then drop n_fft/2 leading samples and keep ``length`` (zero-pad if short);
    irfft ignores Im of the DC and Nyquist bins.
"""
from __future__ import annotations

import numpy as np


def padded_window(n_fft, win_length, window=None):
    w = np.ones(win_length, dtype=np.float64) if window is None else np.asarray(window, np.float64)
    assert w.shape[0] == win_length
    left = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[left:left + win_length] = w
    return out


def reflect_index(i, n):
    """Index into a length-n signal for padded position i (may be <0 or >=n)."""
    if i < 0:
        return -i
    if i >= n:
        return 2 * (n - 1) - i
    return i


def stft(x, n_fft, hop, win_length, window=None):
    """x: [N] float -> X [F, T] complex128."""
    x = np.asarray(x, np.float64)
    N = x.shape[0]
    w = padded_window(n_fft, win_length, window)
    T = 1 + N // hop
    Fb = n_fft // 2 + 1
    n = np.arange(n_fft)
    k = np.arange(Fb)
    W = np.exp(-2j * np.pi * np.outer(k, n) / n_fft)  # [F, n_fft]
    X = np.empty((Fb, T), dtype=np.complex128)
    for t in range(T):
        idx = [reflect_index(t * hop - n_fft // 2 + i, N) for i in range(n_fft)]
        X[:, t] = W @ (x[idx] * w)
    return X


def irfft_def(Xf, n_fft):
    """One-sided spectrum [F] -> n_fft real samples, C2R semantics."""
    Fb = n_fft // 2 + 1
    full = np.zeros(n_fft, dtype=np.complex128)
    full[:Fb] = Xf
    full[0] = Xf[0].real
    full[n_fft // 2] = Xf[n_fft // 2].real
    full[Fb:] = np.conj(Xf[1:n_fft // 2][::-1])
    n = np.arange(n_fft)
    W = np.exp(2j * np.pi * np.outer(n, n) / n_fft)
    return (W @ full).real / n_fft


def envelope(n_fft, hop, win_length, T, window=None):
    w = padded_window(n_fft, win_length, window)
    env = np.zeros(n_fft + hop * (T - 1), dtype=np.float64)
    for t in range(T):
        env[t * hop:t * hop + n_fft] += w * w
    return env


def istft(X, n_fft, hop, win_length, length=None, window=None):
    """X: [F, T] complex -> y [length] float64."""
    X = np.asarray(X, np.complex128)
    T = X.shape[1]
    w = padded_window(n_fft, win_length, window)
    full_len = n_fft + hop * (T - 1)
    y = np.zeros(full_len, dtype=np.float64)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += irfft_def(X[:, t], n_fft) * w
    env = envelope(n_fft, hop, win_length, T, window)
    start = n_fft // 2
    end = start + length if length is not None else full_len - n_fft // 2
    ys, es = y[start:min(end, full_len)], env[start:min(end, full_len)]
    if np.abs(es).min() < 1e-11:
        raise RuntimeError("window overlap add min: 1")
    out = ys / es
    if end > full_len:
        out = np.concatenate([out, np.zeros(end - full_len)])
    return out


def normalize(x):
    x = np.asarray(x, np.float64)
    mu = x.mean(axis=-1, keepdims=True)
    sd = x.std(axis=-1, ddof=1, keepdims=True)
    return (x - mu) / (sd + 1e-7)


def mask_apply(X, mask, mode="log1p"):
    """Phase-preserving gain form of the mask arithmetic in float64:
    rel = X * g(m), irr = X * g(1-m), g(m) = expm1(m*log1p(|X|))/|X| (-> m as |X|->0);
    linear mode: g(m) = m."""
    X = np.asarray(X, np.complex128)
    a = np.abs(X)
    if mode == "linear":
        return X * mask, X * (1 - mask)
    lm = np.log1p(a)
    safe = np.where(a > 0, a, 1.0)
    g_rel = np.where(a > 0, np.expm1(mask * lm) / safe, mask)
    g_irr = np.where(a > 0, np.expm1((1 - mask) * lm) / safe, 1 - mask)
    return X * g_rel, X * g_irr


def lmac_scores(p, theta, q, eps=1e-10):
    """float64 LMAC scores from probabilities (LMAC_metrics.py:31-73)."""
    p, theta, q = (np.asarray(a, np.float64).reshape(-1) for a in (p, theta, q))
    pc = np.where(p > 0.5, p, 1 - p)
    oc = np.where(theta > 0.5, theta, 1 - theta)
    ff = (p - q) * np.sign(p - 0.5)
    fid = ((p > 0.5) == (theta > 0.5)).astype(np.float64)
    ad = np.maximum(pc - oc, 0) / (pc + eps) * 100
    ai = (oc > pc).astype(np.float64) * 100
    ag = np.maximum(oc - pc, 0) / (1 - pc + eps) * 100
    return np.stack([ff, fid, ad, ai, ag], axis=1)
