"""ctypes binding of libaddvisor_sm100.so (the C ABI declared in include/addvisor_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, importing
this module raises; if no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libaddvisor_sm100.so"
# ADV_LIB_PATH: load another build of the same library (A/B variants built with ADV_NVCC_EXTRA); never rebuilt from here
LIB_PATH = os.environ.get("ADV_LIB_PATH") or os.path.join(_HERE, LIB_NAME)
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["capi.cu", "transform_kernels.cu", "transform3_kernels.cu", "transform4_kernels.cu", "transform5_kernels.cu", "pointwise_kernels.cu", "gemm_kernels.cu", "mel_fused_kernels.cu", "conv_tma_kernels.cu",
           "resunit_kernels.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC"]

ADV_OK, ADV_ERR_INVALID, ADV_ERR_UNSUPPORTED, ADV_ERR_NOLA = 0, -1, -2, -3
ADV_ERR_SHORT_INPUT, ADV_ERR_CUDA, ADV_ERR_SHAPE = -4, -5, -6
MASK_LOG1P, MASK_LINEAR = 0, 1
MASK_DROP_OUTSIDE = 0x100
STFT_ZERO_PAD = 1


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if os.environ.get("ADV_LIB_PATH") and os.path.exists(LIB_PATH):
        return LIB_PATH
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(_HERE, "..", "include", "addvisor_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(d) for d in deps)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("ADV_NVCC_EXTRA", "").split()  # e.g. -DADV_NO_F32X2 for A/B builds
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    objdir = os.path.join(_HERE, "..", "build", "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):  # translation units compile in parallel (the transform kernels alone take minutes)
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run([nvcc] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        return obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    for obj, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    tmp = LIB_PATH + ".tmp"
    res = subprocess.run([nvcc, "-shared", "-o", tmp] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_SIGS = {
    # name: (restype, argtypes)
    "adv_version": (C.c_int, []),
    "adv_set_pdl": (C.c_int, [C.c_int]),
    "adv_set_conv_epilogue": (C.c_int, [C.c_int]),
    "adv_strerror": (C.c_char_p, [C.c_int]),
    "adv_last_cuda_error": (C.c_char_p, []),
    "adv_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int]),
    "adv_plan_destroy": (None, [C.c_void_p]),
    "adv_plan_bins": (C.c_int, [C.c_void_p]),
    "adv_plan_frames": (C.c_int, [C.c_void_p]),
    "adv_plan_tiles": (C.c_int, [C.c_void_p, C.c_int]),
    "adv_plan_tiles_istft": (C.c_int, [C.c_void_p, C.c_int]),
    "adv_stft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p]),
    "adv_istft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p]),
    "adv_explain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_explain_spec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_mask_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_row_stats_parts": (C.c_int, [C.c_int]),
    "adv_row_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "adv_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_void_p]),
    "adv_normalize_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "adv_normalize_pair_lmac": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_lmac_blocks": (C.c_int, [C.c_int]),
    "adv_lmac_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_td_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "adv_mask_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                C.c_void_p]),
    "adv_band_swap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]),
}
_SIGS.update({
    "adv_mel_project": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "adv_resample_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "adv_mel_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "adv_conv1d_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_void_p]),
    "adv_conv1d_bf16_tma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "adv_resunit_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "adv_stft_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_void_p]),
    "adv_plan_inv_env": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_mask_grad_linear": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "adv_band_swap_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "adv_xcorr_blocks": (C.c_int, [C.c_int, C.c_int]),
    "adv_xcorr_shift": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adv_xcorr_fd_mac": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "adv_argmax_first": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "adv_mel_to_channels_last": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p]),
    "adv_avg3_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "adv_halo_fix_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "adv_avg_relayout_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "adv_post_conv_tanh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_int, C.c_void_p, C.c_void_p]),
})
_OPTIONAL_SIGS = {}


def lib():
    """The loaded library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_NAME} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). This package has no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(handle, name)  # AttributeError = header / library out of sync
                    fn.restype, fn.argtypes = res, args
                for name, (res, args) in _OPTIONAL_SIGS.items():
                    if hasattr(handle, name):
                        fn = getattr(handle, name)
                        fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    """Map a C status to the exception the reference path would have raised."""
    if rc == ADV_OK:
        return
    L = lib()
    msg = L.adv_strerror(rc).decode()
    if rc == ADV_ERR_NOLA:  # torch.istft's own text
        raise RuntimeError("window overlap add min: 1")
    if rc == ADV_ERR_CUDA:
        raise RuntimeError(f"{what}: CUDA error: {L.adv_last_cuda_error().decode()}")
    if rc == ADV_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc in (ADV_ERR_SHAPE, ADV_ERR_SHORT_INPUT):
        raise RuntimeError(f"{what}: {msg}")
    raise ValueError(f"{what}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("addvisor_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Plan:
    """Owner of one adv_plan (geometry + device tables)."""

    def __init__(self, n_fft, hop, win_length, window, n_frames, n_in, n_out):
        import numpy as np
        require_cuda()
        self.handle = C.c_void_p()
        wptr = C.c_void_p(0)
        if window is not None:
            self._w = np.ascontiguousarray(window, dtype=np.float32)
            if self._w.shape[0] != win_length:
                raise ValueError("window length must equal win_length")
            wptr = C.c_void_p(self._w.ctypes.data)
        check(lib().adv_plan_create(C.byref(self.handle), n_fft, hop, win_length, wptr, n_frames, n_in, n_out),
              "adv_plan_create")
        self.n_fft, self.hop, self.win_length = n_fft, hop, win_length
        self.n_frames, self.n_in, self.n_out = n_frames, n_in, n_out
        self.bins = n_fft // 2 + 1

    def tiles(self, batch):
        return lib().adv_plan_tiles(self.handle, batch)

    def tiles_istft(self, batch):
        return lib().adv_plan_tiles_istft(self.handle, batch)

    def __del__(self):
        try:
            if self.handle:
                lib().adv_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


import collections

# Plans are cached per (geometry, clip length, device) in a bounded LRU: the variable-length callers
# (hifigan.align_waveforms / band_swapped_waveforms over a corpus of files) would otherwise create - and never
# free - one device table set per distinct length.  An evicted plan is destroyed when the last
# caller still holding it lets go (Plan.__del__), never while a launch that uses it is being issued.
PLAN_CACHE_SIZE = int(os.environ.get("ADV_PLAN_CACHE", "64"))
_plans = collections.OrderedDict()
_plans_lock = threading.Lock()


def get_plan(n_fft, hop, win_length, window, n_frames, n_in, n_out):
    import torch
    wkey = None
    if window is not None:
        window = window.detach().float().cpu().numpy() if hasattr(window, "detach") else window
        wkey = window.tobytes()
    key = (n_fft, hop, win_length, wkey, n_frames, n_in, n_out, torch.cuda.current_device())
    with _plans_lock:
        p = _plans.get(key)
        if p is not None:
            _plans.move_to_end(key)
            return p
    p = Plan(n_fft, hop, win_length, window, n_frames, n_in, n_out)
    with _plans_lock:
        _plans[key] = p
        while len(_plans) > max(1, PLAN_CACHE_SIZE):
            _plans.popitem(last=False)
    return p


def plan_cache_info():
    """(entries, capacity) of the plan cache."""
    return len(_plans), PLAN_CACHE_SIZE
