"""addvisor-b200: B200-native (sm_100a) explanation-evaluation hot path of ADDvisor.

Reference-named modules: ``audioprocessor`` (AudioProcessor), ``addvisor`` (mask head / UNet),
``LMAC_metrics`` (metric functions + run_addvisor_metrics), ``classifier_embedder`` (TorchLogReg,
zero_mean_unit_var_norm), ``hifigan`` (mel front-end + vocoder), ``captum_saliency`` (time-domain
saliency masks).  ``ops`` is the tensor-level wrapper over the C ABI in include/addvisor_b200.h.

The directory name contains a hyphen; import it with
``importlib.import_module("xai-audio-deepfakes_b200")`` or through the root-level ``adv_b200`` alias.
Reference scripts keep their own import lines (``from audioprocessor import AudioProcessor``) by
putting the repo's ``dropin/`` directory on ``sys.path``.
"""
from . import _lib  # noqa: F401  (does not load the .so until first use)

__version__ = "0.1.0"


def __getattr__(name):
    import importlib
    if name in ("ops", "audioprocessor", "addvisor", "LMAC_metrics", "classifier_embedder", "distributed",
                "mel", "hifigan", "captum_saliency", "pipeline", "loss_function"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
