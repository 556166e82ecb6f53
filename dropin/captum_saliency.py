"""Drop-in shim: with this directory on ``sys.path``, the reference's own import line
``from captum_saliency import ...`` (e.g. /root/reference/LMAC_metrics.py:4-6, loss_function.py:11-12,
captum_saliency.py:1-2) resolves to the B200 implementation.  The shim replaces itself in
``sys.modules`` with the package module, so ``captum_saliency`` and ``xai-audio-deepfakes_b200.captum_saliency``
are ONE module object (registered classifier, plan cache and module globals are shared)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(1, _root)
sys.modules[__name__] = importlib.import_module("xai-audio-deepfakes_b200.captum_saliency")
