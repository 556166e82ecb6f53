"""Importable alias of the package directory ``xai-audio-deepfakes_b200`` (its name has a hyphen)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
pkg = importlib.import_module("xai-audio-deepfakes_b200")
sys.modules.setdefault("adv_b200_pkg", pkg)


def __getattr__(name):
    return getattr(pkg, name)
