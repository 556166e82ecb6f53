"""Run the UNMODIFIED reference modules without their weights / datasets (TEST INFRASTRUCTURE ONLY).

Used by ``oracle/make_golden.py`` (build container: reads ``/root/reference``) and by ``bench.py --impl reference``
(GPU box: reads ``baseline/_ref``, the git-ignored verbatim copy that ``__graft_entry__.build()`` makes where the
reference exists).  Nothing here is on the product path.

  * ``accelerate`` and ``classifier_embedder`` are stubbed in ``sys.modules`` so that the reference's
    ``audioprocessor.py`` imports verbatim (its own ``compute_stft`` / ``compute_invert_stft`` then run on the CPU);
  * ``lift`` compiles named top-level FunctionDef / ClassDef nodes of a reference file unchanged (modules with
    import-time side effects: LMAC_metrics.py, classifier_embedder.py, addvisor.py);
  * ``statements_on_lines`` extracts the statements of one function that start on given lines, for the inline mask
    arithmetic (LMAC_metrics.py:136-157, loss_function.py:36-47, captum_saliency.py:136-143).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch


def lift(path, names, env):
    """Compile the named top-level FunctionDef/ClassDef nodes of ``path`` into env."""
    tree = ast.parse(open(path).read())
    picked = [n for n in tree.body
              if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert {n.name for n in picked} == set(names), (names, [n.name for n in picked])
    exec(compile(ast.Module(picked, []), path, "exec"), env)
    return env


def statements_on_lines(path, func_name, line_ranges):
    """The statements of ``func_name`` whose first line lies in one of ``line_ranges``."""
    tree = ast.parse(open(path).read())
    picked = []

    def walk(body):
        for st in body:
            if any(lo <= st.lineno <= hi for lo, hi in line_ranges) and \
                    not isinstance(st, (ast.For, ast.With, ast.If)):
                picked.append(st)
            for attr in ("body", "orelse"):
                sub = getattr(st, attr, None)
                if isinstance(sub, list):
                    walk(sub)

    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            walk(node.body)
    return compile(ast.Module(picked, []), path, "exec"), [s.lineno for s in picked]


def import_reference_audioprocessor(ref_dir):
    """(reference ``audioprocessor`` module imported verbatim from ``ref_dir``, the ``classifier_embedder`` stub)."""
    if "audioprocessor" in sys.modules and "classifier_embedder" in sys.modules and \
            getattr(sys.modules["audioprocessor"], "__file__", "").startswith(os.path.abspath(ref_dir)):
        return sys.modules["audioprocessor"], sys.modules["classifier_embedder"]
    from transformers import Wav2Vec2Model  # noqa: F401  (before the stub: transformers probes `accelerate` on import)
    acc = types.ModuleType("accelerate")

    class Accelerator:  # audioprocessor.py:15-16 only reads .device
        def __init__(self):
            self.device = torch.device("cpu")

    acc.Accelerator = Accelerator
    sys.modules["accelerate"] = acc

    ce = types.ModuleType("classifier_embedder")
    ce.wav2vec2 = torch.nn.Identity()
    ce.processor = None
    ce.classifier = types.SimpleNamespace(coef_=np.zeros((1, 1920)), intercept_=np.zeros(1))
    lift(os.path.join(ref_dir, "classifier_embedder.py"), ["zero_mean_unit_var_norm"], ce.__dict__)
    sys.modules["classifier_embedder"] = ce
    sys.modules.pop("audioprocessor", None)
    sys.path.insert(0, os.path.abspath(ref_dir))
    try:
        import audioprocessor  # the reference module, verbatim
    finally:
        sys.path.remove(os.path.abspath(ref_dir))
    return audioprocessor, ce


class ReferencePath:
    """The reference's CPU path for one batch, built from its own code in ``ref_dir``: ``AudioProcessor.compute_stft``,
    the mask arithmetic of LMAC_metrics.py:136-144,151-154 (exec'd unchanged), ``compute_invert_stft`` x 2,
    ``zero_mean_unit_var_norm`` and the lifted ``compute_*`` metric functions."""

    def __init__(self, ref_dir, **ap_kwargs):
        self.ap_mod, self.ce = import_reference_audioprocessor(ref_dir)
        self.ap = self.ap_mod.AudioProcessor(**ap_kwargs)
        lm = os.path.join(ref_dir, "LMAC_metrics.py")
        self.body, self.lines = statements_on_lines(lm, "run_addvisor_metrics", [(136, 144), (151, 154)])
        self.menv = dict(torch=torch, F=torch.nn.functional, device=torch.device("cpu"), eps=1e-10)
        lift(lm, ["compute_fidelity", "get_score_for_predicted_class", "compute_faithfulness", "compute_AD",
                  "compute_AI", "compute_AG"], self.menv)

    @torch.no_grad()
    def step(self, wav, mask, logits):
        """wave [B,n] + mask [B,F,T] + logits [3,B] -> (rel, irr normalised, six float64 sums)."""
        _, magnitude, phase = self.ap.compute_stft(wav)
        env = dict(torch=torch, audio_processor=self.ap, device=torch.device("cpu"), mask=mask, magnitude=magnitude,
                   phase=phase)
        exec(self.body, env)
        rel = self.ce.zero_mean_unit_var_norm(env["istft_waveforms"])
        irr = self.ce.zero_mean_unit_var_norm(env["istft_irr_waveform"])
        p, th, q = (torch.sigmoid(l).unsqueeze(-1) for l in logits)
        m = self.menv
        sums = torch.stack([m["compute_faithfulness"](p, q).double().sum(), m["compute_fidelity"](th, p).double().sum(),
                            m["compute_AD"](th, p).double().sum(), m["compute_AI"](th, p).double().sum(),
                            m["compute_AG"](th, p).double().sum(), torch.tensor(float(p.shape[0]), dtype=torch.float64)])
        return rel, irr, sums
