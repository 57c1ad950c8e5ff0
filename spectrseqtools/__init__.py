"""Drop-in alias: ``import spectrseqtools.mass_explanation`` etc. resolve to the B200 implementation.

Only the modules of the mass-explanation path exist here (masses, mass_table, mass_explanation, common,
fragment_classification); they are registered in ``sys.modules`` under the reference's names.  The rest of the
reference pipeline (prediction, skeleton building, LP, pre-processing, CLI) is not re-implemented: when a checkout of
the reference is reachable (``SPECTRSEQTOOLS_REFERENCE``, /root/reference, or this repo's git-ignored
``baseline/_ref``) its package directory is appended to ``__path__``, so ``import spectrseqtools.prediction`` loads
the reference's own, unmodified file — which then imports the five hot-path modules from here.
"""
import importlib
import os
import pathlib
import sys

from spectrseqtools_b200 import _frame

_frame.install_polars_shim()  # no-op when a real polars is installed

_HOT = ("masses", "mass_table", "mass_explanation", "common", "fragment_classification")
for _name in _HOT:
    _mod = importlib.import_module(f"spectrseqtools_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod

_here = pathlib.Path(__file__).resolve().parent
for _cand in (os.environ.get("SPECTRSEQTOOLS_REFERENCE"), "/root/reference", _here.parent / "baseline" / "_ref"):
    if _cand and (pathlib.Path(_cand) / "spectrseqtools" / "prediction.py").is_file():
        __path__.append(str(pathlib.Path(_cand) / "spectrseqtools"))  # modules that are not ours come from upstream
        break
