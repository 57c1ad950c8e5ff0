"""Drop-in alias: ``import spectrseqtools.mass_explanation`` etc. resolve to the B200 implementation.

Only the modules of the mass-explanation path exist here (masses, mass_table, mass_explanation, common,
fragment_classification);
the rest of the reference pipeline (prediction, skeleton building, LP, pre-processing, CLI) is not
re-implemented — it imports these names from the same places and runs unchanged on top of them.
"""
import importlib
import sys

from spectrseqtools_b200 import _frame

_frame.install_polars_shim()  # no-op when a real polars is installed

for _name in ("masses", "mass_table", "mass_explanation", "common", "fragment_classification"):
    _mod = importlib.import_module(f"spectrseqtools_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
