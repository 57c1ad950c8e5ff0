"""spectrseqtools_b200 — B200-native mass-explanation path of SpectrSeqTools.

Host side in Python (same names as the reference's ``spectrseqtools.masses`` / ``mass_table`` /
``mass_explanation`` modules), compute in hand-written sm_100a CUDA kernels behind a C-ABI
(``include/sst_b200.h``, ``libsst_b200.so``).  No CPU fallback.
"""
__version__ = "0.1.0"
