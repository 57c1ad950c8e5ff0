#!/usr/bin/env python3
"""Table build with and without the fused row-mask write (SST_FUSE_MASKS=0/1): times of the build and of the mask step,
SHA-256 of the table, equality of the row masks with the ones k_transpose_masks makes."""
import hashlib, os, pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from spectrseqtools_b200 import mass_table as MT, synthetic as S

seq = MT.SequenceInformation(max_len=40, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
ref = None
for fuse in ("0", "1", "0", "1"):
    os.environ["SST_FUSE_MASKS"] = fuse
    MT.clear_table_cache()
    dp = MT.DynamicProgrammingTable(S.alphabet_frame(None), 32, 10e-6, 1e-3, seq)
    dev = dp.device_table()
    b, t = [], []
    for _ in range(6):
        dev.rebuild()
        x, y = dev.timings()
        b.append(x); t.append(y)
    n = 200000
    masks = np.concatenate([dev.download_masks(0, n), dev.download_masks(dev.limit - n, n), dev.download_masks(dev.limit // 2, n)])
    if ref is None:
        ref = masks
    sha = hashlib.sha256(dev.download().tobytes()).hexdigest()[:16]
    print(f"fuse {fuse}: build ms min {min(b):.4f} median {sorted(b)[3]:.4f}; masks step ms min {min(t):.4f}; sum {min(b) + min(t):.4f}; table sha {sha}; masks equal {np.array_equal(masks, ref)}")
