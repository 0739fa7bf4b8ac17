"""Writes the inputs of tests/golden/oracle_small_case.npz as raw binaries for make_pcl_golden.cpp:
    python tests/golden/export_golden_inputs.py DIR  ->  DIR/scene.bin, model.bin (16-byte points), particles.bin (32-byte particles)"""
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
out = sys.argv[1]
os.makedirs(out, exist_ok=True)
d = np.load(os.path.join(here, "oracle_small_case.npz"))
for name in ("scene", "model", "particles"):
    np.ascontiguousarray(d[name]).tofile(os.path.join(out, name + ".bin"))
print("written", out)
