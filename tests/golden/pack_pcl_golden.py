"""Packs what make_pcl_golden.cpp wrote (DIR/pcl_*.npy, computed by REAL PCL 1.8.0) into tests/golden/pcl_small_case.npz:
    python tests/golden/pack_pcl_golden.py DIR
Commit the .npz: tests/test_golden.py::test_oracle_matches_pcl_golden_when_present then pins the oracle to PCL."""
import glob
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
arrays = {}
for f in sorted(glob.glob(os.path.join(sys.argv[1], "pcl_*.npy"))):
    a = np.load(f)
    if a.dtype.kind == "V" and a.dtype.itemsize == 16:
        a = a.view(POINT)
    arrays[os.path.basename(f)[4:-4]] = a
np.savez_compressed(os.path.join(here, "pcl_small_case.npz"), **arrays)
print("packed", sorted(arrays))
