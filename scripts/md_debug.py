import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import test_gpu_peer as T
kld = len(sys.argv) > 1 and sys.argv[1] == "kld"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx, g, cloud = T._multi_device_tracker(kld, [0] * n)
ctx1, g1, cloud1 = T._tracker(kld)
for f in range(T.FRAMES):
    g.compute(); g1.compute()
    ranks = [g] + [g.follower(r) for r in range(1, n)]
    ref = (g1.getParticles(), g1.rawWeights(), g1.aabb(), g1.croppedCount() if hasattr(g1, "croppedCount") else None)
    for r, t in enumerate(ranks):
        p, w, a = t.getParticles(), t.rawWeights(), t.aabb()
        bad_w = np.flatnonzero(w.view(np.uint32) != ref[1].view(np.uint32)) if len(w) == len(ref[1]) else "len %d vs %d" % (len(w), len(ref[1]))
        print("frame", f, "rank", r, "n", len(p), "particles equal", np.array_equal(p.view(np.uint32), ref[0].view(np.uint32)) if len(p) == len(ref[0]) else False,
              "aabb equal", np.array_equal(a, ref[2]), "raw mismatches at", bad_w[:12] if not isinstance(bad_w, str) else bad_w, flush=True)
