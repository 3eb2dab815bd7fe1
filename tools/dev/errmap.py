"""Dev: where does the map differ from the oracle?  argv: width height [frames]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import oracle
from ssim_b200 import api
from ssim_b200.synth import synth_pair
W = int(sys.argv[1]); H = int(sys.argv[2])
a, b = synth_pair(W, H, 3)
s, m = api.compute_ssim(a, b, want_map=True)
o, _, om = oracle.oracle_ssim(a, b, want_map=True, taps=oracle.TAPS_TABLE)
d = np.abs(m - om)
print("global", s, o, "max map err", d.max())
bad = np.argwhere(d > 1e-3)
print("bad pixels:", len(bad))
if len(bad):
    rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
    print("bad rows:", rows[:60], "...", rows[-5:])
    print("bad cols: min %d max %d count %d; cols mod 64 hist" % (cols.min(), cols.max(), len(cols)), np.bincount(cols // 64))
    r = rows[0]
    print("row", r, "bad cols", np.where(d[r] > 1e-3)[0][:40])
s2, m2 = api.compute_ssim(a, a.copy(), want_map=True)
bad2 = np.argwhere(m2 != 1.0)
print("identical: ssim", s2, "pixels != 1:", len(bad2), "rows", np.unique(bad2[:, 0])[:40] if len(bad2) else "")
