"""Batch-EDT stage times at 512^3 as a function of what the volume holds: how many z-slices hold an obstacle (s) and how many
columns of such a slice are real.  Surfaces (walls, floors, pillars) rather than random points: the shapes a mapped scene has.
usage: python scratch/edt_regimes.py"""
import sys, json, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
X = Y = Z = 512
def volume(kind):
    t = np.ones((Z, Y, X), np.int8)            # FREE
    rng = np.random.RandomState(3)
    if kind == "floor":                         # one horizontal plane + low clutter: obstacles in 12 slices
        t[100] = 2
        for _ in range(60):
            x, y = rng.randint(0, X - 8), rng.randint(0, Y - 8)
            t[101:112, y:y + 8, x:x + 8] = 2
    elif kind.startswith("walls"):              # vertical walls over a fraction of the height: s = that fraction
        frac = float(kind.split("_")[1])
        z1 = int(Z * frac)
        for x in (40, 200, 330, 470): t[:z1, :, x] = 2
        for y in (60, 250, 420): t[:z1, y, :] = 2
    elif kind == "pillars":                     # 300 thin pillars through the whole height: every slice, few columns
        for _ in range(300):
            x, y = rng.randint(0, X - 2), rng.randint(0, Y - 2)
            t[:, y:y + 2, x:x + 2] = 2
    return t
out = {}
for kind in ("floor", "walls_0.25", "walls_0.5", "walls_1.0", "pillars"):
    t = volume(kind)
    lm = gie.LocMap(0.1, (X, Y, Z), cutoff_grids_sq=2500)
    lm.upload_glb_type(t)
    lm.profile_enable(True)
    acc = {}
    for k in range(8):
        lm.batchEDTUpdate()
        if k >= 3:
            for s, v in lm.profile_last().items():
                acc[s] = acc.get(s, 0.0) + v / 5
    occ = t == 2
    slices = occ.any(axis=(1, 2))
    cols = occ.any(axis=1)[slices].sum(axis=1).mean()
    out[kind] = {"s": round(float(slices.mean()), 3), "real_columns_per_slice": round(float(cols), 1),
                 **{k: round(v, 4) for k, v in acc.items() if k.startswith("edt")}}
    lm.close()
print(json.dumps(out))
