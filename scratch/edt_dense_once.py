"""One dense batch EDT (512^3, random occupancy) for ncu captures.  usage: edt_dense_once.py [density] [reps]"""
import sys, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
dens = float(sys.argv[1]) if len(sys.argv) > 1 else 0.002
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X = Y = Z = 512
rng = np.random.RandomState(5)
t = np.where(rng.rand(Z, Y, X) < dens, 2, 1).astype(np.int8)
lm = gie.LocMap(0.1, (X, Y, Z), cutoff_grids_sq=2500)
lm.upload_glb_type(t)
for _ in range(reps):
    lm.batchEDTUpdate()
print("max aux", int(lm.download(gie.ARR_AUX).max()))
lm.close()
