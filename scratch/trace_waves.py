import sys, os, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
from oracle import oracle_py as O
cfg = gie.scenes.make_config("cfg4")
frames = gie.scenes.make_frames(cfg, 17)
om = O.OracleMapper(cfg)
for k, f in enumerate(frames):
    print("FRAME", k, file=sys.stderr, flush=True)
    om.publishMap(f)
    print(k, om.stats(), flush=True)
