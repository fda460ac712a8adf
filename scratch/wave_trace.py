import os, sys, ctypes as C, numpy as np
os.environ["GIE_WAVE_TRACE"] = "1"
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
cfg = gie.scenes.make_config("cfg4")
frames = gie.scenes.make_frames(cfg, 16)
mp = gie.Mapper(cfg)
for f in frames:
    mp.publishMap(f)
mp.hash_map.sync()
st = mp.hash_map.wave_stats()
print(st)
L = st["levelsC"]
buf = np.zeros(4096 * 10, np.uint64)
lib = mp.loc_map.lib
lib.gie_debug_wave_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
rc = lib.gie_debug_wave_trace(mp.hash_map._h, buf.ctypes.data_as(C.c_void_p), 4096)
print("rc", rc)
ext = buf[4096 * 6:].reshape(4096, 4).astype(np.int64)
t = buf[:4096 * 6].reshape(4096, 6)[:L].astype(np.int64)
d = np.diff(t[:, 1:], axis=1)
print("level n  phase1 bar1 phase2 bar2 (ns)")
for l in list(range(0, min(L, 12))) + list(range(12, L, 20)):
    print(l, t[l, 0] % 1000000000, d[l].tolist(), "| ph2: mins", ext[l, 0] - t[l, 3], "exch", ext[l, 1] - ext[l, 0], "push", t[l, 4] - ext[l, 1], "| sync2", ext[l, 2] - t[l, 4], "totals", t[l, 5] - ext[l, 2])
print("mean per column", d.mean(axis=0), "total ms", (t[L - 1, 5] - t[0, 1]) / 1e6, "sum n", t[:, 0].sum())
