"""CPU tests of the oracle: against brute force, and against the golden fixtures produced by the reference's own CUDA
sources on a B200 (tests/golden/*.npz, generator oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("shape", [(8, 12, 10), (16, 9, 20), (5, 33, 7), (1, 16, 16), (3, 1, 40)])
@pytest.mark.parametrize("density", [0.0, 0.003, 0.05, 0.5, 1.0])
def test_batch_edt_equals_bruteforce(oracle, shape, density):
    Z, Y, X = shape
    rng = np.random.RandomState(X * 31 + Y * 7 + Z + int(density * 1000))
    t = np.where(rng.rand(Z, Y, X) < density, 2, rng.randint(0, 2, (Z, Y, X))).astype(np.int8)
    om = oracle.OracleMapper(dict(local_size=(X, Y, Z), voxel_width=0.2, cutoff_grids_sq=100))
    om.set_glb_type(t)
    om.batch_edt()
    d, c = oracle.batch_edt_bruteforce(t)
    assert np.array_equal(om.aux, d)
    assert np.array_equal(om.coc_aux, c)
    om.close()


def test_batch_edt_tie_rules(oracle):
    """y ties -> larger y; x ties -> smaller x; z ties -> smaller z (reference local_edt_core.h:65-81,95-98,148-151)."""
    t = np.ones((5, 5, 5), np.int8)
    om = oracle.OracleMapper(dict(local_size=(5, 5, 5), voxel_width=0.2, cutoff_grids_sq=100))
    for axis, expect in [("y", 4), ("x", 0), ("z", 0)]:
        t[:] = 1
        if axis == "y":
            t[2, 0, 2] = t[2, 4, 2] = 2
        elif axis == "x":
            t[2, 2, 0] = t[2, 2, 4] = 2
        else:
            t[0, 2, 2] = t[4, 2, 2] = 2
        om.set_glb_type(t)
        om.batch_edt()
        c = int(om.coc_aux[2, 2, 2])
        got = {"x": c & 0x7ff, "y": (c >> 11) & 0x7ff, "z": (c >> 22) & 0x3ff}[axis]
        assert got == expect and om.aux[2, 2, 2] == 4
    om.close()


def test_cuda_atan2f_transliteration(oracle):
    """Sanity of the libdevice atan2f restatement: within 2 ulp of the double result over all quadrants."""
    rng = np.random.RandomState(3)
    l = oracle.lib()
    for _ in range(2000):
        y, x = rng.uniform(-50, 50, 2).astype(np.float32)
        got = l.gor_cuda_atan2f(float(y), float(x))
        ref = np.arctan2(np.float64(y), np.float64(x))
        assert abs(got - ref) <= 4 * np.spacing(np.float32(abs(ref))) + 1e-12
    assert l.gor_cuda_atan2f(0.0, -1.0) == np.float32(np.pi)
    assert l.gor_cuda_atan2f(-0.0, 1.0) == 0.0


def _cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_vs_reference_golden(gie, oracle, path):
    """Frames before any wavefront activity must equal the reference bit-for-bit (occupancy, batch EDT after the
    limited-observation pass, committed (dist, coc)).  Once the wavefronts run, equal-distance offers are resolved by
    arrival order in the reference (id_atomicMin, wave_core.cuh:9-22) and by the smaller coc id here, which can move a
    handful of distances by one step of the propagation: >= 99.9 % of the known voxels must carry the identical distance
    (tests/test_wave_pinning_cpu.py holds the full accounting against ground truth)."""
    g = np.load(path)
    cfg = gie.scenes.small_config(str(g["cfg_name"]), tuple(int(v) for v in g["size"]), cutoff_grids_sq=int(g["cutoff"]))
    frames = gie.scenes.make_frames(cfg, int(g["nframes"]), dynamic=bool(g["dynamic"]))
    om = oracle.OracleMapper(cfg)
    waves_seen = False
    exact_frames = 0
    for k, f in enumerate(frames):
        om.integrate(f)
        rt = g[f"f{k}_glb_type"]
        # occupancy part of glb_type (FNT marks come later in the frame): compare OCC / not-OCC and known / unknown
        om.update_edt()
        st = om.stats()
        waves_now = (st["fA"] + st["fB"] + st["fC"]) > 0
        known = (rt != 0) & (om.glb_type != 0)
        od = (om.pair >> np.uint64(32)).astype(np.int64)
        oid = (om.pair & np.uint64(0xffffffff)).astype(np.int64)
        rd = g[f"f{k}_pair_dist"].astype(np.int64)
        rid = g[f"f{k}_pair_id"].astype(np.int64) & 0xffffffff
        assert np.array_equal(om.glb_type == 2, rt == 2) or waves_seen, f"frame {k}: OCCUPIED set differs"
        assert ((om.glb_type != 0) != (rt != 0)).sum() == 0, f"frame {k}: known/unknown set differs"
        if not waves_seen and not waves_now:
            assert np.array_equal(om.glb_type, rt), f"frame {k}: glb_type"
            assert np.array_equal(om.aux[known], g[f"f{k}_aux"][known]), f"frame {k}: batch dist_sq"
            assert np.array_equal(od[known], rd[known]) and np.array_equal(oid[known], rid[known]), f"frame {k}: pair"
            exact_frames += 1
        else:
            waves_seen = True
            assert (om.glb_type != rt).sum() <= 2, f"frame {k}: glb_type differs in {(om.glb_type != rt).sum()} voxels"
            same = (od[known] == rd[known]).mean()
            assert same >= 0.999, f"frame {k}: only {same:.5f} of known voxels agree"
    om.close()
    assert exact_frames >= 1


def test_golden_box_hash_voxels(gie, oracle):
    """The reference's hash voxels in a halo box around the volume (occupancy value, type, dist, coc) against the oracle's
    block dump, on the static point-cloud case before wavefront activity."""
    g = np.load(os.path.join(GOLDEN, "pc_static.npz"))
    cfg = gie.scenes.small_config(str(g["cfg_name"]), tuple(int(v) for v in g["size"]), cutoff_grids_sq=int(g["cutoff"]))
    frames = gie.scenes.make_frames(cfg, int(g["nframes"]), dynamic=bool(g["dynamic"]))
    om = oracle.OracleMapper(cfg)
    H = int(g["halo"])
    for k in range(3):
        om.publishMap(frames[k])
        box = g[f"f{k}_box"]
        keys, vox = om.export_blocks()
        lut = {tuple(kk): i for i, kk in enumerate(keys.tolist())}
        pvt = om.pivots()[0]
        Zb, Yb, Xb = box.shape
        rng = np.random.RandomState(k)
        for _ in range(4000):
            z, y, x = rng.randint(0, Zb), rng.randint(0, Yb), rng.randint(0, Xb)
            gcoord = (x - H + pvt[0], y - H + pvt[1], z - H + pvt[2])
            b = lut.get((gcoord[0] >> 3, gcoord[1] >> 3, gcoord[2] >> 3))
            r = box[z, y, x]
            if b is None:
                assert r["type"] == 0
                continue
            v = vox[b][(gcoord[0] & 7) * 64 + (gcoord[1] & 7) * 8 + (gcoord[2] & 7)]
            if not r["alloc"]:
                assert v["vox_type"] == 0
                continue
            assert v["vox_type"] == r["type"] and v["occ_val"] == r["occ"], (k, gcoord)
            assert v["dist_sq"] == r["dist"] and tuple(v["coc_glb"]) == tuple(r["coc"]), (k, gcoord)
    om.close()


def test_vlp16_binning_restatement(gie, oracle):
    """convertPyntCld restated: empty bins are INFINITY, the last point of a bin wins, ranges are horizontal."""
    pts = np.zeros(4, dtype=gie.scenes.VLP16_POINT_DTYPE)
    pts["x"], pts["y"], pts["ring"] = [1.0, 2.0, -3.0, 0.0], [0.0, 0.0, 0.0, 5.0], [3, 3, 0, 15]
    img = oracle.vlp16_bin(pts.view(np.uint8), 22, 0, 4, 16, 440, 16, float(np.float32(2 * np.pi / 440)))
    assert np.isinf(img).sum() == 440 * 16 - 2
    assert img[3, 220] == 2.0            # atan2(0, +x) = 0 -> bin (0 + pi) / res = 220; the later point (range 2) wins
    assert np.isinf(img[0]).all()        # atan2(0, -x) = pi -> bin 440 == scan_num: dropped, as in the reference (:137)
    assert img[15, 330] == 5.0
