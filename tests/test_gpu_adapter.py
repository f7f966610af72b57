"""GPU: the C++ drop-in adapter (include/gvs_adapter.hpp: the reference's FeatureExtractor / FeatureMatcher / TrackedKeys
names on top of the C ABI) gives the same bytes as the ctypes path and the oracle."""
import os
import subprocess

import numpy as np
import pytest

import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fnv(b: bytes) -> str:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_cpp_adapter_matches_oracle(tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "test_adapter.bin")
    if not os.path.exists(exe):
        pytest.skip("adapter test driver not built (python __graft_entry__.py build)")
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_pair(w, h, cfg["disparity"], 31)
    (tmp_path / "l.raw").write_bytes(L.tobytes()); (tmp_path / "r.raw").write_bytes(R.tobytes())
    out = subprocess.run([exe, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), str(w), str(h), str(cfg["n_features"]),
                          str(cfg["n_levels"])], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
    kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
    s = cpu_ref.stereo_match(oL, oR, 200.0, 0.1, h, kl, dl, kr, dr)
    g = cpu_ref.assign_grid(kl, w, h)
    t = lines[0].split()
    assert int(t[1]) == len(kl) and int(t[3]) == len(kr) and int(t[5]) == int((s["rightIdxs"] >= 0).sum())
    assert int(t[7]) == len(kl) and int(t[9]) == g["yGrids"]
    t = lines[1].split()
    assert t[1] == fnv(kl.tobytes()) and t[3] == fnv(dl.tobytes())
    assert t[5] == fnv(s["rightIdxs"].tobytes()) and t[7] == fnv(s["estimatedDepth"].tobytes())
    t = lines[2].split()
    assert int(t[1]) == len(kl) and t[2] == fnv(kl.tobytes())
    # projecting a frame's own keys onto itself with a 4-px radius must match nearly all of them
    t = lines[3].split()
    assert int(t[1]) > 0.9 * int(t[3])
    # findOutliersR through the adapter against the oracle's sequential loop on the same inputs
    t = lines[4].split()
    from test_oracle_vs_cv2 import _inv_sigma_factor
    fx, cx, cy = 200.0, w / 2.0, h / 2.0
    z = np.where(s["estimatedDepth"] > 0, s["estimatedDepth"].astype(np.float64), 5.0)
    wp = np.stack([(kl["x"].astype(np.float64) - cx) / fx * z, (kl["y"].astype(np.float64) - cy) / fx * z, z, np.ones(len(kl))], 1)
    m = np.stack([np.arange(len(kl)), s["rightIdxs"]], 1).astype(np.int32)
    swap = (m[:, 1] >= 0) & (np.arange(len(kl)) % 5 == 0)
    m[swap, 1] = (m[swap, 1] + 1) % len(kr)
    T = np.eye(4); TR = np.eye(4); TR[0, 3] = -float(np.float32(0.1))
    out = np.zeros(len(kl), np.uint8)
    arr = [s["estimatedDepth"].copy(), s["close"].astype(np.uint8).copy(), s["rightIdxs"].copy(), s["leftIdxs"].copy()]
    nIn, nSt = cpu_ref.find_outliers(0, T, TR, fx, fx, cx, cy, 7.815, float(np.float32(0.1) * np.float32(40)), _inv_sigma_factor(cfg["n_levels"]),
                                     None, wp, None, None, m, out, kl, kr, *arr)
    assert (int(t[2]), int(t[4]), int(t[6])) == (nIn, nSt, int(out.sum())) and nSt > 0
    assert t[8] == fnv(arr[2].tobytes()) and int(t[10]) == int(arr[1].sum()) and (arr[2] != s["rightIdxs"]).sum() > 0
    # PredictMPsPosition / removeOutOfFrameMPs through the adapter against the oracle's loops on the same inputs
    import math
    n = len(kl)
    wp2 = wp.copy(); wp2[::7, 0] += 1000.0
    m2 = np.stack([np.arange(n), arr[2]], 1).astype(np.int32)          # the stereo arrays as findOutliersR left them
    mL = np.full(n, -1, np.int32); mR = np.full(len(kr), -1, np.int32)
    for i in range(n):
        mL[m2[i, 0]] = i
        if m2[i, 1] >= 0:
            mR[m2[i, 1]] = i
    outl = (np.arange(n) % 11 == 0).astype(np.uint8)
    msd = np.full(n, 30, np.float32); ls = np.full(n, np.float32(math.log(np.float32(1.2))), np.float32); nlev = np.full(n, cfg["n_levels"], np.int32)
    # MapPoint state before the call: predL = the key itself, predR = the key, levels = octave, inFrame = inFrameR = true (driver set-up)
    pl = np.stack([kl["x"], kl["y"]], 1).astype(np.float32); pr_ = pl.copy()
    sl = kl["octave"].astype(np.int32).copy(); sr = sl.copy(); inF = np.ones(n, np.uint8); inFR = np.ones(n, np.uint8)
    cam = (fx, fx, cx, cy, w, h)
    cpu_ref.predict_mappoints(2, T, TR, *cam, None, wp2, msd, ls, nlev, pl, sl, inF, pr_, sr, inFR, m2, mL, mR, outl)
    t = lines[5].split()
    assert (int(t[2]), int(t[4]), int(t[6])) == (int(inF.sum()), int(inFR.sum()), int((sl + 3 * sr).sum()))
    assert t[8] == fnv(m2.tobytes()) and t[10] == fnv(mL.tobytes()) and t[12] == fnv(mR.tobytes())
    valid = (np.arange(n) % 11 != 0).astype(np.uint8)
    keep = cpu_ref.predict_mappoints(0, T, TR, *cam, valid, wp2, msd, ls, nlev, pl, sl, inF, pr_, sr, inFR)
    t = lines[6].split()
    assert int(t[2]) == int(keep.sum()) and int(t[4]) == int(keep.sum()) + int((valid == 0).sum()) and 0 < keep.sum() < n
    # matchByRadius with the reference's seven arguments: right keys of the frame against its left keys (rad 120)
    t = lines[7].split()
    mps = dict(desc=dr, predL=np.stack([kr["x"], kr["y"]], 1).astype(np.float32), scaleLevelL=kr["octave"].astype(np.int32))
    oL = np.full(len(kl), -1, np.int32); ot = np.full((len(kr), 2), -1, np.int32)
    on = cpu_ref.match_projection(cpu_ref.Extractor(**args), 2, mps, kl, dl, None, None, None, None, w, h, 120.0, oL, None, ot)
    assert int(t[1]) == on and int(t[3]) == on and int(t[5]) == 1
    assert t[7] == fnv(ot[:, 0].astype(np.int32).tobytes()) and t[9] == fnv(oL.tobytes())
    # default constructor (2000 features, 8 levels) + imagePyramid mirror + public ssc helper
    t = lines[8].split()
    oM = cpu_ref.Extractor()
    km, dm = oM.extractKeysNew(L)
    pyr = b"".join(oM.level_image(l).tobytes() for l in range(8))
    assert int(t[1]) == len(km) and int(t[3]) == 8 and t[5] == fnv(pyr)
    lw, lh = oM.level_image(1).shape[1], oM.level_image(1).shape[0]
    assert t[7] == f"{lw}x{lh}"
    t = lines[9].split()
    gx, gy = np.meshgrid(np.arange(20, w - 20, 5), np.arange(20, h - 20, 5))
    gx = gx.ravel().astype(np.int32); gy = gy.ravel().astype(np.int32)
    resp = (1 + (gx * 7 + gy * 13) % 200).astype(np.float32)
    _, kept = oM.ssc(int(oM.featurePerLevel[0]), w, h, gx, gy, resp)
    assert int(t[1]) == len(kept) and int(t[3]) == len(gx) and int(t[5]) == int(oM.featurePerLevel[0])
    kk = np.zeros(len(kept), gvs.KEYPOINT_DTYPE)
    kk["x"] = gx[kept]; kk["y"] = gy[kept]; kk["response"] = resp[kept]; kk["angle"] = -1; kk["class_id"] = -1
    assert t[6] == fnv(kk.tobytes())
