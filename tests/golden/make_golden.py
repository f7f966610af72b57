#!/usr/bin/env python3
"""Generate tests/golden/*.npz with oracle/ref_cv2.py, i.e. with the real OpenCV primitives (cv2) under the
reference's control flow.  Run in the build container:  python tests/golden/make_golden.py

Fixtures hold the inputs for the small cases and only seeds + SHA-256 of the inputs for the full-size ones
(the generator gtsam_vslam_b200/synth.py is deterministic).  The reference has no golden vectors of its own
(SURVEY.md section 4); these are what pins both the C++ oracle and the CUDA path.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gtsam_vslam_b200 import synth  # noqa: E402
from oracle import ref_cv2 as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def csr(grid):
    start, items = [0], []
    for row in grid:
        for cell in row:
            items.extend(cell)
            start.append(len(items))
    return np.array(start, np.int32), np.array(items, np.int32)


def stereo_case(name, cfgname, seed, store_images, ramp=False):
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    L, Rr = synth.stereo_pair(w, h, cfg["disparity"], seed, ramp=ramp)
    feL = R.FeatureExtractor(cfg["n_features"], cfg["n_levels"]); feR = R.FeatureExtractor(cfg["n_features"], cfg["n_levels"])
    fm = R.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    st = {}
    k = R.extractORBAndStereoMatch(feL, feR, fm, L, Rr, w, h, st)
    ls, li = csr(k.lkeyGrid); rs, ri = csr(k.rkeyGrid)
    d = dict(cfg=cfgname, seed=seed, ramp=ramp, sha_left=hashlib.sha256(L.tobytes()).hexdigest(),
             sha_right=hashlib.sha256(Rr.tobytes()).hexdigest(),
             kpL=k.keyPoints, kpR=k.rightKeyPoints, descL=k.Desc, descR=k.rightDesc, rightIdxs=k.rightIdxs, leftIdxs=k.leftIdxs,
             depth=k.estimatedDepth, close=k.close, gridL_start=ls, gridL_items=li, gridR_start=rs, gridR_items=ri,
             yGrids=k.yGrids, xMult=np.float32(k.xMult), yMult=np.float32(k.yMult),
             candidates=np.array(st["candidates"]), ssc_passes=np.array(st.get("ssc_passes", [])),
             level_sizes=np.array([im.shape[::-1] for im in feL.imagePyramid]),
             pyr_sha=np.array([hashlib.sha256(im.tobytes()).hexdigest() for im in feL.imagePyramid]),
             scalePyramid=feL.scalePyramid, featurePerLevel=feL.featurePerLevel, umax=np.array(feL.umax))
    if store_images:
        d["left"], d["right"] = L, Rr
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, len(k.keyPoints), len(k.rightKeyPoints), int((k.rightIdxs >= 0).sum()), st.get("ssc_passes"))
    return cfg, feL, feR, fm, k


def projection_cases(cfg, feL, feR, fm, k, name):
    """Synthetic map points = this frame's own keys, jittered, for all three matchers."""
    rng = np.random.default_rng(7)
    w, h = cfg["width"], cfg["height"]
    nl, nr = len(k.keyPoints), len(k.rightKeyPoints)
    M = min(nl, 400)
    sel = rng.permutation(nl)[:M]
    desc = k.Desc[sel].copy()
    flip = rng.integers(0, 256, (M, 3))
    for i in range(M):  # flip a few bits so that distances are not all zero
        for b in flip[i]:
            desc[i, b // 8] ^= np.uint8(1 << (b % 8))
    predL = np.stack([k.keyPoints["x"][sel], k.keyPoints["y"][sel]], 1).astype(np.float32) + rng.normal(0, 1.0, (M, 2)).astype(np.float32)
    ridx = k.rightIdxs[sel]
    predR = predL.copy()
    has = ridx >= 0
    predR[has, 0] = k.rightKeyPoints["x"][ridx[has]]
    predR[has, 1] = k.rightKeyPoints["y"][ridx[has]]
    predR += rng.normal(0, 1.0, (M, 2)).astype(np.float32)
    lvlL = k.keyPoints["octave"][sel].astype(np.int32)
    lvlR = np.clip(lvlL + rng.integers(-1, 2, M), 0, cfg["n_levels"] - 1).astype(np.int32)
    inF = (rng.random(M) > 0.05).astype(np.uint8); inR = (rng.random(M) > 0.3).astype(np.uint8)
    mps = dict(desc=desc, predL=predL, predR=predR, scaleLevelL=lvlL, scaleLevelR=lvlR, inFrame=inF, inFrameR=inR)
    out = dict(**{"mp_" + kk: v for kk, v in mps.items()})
    for rad in (4.0, 10.0, 120.0):
        mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        mt[::17] = [3, -1]  # some map points are already paired and must be skipped (src/FeatureMatcher.cpp:262)
        n = fm.matchByProjectionRPred(mps, k, mL, mR, mt, rad)
        out[f"rpred_{int(rad)}_n"] = n; out[f"rpred_{int(rad)}_mL"] = mL; out[f"rpred_{int(rad)}_mR"] = mR; out[f"rpred_{int(rad)}_mt"] = mt
        mL = np.full(nl, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        n = fm.matchByProjectionMono(mps, k, mL, mt, rad)
        out[f"mono_{int(rad)}_n"] = n; out[f"mono_{int(rad)}_mL"] = mL; out[f"mono_{int(rad)}_mt"] = mt
    # matchByRadius: "last" keys = the right image's keys (a different view of the same scene)
    last = R.TrackedKeys(); last.keyPoints, last.Desc = k.rightKeyPoints, k.rightDesc
    mL = np.full(nl, -1, np.int32)
    n, m = fm.matchByRadius(last, k, mL, 120.0)
    out["radius_n"] = n; out["radius_mL"] = mL; out["radius_match"] = m
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {kk: int(v) for kk, v in out.items() if kk.endswith("_n")})


def lba_inputs(cfg, k, seed=11):
    """Candidate points for matchByProjectionRPredLBA: this frame's keys seen from a 'last key frame' shifted by ~14 px."""
    rng = np.random.default_rng(seed)
    nl, nr = len(k.keyPoints), len(k.rightKeyPoints)
    M = min(nl, 350)
    sel = rng.permutation(nl)[:M]
    desc = k.Desc[sel].copy()
    flip = rng.integers(0, 256, (M, 2))
    for i in range(M):
        for b in flip[i]:
            desc[i, b // 8] ^= np.uint8(1 << (b % 8))
    predL = np.stack([k.keyPoints["x"][sel], k.keyPoints["y"][sel]], 1).astype(np.float32) + rng.normal(0, 0.8, (M, 2)).astype(np.float32)
    ridx = k.rightIdxs[sel]
    predR = np.full((M, 2), -1.0, np.float32)           # not visible in the right image unless stereo matched
    has = ridx >= 0
    predR[has, 0] = k.rightKeyPoints["x"][ridx[has]]; predR[has, 1] = k.rightKeyPoints["y"][ridx[has]]
    predR[has] += rng.normal(0, 0.8, (int(has.sum()), 2)).astype(np.float32)
    predL[rng.random(M) < 0.1] = -5.0                   # some points project outside the left image
    shift = np.where(rng.random((M, 1)) < 0.8, 14.0, 3.0).astype(np.float32)   # 20 % fail the 10-px parallax test
    lastL = predL + shift * np.array([[0.8, 0.6]], np.float32)
    lastR = predR + shift * np.array([[0.6, -0.8]], np.float32)
    predScale = np.clip(k.keyPoints["octave"][sel] + rng.integers(-1, 2, M), 0, cfg["n_levels"] - 1).astype(np.int32)
    unF = np.where(rng.random(nl) < 0.15, 7, -1).astype(np.int32)
    unFR = np.where(rng.random(nr) < 0.15, 9, -1).astype(np.int32)
    return dict(desc=desc, predL=predL, predR=predR, predScale=predScale, lastL=lastL, lastR=lastR), unF, unFR


def lba_case(cfg, feL, feR, fm, k, name):
    pts, unF, unFR = lba_inputs(cfg, k)
    out = dict(**{"pt_" + kk: v for kk, v in pts.items()}, unF=unF, unFR=unFR)
    for rad in (4.0, 15.0):
        n, o = fm.matchByProjectionRPredLBA(pts, k, unF, unFR, rad)
        out[f"lba_{int(rad)}_n"] = n; out[f"lba_{int(rad)}_out"] = o
    # MapPoint::calcDescriptor on random observation sets built from real descriptors
    rng = np.random.default_rng(5)
    sizes = [1, 2, 3, 4, 7, 10, 33, 64]
    descs, offs, best = [], [0], []
    for sz in sizes:
        base = k.Desc[rng.integers(0, len(k.Desc))]
        d = np.repeat(base[None], sz, 0).copy()
        for r in range(sz):
            for b in rng.integers(0, 256, rng.integers(0, 40)):
                d[r, b // 8] ^= np.uint8(1 << (b % 8))
        descs.append(d); offs.append(offs[-1] + sz); best.append(R.calcDescriptor(d))
    out["med_descs"] = np.concatenate(descs); out["med_offsets"] = np.array(offs, np.int32); out["med_best"] = np.array(best, np.int32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {kk: int(v) for kk, v in out.items() if kk.endswith("_n")}, best)


if __name__ == "__main__":
    c = stereo_case("small_s0", "small", 0, True)
    projection_cases(*c, "small_s0_proj")
    lba_case(*c, "small_s0_lba")
    stereo_case("small_s1_ramp", "small", 1, True, ramp=True)
    stereo_case("euroc_s0", "euroc", 0, False)
    stereo_case("kitti_s0", "kitti", 0, False)
