#!/usr/bin/env python3
"""Differential fuzzing of the CUDA path against the oracle on random constructor parameters, image sizes and image
kinds (run on a GPU box):  python tools/fuzz_parity.py [iterations] [seed] [only=it1,it2,...]
Every mismatch prints the parameter set; `only=` replays the same random stream but runs (and diagnoses, level by level)
just the listed iterations.  Exit code 1 if any mismatch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs  # noqa: E402
from gtsam_vslam_b200 import synth  # noqa: E402
from oracle import cpu_ref  # noqa: E402


def box(img, k):
    c = np.cumsum(np.cumsum(np.pad(img.astype(np.float64), ((k, 0), (k, 0))), 0), 1)
    return (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)


def make_image(rng, kind, w, h):
    if kind == "rects":
        return synth.stereo_pair(w, h, 0, int(rng.integers(1 << 30)))[0]
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "smooth":
        k = int(rng.integers(2, 6))
        b = box(rng.integers(0, 256, (h + k, w + k)).astype(np.float64), k)[:h, :w]
        b = (b - b.min()) / max(b.max() - b.min(), 1e-9)
        return (b * 255).astype(np.uint8)
    if kind == "checker":
        c = int(rng.integers(5, 40))
        yy, xx = np.mgrid[0:h, 0:w]
        lo, hi = sorted(rng.integers(0, 256, 2).tolist())
        return np.where(((yy // c) + (xx // c)) % 2 == 0, lo, hi).astype(np.uint8)
    if kind == "lowcontrast":
        base = synth.stereo_pair(w, h, 0, int(rng.integers(1 << 30)))[0].astype(np.int32)
        return (100 + (base - 128) // int(rng.integers(4, 12))).clip(0, 255).astype(np.uint8)
    if kind == "dots":
        img = np.full((h, w), int(rng.integers(20, 230)), np.uint8)
        n = int(rng.integers(5, 400))
        ys = rng.integers(0, h - 3, n); xs = rng.integers(0, w - 3, n)
        for y, x in zip(ys, xs):
            img[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4))] = int(rng.integers(0, 256))
        return img
    if kind == "saturated":
        base = synth.stereo_pair(w, h, 0, int(rng.integers(1 << 30)))[0].astype(np.int32)
        return ((base - 128) * 4 + 128).clip(0, 255).astype(np.uint8)
    raise ValueError(kind)


KINDS = ["rects", "noise", "smooth", "checker", "lowcontrast", "dots", "saturated"]


def same_keys(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def diagnose(fe, o, nL, side):
    for l in range(nL):
        a = fe.level_image(0, l); b = o.level_image(l)
        x, y, s = fe.level_candidates(0, l); c = o.level_keys(l, False)
        same_c = len(x) == len(c["x"]) and np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(s, c["response"].astype(np.int32))
        xk, yk, sk = fe.level_kept(0, l); ck = o.level_keys(l, True)
        same_k = len(xk) == len(ck["x"]) and np.array_equal(xk, ck["x"]) and np.array_equal(yk, ck["y"])
        msg = f"   {side} L{l} {a.shape[1]}x{a.shape[0]} pyr {'ok' if np.array_equal(a, b) else 'DIFF ' + str(int((a != b).sum()))} cand {len(x)}/{len(c['x'])} {'ok' if same_c else 'DIFF'} kept {len(xk)}/{len(ck['x'])} {'ok' if same_k else 'DIFF'}"
        if not same_c:
            sa = set(zip(x.tolist(), y.tolist(), s.tolist())); sb = set(zip(c["x"].tolist(), c["y"].tolist(), c["response"].astype(int).tolist()))
            msg += f" only-gpu {sorted(sa - sb)[:6]} only-oracle {sorted(sb - sa)[:6]}"
        print(msg, flush=True)


def fuzz_matchers(rng, fm, oL, k, kl, dl, kr, dr, s, w, h, nL, tag):
    """Random map points against the projection matchers (modes 0/1/2) and the LBA matcher; returns #mismatches."""
    nl, nr = len(kl), len(kr)
    if nl < 20 or nr < 20:
        return 0
    bad = 0
    M = int(rng.integers(1, 600))
    src = rng.integers(0, nl, M)
    desc = dl[src].copy()
    flip = rng.integers(0, 256, (M, 32)) < int(rng.integers(0, 40))
    desc ^= (flip * rng.integers(1, 256, (M, 32))).astype(np.uint8)
    jit = float(rng.choice([0.5, 2.0, 6.0]))
    predL = (np.stack([kl["x"][src], kl["y"][src]], 1) + rng.normal(0, jit, (M, 2))).astype(np.float32)
    predR = (predL - np.array([float(rng.integers(0, 30)), 0], np.float32)).astype(np.float32)
    if rng.random() < 0.3:
        predL[rng.random(M) < 0.1] = -5.0          # out-of-frame predictions
    lv = lambda: np.clip(kl["octave"][src] + rng.integers(-1, 2, M), 0, nL - 1).astype(np.int32)
    mps = dict(desc=desc, predL=predL, predR=predR, scaleLevelL=lv(), scaleLevelR=lv(),
               inFrame=(rng.random(M) < 0.9).astype(np.uint8), inFrameR=(rng.random(M) < 0.8).astype(np.uint8))
    rad = float(rng.choice([1.0, 3.0, 8.0, 20.0, 120.0]))
    for mode in (0, 1, 2):
        mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32); mt = np.full((M, 2), -1, np.int32)
        pre = rng.random(nl) < 0.1
        mL[pre] = 7
        if mode != 2:
            mt[rng.random(M) < 0.1] = [3, -1]
        a = [mL.copy(), mR.copy(), mt.copy()]; b = [mL.copy(), mR.copy(), mt.copy()]
        if mode == 0:
            n1 = fm.matchByProjectionRPred(mps, 0, a[0], a[1], a[2], rad)
            n2 = cpu_ref.match_projection(oL, 0, mps, kl, dl, kr, dr, s["leftIdxs"], s["rightIdxs"], w, h, rad, b[0], b[1], b[2])
        elif mode == 1:
            n1 = fm.matchByProjectionMono(mps, 0, a[0], a[2], rad)
            n2 = cpu_ref.match_projection(oL, 1, mps, kl, dl, None, None, None, None, w, h, rad, b[0], None, b[2])
        else:
            last = dict(desc=desc, predL=predL, scaleLevelL=mps["scaleLevelL"])
            n1 = fm.matchByRadius(last, 0, a[0], a[2], rad)
            n2 = cpu_ref.match_projection(oL, 2, last, kl, dl, None, None, None, None, w, h, rad, b[0], None, b[2])
        if n1 != n2 or not all(np.array_equal(x, y) for x, y in zip(a, b)):
            bad += 1
            print("MISMATCH matcher mode", mode, tag, f"M {M} rad {rad} n {n1}/{n2}", flush=True)
    pts = dict(desc=desc, predL=predL, predR=predR, predScale=mps["scaleLevelL"],
               lastL=(predL + rng.normal(0, 12, (M, 2))).astype(np.float32), lastR=(predR + rng.normal(0, 12, (M, 2))).astype(np.float32))
    unF = np.where(rng.random(nl) < 0.2, 5, -1).astype(np.int32); unFR = np.where(rng.random(nr) < 0.2, 5, -1).astype(np.int32)
    n1, o1 = fm.matchByProjectionRPredLBA(pts, kl, dl, kr, dr, s["leftIdxs"], s["rightIdxs"], unF, unFR, w, h, rad)
    n2, o2 = cpu_ref.match_lba(oL, pts, kl, dl, kr, dr, s["leftIdxs"], s["rightIdxs"], unF, unFR, w, h, rad)
    if n1 != n2 or not np.array_equal(o1, o2):
        bad += 1
        print("MISMATCH matcher LBA", tag, f"M {M} rad {rad} n {n1}/{n2}", flush=True)
    return bad


def fuzz_tracker_loops(rng, fm, feL, k, kl, kr, s, w, h, tag):
    """Random poses and map points through findOutliersR / findOutliersMono and removeOutOfFrameMPs / PredictMPsPosition (SURVEY 8(f)
    rank 3) against the oracle's sequential loops; returns #mismatches.  Runs last: findOutliersR edits the frame's stereo arrays."""
    nl, nr = len(kl), len(kr)
    if nl < 20 or nr < 20:
        return 0
    bad = 0
    M = int(rng.integers(1, 1500))
    fx = float(rng.uniform(150, 900)); cx, cy = w / 2.0 + float(rng.normal(0, 5)), h / 2.0 + float(rng.normal(0, 5))
    ang = float(rng.normal(0, 0.05))
    T = np.eye(4); T[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]; T[:3, 3] = rng.normal(0, 0.3, 3)
    Tr = np.eye(4); Tr[0, 3] = -float(rng.uniform(0.05, 0.6))
    TR = Tr @ T
    first = rng.integers(0, nl, M)
    z = np.where(s["estimatedDepth"][first] > 0, s["estimatedDepth"][first], rng.uniform(0.5, 50, M)).astype(np.float64)
    pc = np.stack([(kl["x"][first] - cx) / fx * z, (kl["y"][first] - cy) / fx * z, z, np.ones(M)], 1)
    pc[:, :3] += rng.normal(0, 1, (M, 3)) * np.where(rng.random(M) < 0.7, 0.002, 0.2)[:, None] * z[:, None]
    pc[rng.random(M) < 0.03, 2] *= -1
    wPos = (np.linalg.inv(T) @ pc.T).T.copy()
    matches = np.stack([first, np.where(rng.random(M) < 0.6, s["rightIdxs"][first], rng.integers(-1, nr, M))], 1).astype(np.int32)
    matches[rng.random(M) < 0.15, 0] = -1
    valid = (rng.random(M) < 0.95).astype(np.uint8); inF = (rng.random(M) < 0.95).astype(np.uint8); inFR = (rng.random(M) < 0.9).astype(np.uint8)
    cam = (fx, fx, cx, cy); thres = float(rng.choice([5.991, 7.815, 20.0])); closeDist = float(np.float32(-Tr[0, 3]) * np.float32(40))
    for mode in (1, 0):      # mono first: it leaves the stereo arrays alone
        a = [matches.copy(), np.full(M, 7, np.uint8)]; b = [matches.copy(), np.full(M, 7, np.uint8)]
        cur = [k.estimatedDepth.copy(), k.close.astype(np.uint8), k.rightIdxs.copy(), k.leftIdxs.copy()]
        if mode == 0:
            n1 = fm.findOutliersR(0, T, TR, *cam, thres, closeDist, valid, wPos, inF, inFR, a[0], a[1], keys=k)
        else:
            n1 = (fm.findOutliersMono(0, T, *cam, thres, valid, wPos, inF, a[0], a[1]), 0)
        n2 = cpu_ref.find_outliers(mode, T, TR, *cam, thres, closeDist, feL.InvSigmaFactor, valid, wPos, inF, inFR, b[0], b[1], kl, kr, *cur)
        same = tuple(n1) == tuple(n2) and all(np.array_equal(x, y) for x, y in zip(a, b))
        if mode == 0:
            same = same and all(np.array_equal(x, y) for x, y in zip((k.estimatedDepth, k.close, k.rightIdxs, k.leftIdxs), cur))
        if not same:
            bad += 1
            print("MISMATCH findOutliers mode", mode, tag, f"M {M} n {n1}/{n2}", flush=True)
    msd = rng.uniform(2, 80, M).astype(np.float32); ls = np.full(M, np.float32(np.log(np.float32(1.2))), np.float32); nlev = np.full(M, 8, np.int32)
    mL = np.full(nl, -1, np.int32); mR = np.full(nr, -1, np.int32)
    mt = matches.copy()
    mL[mt[mt[:, 0] >= 0, 0]] = 3; mR[mt[mt[:, 1] >= 0, 1]] = 3
    outl = (rng.random(M) < 0.2).astype(np.uint8)
    for mode in (0, 1, 2):
        def run(fn):
            st = [np.full((M, 2), -7, np.float32), np.full(M, -7, np.int32), np.full(M, 7, np.uint8), np.full((M, 2), -7, np.float32),
                  np.full(M, -7, np.int32), np.full(M, 7, np.uint8), mt.copy(), mL.copy(), mR.copy(), outl.copy()]
            two = mode != 1
            keep = fn(mode, T, TR if two else None, *cam, w, h, valid, wPos, msd, ls, nlev, st[0], st[1], st[2], st[3] if two else None,
                      st[4] if two else None, st[5] if two else None, *(st[6:] if mode == 2 else (None, None, None, None)))
            return ([] if keep is None else [keep]) + st
        ra, rb = run(fm.predictMapPoints), run(cpu_ref.predict_mappoints)
        if not all(np.array_equal(x, y) for x, y in zip(ra, rb)):
            bad += 1
            print("MISMATCH predictMapPoints mode", mode, tag, f"M {M}", flush=True)
    return bad


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    only = None
    for a in sys.argv[3:]:
        if a.startswith("only="):
            only = set(int(v) for v in a[5:].split(","))
    bad = 0
    t0 = time.time()
    for it in range(iters):
        w = int(rng.integers(90, 1400)); h = int(rng.integers(80, 900))
        nL = int(rng.integers(1, 11)); scale = float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5, 2.0]))
        edge = int(rng.choice([19, 19, 19, 22, 31]))
        while nL > 1 and min(w, h) / scale ** (nL - 1) < 2 * edge + 36:   # the coarsest level must keep one 35-px cell
            nL -= 1
        nL = max(nL, 1)
        nF = int(rng.choice([50, 200, 500, 1000, 2000, 5000]))
        maxF = int(rng.integers(5, 60)); minF = int(rng.integers(1, maxF + 1))
        kind = str(rng.choice(KINDS)); disp = int(rng.integers(0, 40))
        B = int(rng.choice([1, 1, 1, 2, 5]))
        Ls, Rs = [], []
        for f in range(B):
            L = make_image(rng, kind if f == 0 else str(rng.choice(KINDS)), w + disp, h)
            R = np.ascontiguousarray(L[:, disp:disp + w]) if disp else L.copy()
            L = np.ascontiguousarray(L[:, :w])
            if rng.random() < 0.5:
                R = np.clip(R.astype(np.int32) + rng.integers(-2, 3, R.shape), 0, 255).astype(np.uint8)
            Ls.append(L); Rs.append(R)
        tag = f"it {it}: {w}x{h} x{B} nF {nF} nL {nL} scale {scale} edge {edge} fast {maxF}/{minF} {kind} disp {disp}"
        if only is not None and it not in only:
            continue
        try:
            feL = gvs.FeatureExtractor(nF, nL, scale, edge, 31, maxF, minF, width=w, height=h, max_images=B)
            feR = gvs.FeatureExtractor(nF, nL, scale, edge, 31, maxF, minF, width=w, height=h, max_images=B)
            fm = gvs.FeatureMatcher(300.0, 0.2, feL, feR, h)
            fm.track_stereo_batch(np.stack(Ls), np.stack(Rs))
            ks = [gvs.tracked_keys(fm, f, w, h) for f in range(B)]
        except gvs.GvsError as e:
            if "overflow" in str(e).lower() or "taller than 8x" in str(e):   # documented limits, loud
                print("LIMIT", tag, e, flush=True)
                continue
            print("ERROR", tag, e, flush=True); bad += 1
            continue
        ok2 = True
        for f in range(B):
            L, R, k = Ls[f], Rs[f], ks[f]
            oL, oR = cpu_ref.Extractor(nF, nL, scale, edge, 31, maxF, minF), cpu_ref.Extractor(nF, nL, scale, edge, 31, maxF, minF)
            kl, dl = oL.extractKeysNew(L); kr, dr = oR.extractKeysNew(R)
            s = cpu_ref.stereo_match(oL, oR, 300.0, 0.2, h, kl, dl, kr, dr)
            ok = same_keys(k.keyPoints, kl) and same_keys(k.rightKeyPoints, kr) and np.array_equal(k.Desc, dl) and np.array_equal(k.rightDesc, dr)
            okf = ok and all(np.array_equal(getattr(k, a), s[a]) for a in ("rightIdxs", "leftIdxs", "estimatedDepth", "close"))
            if okf:
                gl = cpu_ref.assign_grid(kl, w, h); gr = cpu_ref.assign_grid(kr, w, h)
                okf = all(np.array_equal(k.lkeyGrid[a], gl[a]) and np.array_equal(k.rkeyGrid[a], gr[a]) for a in ("start", "items"))
                ok = okf or ok
            if not okf:
                ok2 = False
                bad += 1
                print("MISMATCH", "extract" if not ok else "stereo/grid", f"frame {f}", tag, f"gpu {len(k.keyPoints)}/{len(k.rightKeyPoints)} oracle {len(kl)}/{len(kr)}", flush=True)
                if only is not None and B == 1:
                    diagnose(feL, oL, nL, "L"); diagnose(feR, oR, nL, "R")
            elif f == 0:
                bad += fuzz_matchers(rng, fm, oL, k, kl, dl, kr, dr, s, w, h, nL, tag)
                bad += fuzz_tracker_loops(rng, fm, feL, k, kl, kr, s, w, h, tag)
        if ok2 and it % 10 == 0:
            print("ok", tag, f"keys {len(kl)}/{len(kr)} matches {(s['rightIdxs'] >= 0).sum()}", flush=True)
    print(f"{iters} cases, {bad} bad, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
