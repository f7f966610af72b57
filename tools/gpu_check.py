#!/usr/bin/env python3
"""Stage-by-stage parity report of the CUDA path against the C++ oracle (run on a GPU box).
Prints one line per stage and config; exits non-zero if anything differs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
from oracle import cpu_ref

bad = 0
def report(name, ok, extra=""):
    global bad
    print(f"  {'ok ' if ok else 'BAD'} {name} {extra}", flush=True)
    bad += 0 if ok else 1

def run(cfgname, seeds):
    cfg = synth.CONFIGS[cfgname]
    w, h = cfg["width"], cfg["height"]
    n = len(seeds)
    Ls, Rs = synth.stereo_batch(w, h, cfg["disparity"], seeds)
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    feL = gvs.FeatureExtractor(**args, width=w, height=h, max_images=n)
    feR = gvs.FeatureExtractor(**args, width=w, height=h, max_images=n)
    fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    t = time.time()
    fm.track_stereo_batch(Ls, Rs); fm.sync()
    print(f"[{cfgname}] {n} frames gpu {1e3*(time.time()-t):.1f} ms (first call)")
    offL = feL.offsets(n); offR = feR.offsets(n)
    for f in range(n):
        oL, oR = cpu_ref.Extractor(**args), cpu_ref.Extractor(**args)
        kl, dl = oL.extractKeysNew(Ls[f]); kr, dr = oR.extractKeysNew(Rs[f])
        print(f" frame {f}: oracle {len(kl)}/{len(kr)} keys, gpu {offL[f+1]-offL[f]}/{offR[f+1]-offR[f]}")
        for side, fe, o, img in (("L", feL, oL, Ls[f]), ("R", feR, oR, Rs[f])):
            for l in range(cfg["n_levels"]):
                a = fe.level_image(f, l); b = o.level_image(l)
                report(f"{side} pyr L{l}", np.array_equal(a, b), f"diff px {(a != b).sum()}" if a.shape == b.shape else f"shape {a.shape} {b.shape}")
                a = fe.level_image(f, l, True); b = o.level_image(l, True)
                report(f"{side} blur L{l}", np.array_equal(a, b), f"diff px {(a != b).sum()}")
                x, y, s = fe.level_candidates(f, l); c = o.level_keys(l, False)
                same = len(x) == len(c["x"]) and np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(s, c["response"].astype(np.int32))
                extra = f"n {len(x)} vs {len(c['x'])}"
                if not same and len(x) == len(c["x"]):
                    extra += f" first diff {np.nonzero((x != c['x']) | (y != c['y']) | (s != c['response']))[0][:5]}"
                elif not same:
                    sa = set(zip(x.tolist(), y.tolist(), s.tolist())); sb = set(zip(c['x'].tolist(), c['y'].tolist(), c['response'].astype(int).tolist()))
                    extra += f" only gpu {len(sa - sb)} only ref {len(sb - sa)} ex {list(sa - sb)[:3]} {list(sb - sa)[:3]}"
                report(f"{side} cand L{l}", same, extra)
                x, y, s = fe.level_kept(f, l); c = o.level_keys(l, True)
                same = len(x) == len(c["x"]) and np.array_equal(x, c["x"]) and np.array_equal(y, c["y"]) and np.array_equal(s, c["response"].astype(np.int32))
                report(f"{side} kept L{l}", same, f"n {len(x)} vs {len(c['x'])} passes {o.ssc_passes(l)}")
        gk, gd = feL.download(f, 1, int(offL[f+1]-offL[f])); rk, rd = feR.download(f, 1, int(offR[f+1]-offR[f]))
        for nm, g_, o_ in (("L", gk, kl), ("R", rk, kr)):
            if len(g_) != len(o_):
                report(f"keypoints {nm}", False, "count"); continue
            for fld in o_.dtype.names:
                d = g_[fld] != o_[fld]
                report(f"keypoints {nm}.{fld}", not d.any(), f"{d.sum()} differ" + (f" max {np.abs(g_[fld].astype(np.float64)-o_[fld]).max():.3g}" if d.any() else ""))
        for nm, g_, o_ in (("L", gd, dl), ("R", rd, dr)):
            if g_.shape == o_.shape:
                nb = int(np.unpackbits(g_ ^ o_).sum())
                report(f"desc {nm}", nb == 0, f"{nb} bits of {g_.size*8}, rows {(g_ != o_).any(axis=1).sum()}")
        if len(gk) == len(kl) and len(rk) == len(kr):
            s = cpu_ref.stereo_match(oL, oR, cfg["fx"], cfg["baseline"], h, kl, dl, kr, dr)
            ridx, lidx, dep, cl = fm.stereo_download(f, 1)
            report("stereo rightIdxs", np.array_equal(ridx, s["rightIdxs"]), f"{(ridx != s['rightIdxs']).sum()} differ, matches {(ridx>=0).sum()} vs {(s['rightIdxs']>=0).sum()}")
            report("stereo leftIdxs", np.array_equal(lidx, s["leftIdxs"]), f"{(lidx != s['leftIdxs']).sum()} differ")
            report("stereo depth", np.array_equal(dep, s["estimatedDepth"]), f"{(dep != s['estimatedDepth']).sum()} differ")
            report("stereo close", np.array_equal(cl, s["close"]))
            for side, kk in ((0, kl), (1, kr)):
                g_ = fm.grid_download(f, side); o_ = cpu_ref.assign_grid(kk, w, h)
                report(f"grid side {side}", g_["yGrids"] == o_["yGrids"] and np.array_equal(g_["start"], o_["start"]) and np.array_equal(g_["items"], o_["items"]) and g_["xMult"] == o_["xMult"] and g_["yMult"] == o_["yMult"])
    st = feL.stage_times()
    return feL, feR, fm

if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["small", "euroc", "kitti"]
    for c in cfgs:
        run(c, [0, 1] if c != "hires" else [0])
    print("TOTAL BAD", bad)
    sys.exit(1 if bad else 0)
