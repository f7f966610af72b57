"""GPU (-m gpu): one-process multi-device front end (gvs_multi) and host-thread concurrency.

On a single-GPU box the devices list names GPU 0 several times: the shards are separate handle sets driven by separate host
threads, which exercises the sharding, the host gather and the thread safety promised in include/gvs.h; with more GPUs visible
the same test spreads the shards over them."""
import threading

import numpy as np
import pytest

import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth

pytestmark = pytest.mark.gpu


def _single(cfg, L, R, motion=None):
    w, h = cfg["width"], cfg["height"]
    n = len(L)
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=n)
    feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
    fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    fm.track_stereo_batch(L, R)
    offL = feL.offsets(n); offR = feR.offsets(n)
    kl, dl = feL.download(0, n, int(offL[-1])); kr, dr = feR.download(0, n, int(offR[-1]))
    ri, li, de, cl = fm.stereo_download(0, n)
    out = dict(offL=offL, offR=offR, kpL=kl, descL=dl, kpR=kr, descR=dr, rightIdxs=ri, leftIdxs=li, estimatedDepth=de, close=cl)
    if motion is not None:
        fm.track_temporal_batch(n, motion["prev"], motion["pose_prev"], motion["to_camera"], motion["to_camera_r"], motion["fx"], motion["fy"],
                                motion["cx"], motion["cy"], w, h, 10.0)
        t = fm.temporal_download(0, n)
        out.update(mp_offsets=t["mp_offsets"], mp_key=t["mp_key"], matches=t["matches"], n_matches=t["n_matches"])
    return out


@pytest.mark.parametrize("n_shards,n_frames", [(2, 6), (3, 7), (4, 4), (2, 1)])
def test_multi_device_gather_equals_single_handle(n_shards, n_frames):
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_sequence(w, h, cfg["disparity"], n_frames, seed=70)
    ndev = gvs.device_count()
    devices = [g % ndev for g in range(n_shards)]
    mt = gvs.MultiGpuTracker(cfg["fx"], cfg["baseline"], devices, width=w, height=h, frames_per_device=(n_frames + n_shards - 1) // n_shards + 1,
                             nFeatures=cfg["n_features"], nLevels=cfg["n_levels"])
    # predecessors stay inside each device's contiguous range [g*n/G, (g+1)*n/G)
    first = [g * n_frames // n_shards for g in range(n_shards + 1)]
    prev = np.full(n_frames, -1, np.int32)
    for g in range(n_shards):
        for f in range(first[g] + 1, first[g + 1]):
            prev[f] = f - 1
    mo = synth.sequence_motion(cfg, n_frames, prev)
    got = mt.track_stereo(L, R, mo)
    want = _single(cfg, L, R, mo)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    assert got["n_matches"].sum() > 0 or (prev < 0).all()
    # a predecessor on another device is refused
    if n_frames >= 4 and n_shards >= 2:
        bad = dict(mo); bad["prev"] = prev.copy(); bad["prev"][first[1]] = first[1] - 1
        with pytest.raises(gvs.GvsError):
            mt.track_stereo(L, R, bad)
    # without temporal tracking
    got2 = mt.track_stereo(L, R)
    for k in ("offL", "kpL", "descR", "rightIdxs", "estimatedDepth"):
        assert np.array_equal(got2[k], want[k]), k


def test_handles_driven_from_concurrent_host_threads():
    """include/gvs.h: different handles may be driven from different host threads concurrently (the reference runs its two
    extractors on two std::threads, src/FeatureTracker.cpp:58-61; LocalMapper calls the LBA matcher from its own thread)."""
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_batch(w, h, cfg["disparity"], range(80, 86))
    want = _single(cfg, L, R)
    errors, results = [], {}

    def extract_worker(name, images):
        try:
            fe = gvs.FeatureExtractor(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=len(images))
            for _ in range(5):
                fe.extract_batch(images)
                off = fe.offsets(len(images))
                results[name] = (off, *fe.download(0, len(images), int(off[-1])))
        except Exception as e:  # noqa: BLE001
            errors.append((name, e))

    def lba_worker():
        try:
            args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=1)
            fe1, fe2 = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
            fm = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], fe1, fe2, h)
            rng = np.random.default_rng(1)
            a = rng.integers(0, 256, (500, 32)).astype(np.uint8); b = rng.integers(0, 256, (500, 32)).astype(np.uint8)
            ref = np.unpackbits(a ^ b, axis=1).sum(1)
            for _ in range(20):
                assert np.array_equal(fm.DescriptorDistance(a, b), ref)
                off = np.array([0, 3, 10, 500], np.int32)
                assert len(fm.median_descriptor(a, off)) == 3
        except Exception as e:  # noqa: BLE001
            errors.append(("lba", e))

    threads = [threading.Thread(target=extract_worker, args=("L", L)), threading.Thread(target=extract_worker, args=("R", R)),
               threading.Thread(target=lba_worker)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(results["L"][0], want["offL"]) and np.array_equal(results["L"][1], want["kpL"]) and np.array_equal(results["L"][2], want["descL"])
    assert np.array_equal(results["R"][0], want["offR"]) and np.array_equal(results["R"][1], want["kpR"]) and np.array_equal(results["R"][2], want["descR"])


def test_guard_tails_and_poisoned_allocations():
    """compute-sanitizer is closed on the GPU pool: every device buffer gets a guard tail (GVS_GUARD) and a poison fill
    (GVS_POISON), the whole step / pipeline / multi-device front end runs (tools/sanitize_target.py, bit-exact smoke included) and
    no tail may have been overwritten."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GVS_GUARD="1", GVS_POISON="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_target.py")], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    line = [l for l in out.stdout.splitlines() if l.startswith("guard tails:")][0]
    assert line.startswith("guard tails: 0 overwritten of ") and int(line.split()[5]) > 100, line


def test_second_matcher_on_the_same_extractors_from_another_thread(golden_dir):
    """The reference wires two FeatureMatcher objects over one extractor pair: FeatureTracker::fm (tracking thread) and
    System::mFeatureMatcher (LocalMapper thread: matchByProjectionRPredLBA, src/OptimizationBA.cpp:367).  The calls that do not read
    the resident frame run on the matcher's own stream and scratch, so the second matcher is safe next to a tracking loop."""
    import os
    cfg = synth.CONFIGS["small"]
    w, h = cfg["width"], cfg["height"]
    L, R = synth.stereo_batch(w, h, cfg["disparity"], range(90, 92))
    args = dict(nFeatures=cfg["n_features"], nLevels=cfg["n_levels"], width=w, height=h, max_images=2)
    feL, feR = gvs.FeatureExtractor(**args), gvs.FeatureExtractor(**args)
    fmTrack = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    fmMap = gvs.FeatureMatcher(cfg["fx"], cfg["baseline"], feL, feR, h)
    g = np.load(os.path.join(golden_dir, "small_s0.npz")); p = np.load(os.path.join(golden_dir, "small_s0_lba.npz"))
    pts = {k[3:]: p[k] for k in p.files if k.startswith("pt_")}
    want_n, want_out = int(p["lba_15_n"]), p["lba_15_out"]
    errors, first = [], {}

    def tracker():
        try:
            for i in range(30):
                fmTrack.track_stereo_batch(L, R)
                ri, li, de, cl = fmTrack.stereo_download(0, 2)
                if i == 0:
                    first["ri"] = ri.copy()
                assert np.array_equal(ri, first["ri"])
        except Exception as e:  # noqa: BLE001
            errors.append(("tracker", e))

    def mapper():
        try:
            for _ in range(30):
                n, out = fmMap.matchByProjectionRPredLBA(pts, g["kpL"], g["descL"], g["kpR"], g["descR"], g["leftIdxs"], g["rightIdxs"],
                                                         p["unF"], p["unFR"], w, h, 15.0)
                assert n == want_n and np.array_equal(out, want_out)
        except Exception as e:  # noqa: BLE001
            errors.append(("mapper", e))

    ts = [threading.Thread(target=tracker), threading.Thread(target=mapper)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
