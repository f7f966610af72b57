import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
cfg = synth.CONFIGS['euroc']; w,h = cfg['width'], cfg['height']; F = 256
base = synth.stereo_batch(w, h, cfg['disparity'], range(16))
L = np.ascontiguousarray(np.tile(base[0], (F//16,1,1))); R = np.ascontiguousarray(np.tile(base[1], (F//16,1,1)))
hL = torch.from_numpy(L).pin_memory().numpy(); hR = torch.from_numpy(R).pin_memory().numpy()
pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory().numpy()
for C in (1, 2, 4, 8):
    Fc = F // C
    hs = []
    for c in range(C):
        a = dict(nFeatures=cfg['n_features'], nLevels=8, width=w, height=h, max_images=Fc)
        feL, feR = gvs.FeatureExtractor(**a), gvs.FeatureExtractor(**a)
        fm = gvs.FeatureMatcher(cfg['fx'], cfg['baseline'], feL, feR, h)
        cap = Fc * feL.cap_per_image
        bufs = dict(kl=pin(cap*28, torch.uint8).view(gvs.KEYPOINT_DTYPE), kr=pin(cap*28, torch.uint8).view(gvs.KEYPOINT_DTYPE), dl=pin(cap*32, torch.uint8), dr=pin(cap*32, torch.uint8),
                    ri=pin(cap, torch.int32), li=pin(cap, torch.int32), de=pin(cap, torch.float32), cl=pin(cap, torch.uint8))
        hs.append((feL, feR, fm, bufs))
    def step():
        for c, (feL, feR, fm, b) in enumerate(hs):
            fm.track_stereo_batch(hL[c*Fc:(c+1)*Fc], hR[c*Fc:(c+1)*Fc])
        tot = 0
        for c, (feL, feR, fm, b) in enumerate(hs):
            nl = feL.download_into(0, Fc, b['kl'], b['dl']); nr = feR.download_into(0, Fc, b['kr'], b['dr'])
            fm.stereo_download_into(0, Fc, b['ri'], b['li'], b['de'], b['cl'])
            tot += nl + nr
        return tot
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): tot = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"chunks {C}: {dt*1e3:.2f} ms/step -> {F/dt:.0f} frames/s  (keys {tot})", flush=True)
    del hs
