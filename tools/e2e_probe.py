import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gtsam_vslam_b200 as gvs
from gtsam_vslam_b200 import synth
cfg = synth.CONFIGS['euroc']; w,h = cfg['width'], cfg['height']; F = 256
base = synth.stereo_batch(w, h, cfg['disparity'], range(16))
L = np.ascontiguousarray(np.tile(base[0], (F//16,1,1))); R = np.ascontiguousarray(np.tile(base[1], (F//16,1,1)))
hL = torch.from_numpy(L).pin_memory(); hR = torch.from_numpy(R).pin_memory()
a = dict(nFeatures=cfg['n_features'], nLevels=8, width=w, height=h, max_images=F)
feL, feR = gvs.FeatureExtractor(**a), gvs.FeatureExtractor(**a)
fm = gvs.FeatureMatcher(cfg['fx'], cfg['baseline'], feL, feR, h)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); fm.track_stereo_batch(hL.numpy(), hR.numpy()); t1 = time.perf_counter(); fm.sync(); t2 = T()
    offL, offR = feL.offsets(F), feR.offsets(F); t3 = T()
    kl, dl = feL.download(0, F, int(offL[F])); kr, dr = feR.download(0, F, int(offR[F])); t4 = T()
    s = fm.stereo_download(0, F); t5 = T()
    print(f"launch {1e3*(t1-t0):.2f} run {1e3*(t2-t0):.2f} offsets {1e3*(t3-t2):.2f} dl keys {1e3*(t4-t3):.2f} dl stereo {1e3*(t5-t4):.2f} total {1e3*(t5-t0):.2f}")
# raw H2D bandwidth
d = torch.empty_like(hL, device='cuda')
for _ in range(3):
    t0 = T(); d.copy_(hL, non_blocking=True); t1 = T(); print(f"H2D {hL.numel()/1e6:.0f} MB in {1e3*(t1-t0):.2f} ms = {hL.numel()/(t1-t0)/1e9:.1f} GB/s")
