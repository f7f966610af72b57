"""Deterministic synthetic stereo imagery.

A canvas of random axis-aligned rectangles, lightly smoothed, cropped at two horizontal offsets so
that x_R = x_L - d, plus independent sensor noise.  Pure numpy so that the same seed gives the same
bytes in the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np

__all__ = ["stereo_pair", "stereo_batch", "shifted_sequence", "stereo_sequence", "sequence_motion", "CONFIGS"]

# FE params follow reference config/config_kitti_00.yaml:48-55, config/config_MH_01.yaml:101-108,
# include/FeatureExtractor.h:80 (defaults, mono) ; camera values :20,:45 / :33,:98.
CONFIGS = {
    "kitti": dict(width=1241, height=376, n_features=2000, n_levels=8, fx=718.856, baseline=0.53716, disparity=24),
    "euroc": dict(width=752, height=480, n_features=1000, n_levels=8, fx=435.2047, baseline=0.110074, disparity=12),
    "mono": dict(width=752, height=480, n_features=2000, n_levels=8, fx=435.2047, baseline=0.110074, disparity=12),
    "hires": dict(width=2048, height=1536, n_features=8000, n_levels=8, fx=1200.0, baseline=0.12, disparity=40),
    "small": dict(width=320, height=240, n_features=300, n_levels=4, fx=200.0, baseline=0.1, disparity=8),
}


def _smooth(img: np.ndarray, sigma: float = 0.8) -> np.ndarray:
    r = 3
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    h = np.zeros_like(img)
    for i in range(2 * r + 1):
        h += k[i] * p[:, i:i + img.shape[1]]
    p = np.pad(h, ((r, r), (0, 0)), mode="reflect")
    v = np.zeros_like(img)
    for i in range(2 * r + 1):
        v += k[i] * p[i:i + img.shape[0], :]
    return v


def _canvas(rng: np.random.Generator, height: int, cw: int, n_rect: int) -> np.ndarray:
    c = np.full((height, cw), 128.0, dtype=np.float32)
    x0 = rng.integers(0, cw, n_rect)
    y0 = rng.integers(0, height, n_rect)
    ww = rng.integers(3, 40, n_rect)
    hh = rng.integers(3, 40, n_rect)
    g = rng.uniform(0.0, 256.0, n_rect).astype(np.float32)
    for i in range(n_rect):
        c[y0[i]:y0[i] + hh[i], x0[i]:x0[i] + ww[i]] = g[i]
    return _smooth(c)


def _noisy_u8(rng: np.random.Generator, img: np.ndarray) -> np.ndarray:
    n = rng.normal(0.0, 1.5, img.shape).astype(np.float32)
    return np.clip(np.rint(img + n), 0, 255).astype(np.uint8)


def stereo_pair(width: int, height: int, disparity: int = 24, seed: int = 0, ramp: bool = False):
    """Return (left, right) uint8 images, C-contiguous, shape (height, width).

    ramp=True makes the disparity depend on the row, d(y) = 8 + 0.15*y (rounded), as in section 8(d).
    """
    maxd = disparity if not ramp else int(8 + 0.15 * height) + 1
    rng = np.random.default_rng(seed)
    cw = width + maxd + 32
    c = _canvas(rng, height, cw, (width * height) // 90)
    left = c[:, :width]
    if ramp:
        right = np.empty_like(left)
        for y in range(height):
            d = int(round(8 + 0.15 * y))
            right[y] = c[y, d:d + width]
    else:
        right = c[:, disparity:disparity + width]
    return _noisy_u8(rng, left), _noisy_u8(rng, right)


def stereo_batch(width: int, height: int, disparity: int, seeds) -> tuple[np.ndarray, np.ndarray]:
    """Independent frames, one per seed: two (F, H, W) uint8 arrays."""
    seeds = list(seeds)
    L = np.empty((len(seeds), height, width), np.uint8)
    R = np.empty_like(L)
    for i, s in enumerate(seeds):
        L[i], R[i] = stereo_pair(width, height, disparity, s)
    return L, R


def shifted_sequence(width: int, height: int, n_frames: int, seed: int = 0, step=(2, 1)) -> np.ndarray:
    """Temporal sequence: frame k is the (step*k)-shifted crop of one big canvas (section 8(d) config 4)."""
    rng = np.random.default_rng(seed)
    cw = width + step[0] * n_frames + 32
    ch = height + step[1] * n_frames + 32
    c = _canvas(rng, ch, cw, (cw * ch) // 90)
    out = np.empty((n_frames, height, width), np.uint8)
    for k in range(n_frames):
        out[k] = _noisy_u8(rng, c[step[1] * k: step[1] * k + height, step[0] * k: step[0] * k + width])
    return out


def stereo_sequence(width: int, height: int, disparity: int, n_frames: int, seed: int = 0, step=(2, 1)):
    """Temporal stereo sequence: frame k is the crop of one big canvas at offset step*k (left) and step*k + (disparity, 0) (right),
    i.e. a fronto-parallel plane at depth fx*b/disparity seen by a camera that translates by step pixels per frame.  Two
    (n_frames, H, W) uint8 arrays.  Bounded canvas: the walk turns around every 64 frames so that long batches stay small."""
    rng = np.random.default_rng(seed)
    span = 64
    cw = width + step[0] * span + disparity + 32
    ch = height + step[1] * span + 32
    c = _canvas(rng, ch, cw, (cw * ch) // 90)
    L = np.empty((n_frames, height, width), np.uint8)
    R = np.empty_like(L)
    for k in range(n_frames):
        ox, oy = sequence_offset(k, step, span)
        L[k] = _noisy_u8(rng, c[oy: oy + height, ox: ox + width])
        R[k] = _noisy_u8(rng, c[oy: oy + height, ox + disparity: ox + disparity + width])
    return L, R


def sequence_offset(k: int, step=(2, 1), span: int = 64):
    """Canvas offset of frame k of stereo_sequence: a triangle wave of period 2*span."""
    m = k % (2 * span)
    t = m if m <= span else 2 * span - m
    return step[0] * t, step[1] * t


def sequence_motion(cfg: dict, n_frames: int, prev, step=(2, 1), span: int = 64):
    """Camera poses that explain stereo_sequence: pose_k = translation of offset_k * z0 / f, z0 = fx * baseline / disparity
    (identity rotation, fy = fx, principal point at the image centre).  Returns dict(fx, fy, cx, cy, extrinsics, pose (n,4,4),
    pose_prev, to_camera, to_camera_r (n,4,4) for the given predecessor list)."""
    fx = float(cfg["fx"]); b = float(cfg["baseline"]); d = float(cfg["disparity"])
    z0 = fx * b / d
    pose = np.tile(np.eye(4), (n_frames, 1, 1))
    for k in range(n_frames):
        ox, oy = sequence_offset(k, step, span)
        pose[k, 0, 3] = ox * z0 / fx
        pose[k, 1, 3] = oy * z0 / fx
    ext = np.eye(4); ext[0, 3] = b
    prev = np.asarray(prev, np.int32)
    pose_prev = np.stack([pose[p] if p >= 0 else np.eye(4) for p in prev])
    to_cam = np.stack([np.linalg.inv(pose[k]) for k in range(n_frames)])
    to_cam_r = np.stack([np.linalg.inv(pose[k] @ ext) for k in range(n_frames)])
    return dict(fx=fx, fy=fx, cx=cfg["width"] / 2.0, cy=cfg["height"] / 2.0, extrinsics=ext, pose=pose, pose_prev=pose_prev,
                to_camera=to_cam, to_camera_r=to_cam_r, prev=prev)
