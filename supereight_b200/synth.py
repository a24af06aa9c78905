"""Synthetic 640x480 depth streams for the hot-path benchmark and the parity tests.

SURVEY.md section 8(d) "Concrete synthetic inputs":
  S1 planar sweep : infinite wall z_world = 0.75*dim, camera at (0.5, 0.5, 0.1)*dim looking +z,
                    pose_i = translate x by 2 mm * i and yaw by 0.1 deg * i.
  S2 box room     : axis-aligned room [0.1, 0.9]*dim cubed seen from its centre, one full 360 deg
                    yaw over `n_frames` frames.
Depth is the exact z-depth of the ray through the pixel centre (integer pixel coordinates, the
convention of the reference's raycast, rendering.cpp:63), rounded to uint16 millimetres, which is
what the reference's readers hand to DenseSLAMSystem::preprocessing.  Optional Gaussian noise
(sigma 2 mm, MT19937 seed 1234 + g for sequence g) and 1 % drop-outs (seed 4321 + g) exercise the depth==0 paths.

Pure numpy, deterministic, no GPU: inputs only -- not part of the oracle and not a compute path.
"""
from __future__ import annotations

import numpy as np

DEFAULT_K = (481.2, 480.0, 320.0, 240.0)  # fx, fy, cx, cy (positive fy for synthetic data)


def yaw_pose(tx: float, ty: float, tz: float, yaw_rad: float) -> np.ndarray:
    """Camera-to-world 4x4 (row-major float32): rotation about the world y axis, then translation."""
    c, s = np.cos(yaw_rad), np.sin(yaw_rad)
    T = np.array([[c, 0.0, s, tx],
                  [0.0, 1.0, 0.0, ty],
                  [-s, 0.0, c, tz],
                  [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)
    return T.astype(np.float32)


def _rays(W: int, H: int, k, pose: np.ndarray):
    fx, fy, cx, cy = k
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)  # z-depth parametrisation
    R = pose[:3, :3].astype(np.float64)
    t = pose[:3, 3].astype(np.float64)
    dw = dc @ R.T
    return dw, t


def _finish(depth_m: np.ndarray, frame: int, noise_mm: float, dropout: float, seed: int) -> np.ndarray:
    """`seed` = g, the index of the sequence (BASELINE.json config 5: "seeds 1234+g"): its noise comes from MT19937 seeded
    1234 + g and its drop-outs from MT19937 seeded 4321 + g, one independent sub-stream per frame (SeedSequence spawn key =
    the frame number), so frames can be generated in any order."""
    d = depth_m * 1000.0
    if noise_mm > 0:
        rng = np.random.Generator(np.random.MT19937(np.random.SeedSequence(1234 + seed, spawn_key=(frame,))))
        d = d + rng.normal(0.0, noise_mm, size=d.shape)
    d = np.where(np.isfinite(d) & (d > 0) & (d < 65535), d, 0.0)
    out = np.rint(d).astype(np.uint16)
    if dropout > 0:
        rng = np.random.Generator(np.random.MT19937(np.random.SeedSequence(4321 + seed, spawn_key=(frame,))))
        out[rng.random(out.shape) < dropout] = 0
    return np.ascontiguousarray(out)


def planar_sweep_pose(frame: int, dim: float) -> np.ndarray:
    return yaw_pose(0.5 * dim + 0.002 * frame, 0.5 * dim, 0.1 * dim, np.deg2rad(0.1 * frame))


def planar_sweep(frame: int, dim: float, W: int = 640, H: int = 480, k=DEFAULT_K,
                 noise_mm: float = 0.0, dropout: float = 0.01, seed: int = 0):
    """S1.  Returns (depth_mm uint16 [H, W], pose float32 [4, 4])."""
    pose = planar_sweep_pose(frame, dim)
    dw, t = _rays(W, H, k, pose)
    zw = 0.75 * dim
    with np.errstate(divide="ignore", invalid="ignore"):
        d = (zw - t[2]) / dw[..., 2]
    d = np.where(dw[..., 2] > 1e-9, d, 0.0)
    return _finish(d, frame, noise_mm, dropout, seed), pose


def box_room_pose(frame: int, dim: float, n_frames: int = 300) -> np.ndarray:
    return yaw_pose(0.5 * dim, 0.5 * dim, 0.5 * dim, 2.0 * np.pi * frame / n_frames)


def box_room(frame: int, dim: float, W: int = 640, H: int = 480, k=DEFAULT_K, n_frames: int = 300,
             noise_mm: float = 0.0, dropout: float = 0.01, seed: int = 0):
    """S2.  Returns (depth_mm uint16 [H, W], pose float32 [4, 4])."""
    pose = box_room_pose(frame, dim, n_frames)
    dw, t = _rays(W, H, k, pose)
    lo, hi = 0.1 * dim, 0.9 * dim
    best = np.full(dw.shape[:2], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            for plane in (lo, hi):
                s = (plane - t[a]) / dw[..., a]
                s = np.where(s > 1e-9, s, np.inf)
                best = np.minimum(best, s)
    return _finish(best, frame, noise_mm, dropout, seed), pose


def look_at_pose(eye, target, roll_rad: float = 0.0) -> np.ndarray:
    """Camera-to-world pose with +z looking from `eye` to `target`, y roughly down-ish (right-handed)."""
    eye = np.asarray(eye, np.float64); target = np.asarray(target, np.float64)
    z = target - eye; z /= np.linalg.norm(z)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    c, s = np.cos(roll_rad), np.sin(roll_rad)
    x, y = c * x + s * y, -s * x + c * y
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T.astype(np.float32)


def corner_view_pose(frame: int, dim: float) -> np.ndarray:
    """S3 (tracking tests): the camera looks into a corner of the box room (three mutually perpendicular
    planes in view, so ICP is well conditioned) while drifting 3 mm and ~0.15 deg per frame."""
    eye = np.array([0.45, 0.5, 0.42]) * dim + np.array([0.003, -0.002, 0.0025]) * frame
    target = np.array([0.9, 0.9, 0.9]) * dim + np.array([-0.004, 0.003, 0.0]) * frame
    return look_at_pose(eye, target, roll_rad=np.deg2rad(0.05 * frame))


def corner_view(frame: int, dim: float, W: int = 640, H: int = 480, k=DEFAULT_K, noise_mm: float = 0.0,
                dropout: float = 0.0, seed: int = 0):
    """Depth of the box room [0.1, 0.9]*dim^3 from corner_view_pose.  Returns (depth_mm, pose)."""
    pose = corner_view_pose(frame, dim)
    dw, t = _rays(W, H, k, pose)
    lo, hi = 0.1 * dim, 0.9 * dim
    best = np.full(dw.shape[:2], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            for plane in (lo, hi):
                s = (plane - t[a]) / dw[..., a]
                s = np.where(s > 1e-9, s, np.inf)
                best = np.minimum(best, s)
    return _finish(best, frame, noise_mm, dropout, seed), pose
