"""Synthetic frame-pair generator — the reference's simulation, vectorised.

Restates the data distribution of the reference's experiment generator so every
BASELINE.json config can be produced without files:

  poses          BaseExperiments::SamplePoses      src/simulation/experiments.cc:43-70
  points         BaseExperiments::SamplePoints     src/simulation/experiments.cc:72-129
  covariances    CovFromParameters                 src/simulation/experiments.cc:174-185
                 StandardExperiments::SampleCovariances
                                                   src/simulation/standard_experiments.cc:86-123
  noise          StandardExperiments::AddNoise     src/simulation/standard_experiments.cc:125-157
  bearing + cov  GetFeatures + UnscentedTransform  src/simulation/sim_common.cc:72-107,
                                                   src/common/common.cc:467-525
  start pose     ReadExperiments                   src/simulation/sim_common.cc:205-231

The reference draws from std::mt19937 / std::default_random_engine streams; it
ships no generated data or seeds to compare against, so this module follows the
*distributions*, not the bit streams (numpy PCG64, one stream per call).

Output layout is the C-ABI's (include/pnec_b200.h): bearing vectors [B*N][3],
covariances [B*N][9] column-major, poses [B][7] = (qx qy qz qw tx ty tz).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

OMNIDIRECTIONAL = "omnidirectional"
PINHOLE = "pinhole"

NOISE_TYPES = (
    "isotropic_homogenous",
    "anisotropic_homogenous",
    "isotropic_inhomogenous",
    "anisotropic_inhomogenous",
)

FOCAL_LENGTH = 800.0


@dataclass
class FramePairBatch:
    """B frame pairs, flat C-ABI layout.  `offsets` is None for uniform batches."""

    bvs_host: np.ndarray  # (total, 3) f1
    bvs_target: np.ndarray  # (total, 3) f2
    covs_target: np.ndarray  # (total, 9) column-major 3x3
    init_poses: np.ndarray  # (B, 7)
    gt_poses: np.ndarray  # (B, 7), unit translation
    n_per_problem: int
    offsets: np.ndarray | None = None
    covs_host: np.ndarray | None = None

    @property
    def num_problems(self) -> int:
        return self.init_poses.shape[0]

    @property
    def total(self) -> int:
        return self.bvs_host.shape[0]

    def problem(self, b: int):
        s, e = self.range(b)
        ch = None if self.covs_host is None else self.covs_host[s:e]
        return self.bvs_host[s:e], self.bvs_target[s:e], self.covs_target[s:e], ch

    def range(self, b: int):
        if self.offsets is None:
            return b * self.n_per_problem, (b + 1) * self.n_per_problem
        return int(self.offsets[b]), int(self.offsets[b + 1])


# ------------------------------------------------------------------ geometry


def skew(v: np.ndarray) -> np.ndarray:
    """[v]x for (...,3) -> (...,3,3); pnec::common::SkewFromVector, common.cc:96-101."""
    z = np.zeros_like(v[..., 0])
    return np.stack(
        [
            np.stack([z, -v[..., 2], v[..., 1]], -1),
            np.stack([v[..., 2], z, -v[..., 0]], -1),
            np.stack([-v[..., 1], v[..., 0], z], -1),
        ],
        -2,
    )


def rotation_between_points(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """RotationBetweenPoints, common.cc:118-124: I + [v]x + [v]x^2/(1+c), v = p1 x p2."""
    v = np.cross(p1, p2)
    c = np.sum(p1 * p2, -1)
    vh = skew(v)
    return np.eye(3) + vh + (vh @ vh) / (1.0 + c)[..., None, None]


def angle_axis_to_matrix(angle: np.ndarray, axis: np.ndarray) -> np.ndarray:
    k = skew(axis)
    s = np.sin(angle)[..., None, None]
    c = np.cos(angle)[..., None, None]
    return np.eye(3) + s * k + (1.0 - c) * (k @ k)


def matrix_to_quaternion(R: np.ndarray) -> np.ndarray:
    """Eigen::Quaterniond(Matrix3d) (Shepperd branches), returns (...,4) = x y z w."""
    R = np.asarray(R, dtype=np.float64)
    shape = R.shape[:-2]
    Rf = R.reshape(-1, 3, 3)
    q = np.empty((Rf.shape[0], 4))
    tr = Rf[:, 0, 0] + Rf[:, 1, 1] + Rf[:, 2, 2]
    pos = tr > 0
    if np.any(pos):
        M = Rf[pos]
        t = np.sqrt(tr[pos] + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q[pos, 3] = w
        q[pos, 0] = (M[:, 2, 1] - M[:, 1, 2]) * t
        q[pos, 1] = (M[:, 0, 2] - M[:, 2, 0]) * t
        q[pos, 2] = (M[:, 1, 0] - M[:, 0, 1]) * t
    neg = np.nonzero(~pos)[0]
    for idx in neg:
        M = Rf[idx]
        i = 0
        if M[1, 1] > M[0, 0]:
            i = 1
        if M[2, 2] > M[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(M[i, i] - M[j, j] - M[k, k] + 1.0)
        q[idx, i] = 0.5 * t
        t = 0.5 / t
        q[idx, 3] = (M[k, j] - M[j, k]) * t
        q[idx, j] = (M[j, i] + M[i, j]) * t
        q[idx, k] = (M[k, i] + M[i, k]) * t
    return q.reshape(shape + (4,))


def quaternion_to_matrix(q: np.ndarray) -> np.ndarray:
    """Eigen::Quaternion::toRotationMatrix for (...,4) x y z w."""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.stack(
        [
            np.stack([1 - (tyy + tzz), txy - twz, txz + twy], -1),
            np.stack([txy + twz, 1 - (txx + tzz), tyz - twx], -1),
            np.stack([txz - twy, tyz + twx, 1 - (txx + tyy)], -1),
        ],
        -2,
    )


def pose7_from_rt(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    return np.concatenate([matrix_to_quaternion(R), t], -1)


def _uniform_sphere(rng: np.random.Generator, shape) -> np.ndarray:
    theta = 2 * np.pi * rng.random(shape)
    phi = np.arccos(1.0 - 2.0 * rng.random(shape))
    return np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], -1)


def _normalize(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


# --------------------------------------------------------------- generators


def sample_poses(rng, B: int, max_euler_angle: float = 0.5, translation: bool = True):
    """Relative pose (R, t) of frame 2 in frame 1; experiments.cc:43-70."""
    roll, pitch, yaw = ((rng.random(B) * 2.0 - 1.0) * max_euler_angle for _ in range(3))
    ex, ey, ez = np.eye(3)
    R = (
        angle_axis_to_matrix(roll, np.broadcast_to(ex, (B, 3)))
        @ angle_axis_to_matrix(pitch, np.broadcast_to(ey, (B, 3)))
        @ angle_axis_to_matrix(yaw, np.broadcast_to(ez, (B, 3)))
    )
    if translation:
        direction = _uniform_sphere(rng, (B,))
        t = (2.0 * rng.random(B))[:, None] * direction
    else:
        t = np.zeros((B, 3))
    return R, t


def sample_points(rng, R: np.ndarray, t: np.ndarray, N: int, camera: str):
    """Image-space points (|p| or p_z = focal length) in both frames; experiments.cc:72-129."""
    B = R.shape[0]
    if camera == PINHOLE:
        a, width, height, max_depth = 0.4, 1.0, 1.5, 5.0
        depth = ((1 - a) * rng.random((B, N)) + a) * max_depth
        x = (rng.random((B, N)) - 0.5) * width
        y = (rng.random((B, N)) - 0.5) * height
        point = np.stack([x, y, np.ones_like(x)], -1) * depth[..., None]
    elif camera == OMNIDIRECTIONAL:
        direction = _uniform_sphere(rng, (B, N))
        point = (4.0 * rng.random((B, N)) + 4.0)[..., None] * direction
    else:
        raise ValueError(camera)
    # pose_2.inverse() * point = R^T (point - t)
    p2 = np.einsum("bji,bnj->bni", R, point - t[:, None, :])
    p1n, p2n = _normalize(point), _normalize(p2)
    if camera == PINHOLE:
        p1 = p1n / p1n[..., 2:3] * FOCAL_LENGTH
        p2 = p2n / p2n[..., 2:3] * FOCAL_LENGTH
    else:
        p1 = p1n * FOCAL_LENGTH
        p2 = p2n * FOCAL_LENGTH
    return p1, p2


def sample_covariances_2d(rng, shape, noise_level: float, noise_type: str) -> np.ndarray:
    """2x2 image-plane covariances (...,2,2); standard_experiments.cc:86-123 +
    CovFromParameters experiments.cc:174-185."""
    if noise_type not in NOISE_TYPES:
        raise ValueError(noise_type)
    alpha = np.zeros(shape)
    beta = np.full(shape, 0.5)
    scale = np.ones(shape)
    if noise_type == "anisotropic_homogenous":
        # beta is drawn once per experiment before the loop (:98) and kept
        alpha = rng.random(shape) * np.pi
        beta = np.broadcast_to(((rng.random(shape[:-1]) + 1.0) / 2.0)[..., None], shape).copy()
    elif noise_type == "isotropic_inhomogenous":
        scale = rng.random(shape) + 0.5
    elif noise_type == "anisotropic_inhomogenous":
        alpha = rng.random(shape) * np.pi
        beta = (rng.random(shape) + 1.0) / 2.0
        scale = rng.random(shape) + 0.5
    ca, sa = np.cos(alpha), np.sin(alpha)
    d0, d1 = beta, 1.0 - beta
    c00 = ca * ca * d0 + sa * sa * d1
    c01 = ca * sa * (d0 - d1)
    c11 = sa * sa * d0 + ca * ca * d1
    cov = np.stack([np.stack([c00, c01], -1), np.stack([c01, c11], -1)], -2)
    return cov * (noise_level * scale)[..., None, None]


def _chol2(cov2: np.ndarray):
    """Lower Cholesky factor entries of (...,2,2)."""
    l00 = np.sqrt(cov2[..., 0, 0])
    l10 = cov2[..., 1, 0] / l00
    l11 = np.sqrt(cov2[..., 1, 1] - l10 * l10)
    return l00, l10, l11


def add_noise(rng, p2: np.ndarray, cov2: np.ndarray, camera: str):
    """Noisy frame-2 points and their 3x3 image-plane covariances;
    standard_experiments.cc:125-157."""
    l00, l10, l11 = _chol2(cov2)
    n0 = rng.standard_normal(p2.shape[:-1])
    n1 = rng.standard_normal(p2.shape[:-1])
    local = np.stack([l00 * n0, l10 * n0 + l11 * n1, np.zeros_like(n0)], -1)
    cov3 = np.zeros(p2.shape[:-1] + (3, 3))
    cov3[..., :2, :2] = cov2
    if camera == OMNIDIRECTIONAL:
        ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), p2.shape)
        Rp = rotation_between_points(ez, _normalize(p2))
        noise = np.einsum("...ij,...j->...i", Rp, local)
        cov3 = Rp @ cov3 @ np.swapaxes(Rp, -1, -2)
    else:
        noise = local
    return p2 + noise, cov3


def unscented_transform(mu: np.ndarray, cov: np.ndarray, camera: str, kappa: float = 1.0):
    """pnec::common::UnscentedTransform with K_inv = I, common.cc:467-525.
    mu (...,3) image-space point, cov (...,3,3) -> bearing covariance (...,3,3)."""
    n = 2
    if camera == OMNIDIRECTIONAL:
        ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), mu.shape)
        rot = rotation_between_points(ez, _normalize(mu))
        local = (np.swapaxes(rot, -1, -2) @ cov @ rot)[..., :2, :2]
    else:
        rot = None
        local = cov[..., :2, :2]
    l00, l10, l11 = _chol2(local)
    z = np.zeros_like(l00)
    c0 = np.stack([l00, l10, z], -1)  # C.col(0)
    c1 = np.stack([z, l11, z], -1)  # C.col(1)
    if rot is not None:
        c0 = np.einsum("...ij,...j->...i", rot, c0)
        c1 = np.einsum("...ij,...j->...i", rot, c1)
    w0 = kappa / (n + kappa)
    wi = 0.5 / (n + kappa)
    pts = [mu, mu + c0, mu + c1, mu - c0, mu - c1]
    ws = [w0, wi, wi, wi, wi]
    tp = [_normalize(p) for p in pts]
    mean = sum(w * p for w, p in zip(ws, tp))
    sigma = sum(w * (p - mean)[..., :, None] * (p - mean)[..., None, :] for w, p in zip(ws, tp))
    return sigma


def perturb_poses(rng, R: np.ndarray, t: np.ndarray, init_scaling: float = 1.0):
    """Start pose near the ground truth; sim_common.cc:205-231."""
    B = R.shape[0]
    axis = _uniform_sphere(rng, (B,))
    angle = np.sqrt(rng.random(B)) * 0.01 * init_scaling
    R_off = angle_axis_to_matrix(angle, axis)
    t_off = (np.sqrt(rng.random(B)) * 0.01 * init_scaling)[:, None] * _uniform_sphere(rng, (B,))
    R_init = R_off @ R
    t_init = np.einsum("bij,bj->bi", R_off, t) + t_off
    return R_init, _normalize(t_init)


def make_batch(
    num_problems: int,
    n_per_problem: int,
    *,
    seed: int = 1,
    camera: str = OMNIDIRECTIONAL,
    noise_type: str = "anisotropic_inhomogenous",
    noise_level: float = 1.0,
    translation: bool = True,
    init_scaling: float = 1.0,
    counts: np.ndarray | None = None,
    chunk: int = 512,
) -> FramePairBatch:
    """B frame pairs of N correspondences (or ragged `counts[b]` each) in C-ABI layout."""
    rng = np.random.default_rng(seed)
    B = int(num_problems)
    if counts is not None:
        counts = np.asarray(counts, dtype=np.int64)
        assert counts.shape == (B,)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        total = int(offsets[-1])
        n_gen = int(counts.max()) if B else 0
    else:
        offsets = None
        total = B * n_per_problem
        n_gen = n_per_problem
    f1 = np.empty((total, 3))
    f2 = np.empty((total, 3))
    cov = np.empty((total, 9))
    init = np.empty((B, 7))
    gt = np.empty((B, 7))
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        R, t = sample_poses(rng, e - s, translation=translation)
        p1, p2 = sample_points(rng, R, t, n_gen, camera)
        cov2 = sample_covariances_2d(rng, (e - s, n_gen), noise_level, noise_type)
        p2n, cov3 = add_noise(rng, p2, cov2, camera)
        sig = unscented_transform(p2n, cov3, camera)
        b1, b2 = _normalize(p1), _normalize(p2n)
        # column-major 3x3 == transpose of the row-major numpy matrix
        sig9 = np.swapaxes(sig, -1, -2).reshape(e - s, n_gen, 9)
        R_init, t_init = perturb_poses(rng, R, t, init_scaling)
        init[s:e] = pose7_from_rt(R_init, t_init)
        tn = np.linalg.norm(t, axis=-1, keepdims=True)
        gt[s:e] = pose7_from_rt(R, np.where(tn > 0, t / np.where(tn > 0, tn, 1.0), t))
        if offsets is None:
            f1[s * n_gen : e * n_gen] = b1.reshape(-1, 3)
            f2[s * n_gen : e * n_gen] = b2.reshape(-1, 3)
            cov[s * n_gen : e * n_gen] = sig9.reshape(-1, 9)
        else:
            for b in range(s, e):
                o0, o1 = offsets[b], offsets[b + 1]
                k = o1 - o0
                f1[o0:o1] = b1[b - s, :k]
                f2[o0:o1] = b2[b - s, :k]
                cov[o0:o1] = sig9[b - s, :k]
    return FramePairBatch(f1, f2, cov, init, gt, 0 if offsets is not None else n_per_problem, offsets)


def kitti_like_counts(num_pairs: int = 4540, mean: float = 2000.0, std: float = 300.0,
                      lo: int = 500, hi: int = 3500, seed: int = 7) -> np.ndarray:
    """Per-pair correspondence counts of BASELINE config C4."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(mean, std, num_pairs)), lo, hi).astype(np.int64)


def with_host_covariances(batch: FramePairBatch, seed: int = 11, noise_level: float = 1.0,
                          camera: str = OMNIDIRECTIONAL) -> FramePairBatch:
    """Adds frame-1 covariances (for the SYMMETRIC variant, pypnec.pyceres).  The
    reference's simulator never produces these (covs_1 is unused,
    sim_common.cc:157-159), so they are drawn from the same family as frame 2."""
    rng = np.random.default_rng(seed)
    total = batch.total
    cov2 = sample_covariances_2d(rng, (1, total), noise_level, "anisotropic_inhomogenous")[0]
    cov3 = np.zeros((total, 3, 3))
    cov3[:, :2, :2] = cov2
    if camera == PINHOLE:
        mu = batch.bvs_host / batch.bvs_host[:, 2:3] * FOCAL_LENGTH
    else:
        mu = batch.bvs_host * FOCAL_LENGTH
    if camera == OMNIDIRECTIONAL:
        ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), mu.shape)
        Rp = rotation_between_points(ez, _normalize(mu))
        cov3 = Rp @ cov3 @ np.swapaxes(Rp, -1, -2)
    sig = unscented_transform(mu, cov3, camera)
    batch.covs_host = np.swapaxes(sig, -1, -2).reshape(total, 9).copy()
    return batch
