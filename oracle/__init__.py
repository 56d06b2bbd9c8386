"""ctypes front-end of the CPU oracle (oracle/pnec_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(pnec_b200/) never imports this module.

Parity status: residual functors pinned against the reference's numpy energies
(tests/golden/); the Ceres LM loop is restated from upstream behaviour and is
PARITY UNPINNED (Ceres is neither vendored by the reference nor installed here).
Front stages of PNEC::Solve (oracle/pnec_oracle_frame.c): the eigensolver's LM is
pinned against scipy's MINPACK, its eigenvalue/gradient against numpy; that the
restated steps are what opengv executes is PARITY UNPINNED (opengv unavailable).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpnec_oracle.so")

NEC, TARGET, HOST, SYMMETRIC = 0, 1, 2, 3
JAC_NUMERIC_CENTRAL, JAC_ANALYTIC = 0, 1

STATUS_NAMES = (
    "converged_function",
    "converged_parameter",
    "converged_gradient",
    "converged_radius",
    "max_iterations",
    "failure",
    "nonfinite",
    "empty",
)


class OracleOpts(ctypes.Structure):
    _fields_ = [
        ("variant", ctypes.c_int32),
        ("max_num_iterations", ctypes.c_int32),
        ("max_num_consecutive_invalid_steps", ctypes.c_int32),
        ("jacobi_scaling", ctypes.c_int32),
        ("regularization", ctypes.c_double),
        ("function_tolerance", ctypes.c_double),
        ("gradient_tolerance", ctypes.c_double),
        ("parameter_tolerance", ctypes.c_double),
        ("initial_trust_region_radius", ctypes.c_double),
        ("max_trust_region_radius", ctypes.c_double),
        ("min_trust_region_radius", ctypes.c_double),
        ("min_relative_decrease", ctypes.c_double),
        ("min_lm_diagonal", ctypes.c_double),
        ("max_lm_diagonal", ctypes.c_double),
        ("jacobian_mode", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class OracleInfo(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("iterations", ctypes.c_int32),
        ("num_successful_steps", ctypes.c_int32),
        ("num_unsuccessful_steps", ctypes.c_int32),
        ("initial_cost", ctypes.c_double),
        ("final_cost", ctypes.c_double),
        ("final_radius", ctypes.c_double),
        ("final_gradient_max_norm", ctypes.c_double),
    ]


class OracleEsInfo(ctypes.Structure):
    """Outcome of the restated opengv eigensolver (oracle/pnec_oracle_frame.c)."""
    _fields_ = [
        ("lm_info", ctypes.c_int32),  # MINPACK info code
        ("nfev", ctypes.c_int32),
        ("iterations", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("smallest_ev", ctypes.c_double),
        ("cayley", ctypes.c_double * 3),
    ]


class OracleRansacOpts(ctypes.Structure):
    """opengv::sac::Ransac settings as PNEC::Eigensolver uses them (pnec.cc:241-251)."""
    _fields_ = [
        ("max_iterations", ctypes.c_int32),
        ("sample_size", ctypes.c_int32),
        ("sequential", ctypes.c_int32),  # 1: opengv's sequential state (persistent shuffle, chained start)
        ("reserved", ctypes.c_int32),
        ("threshold", ctypes.c_double),
        ("probability", ctypes.c_double),
        ("max_variation", ctypes.c_double),
        ("seed", ctypes.c_uint64),
    ]


class OracleFrameOpts(ctypes.Structure):
    """pnec::rel_pose_estimation::Options as PNEC::Solve reads it (pnec_config.h:46-65)."""
    _fields_ = [
        ("use_nec", ctypes.c_int32),
        ("use_ceres", ctypes.c_int32),
        ("weighted_iterations", ctypes.c_int32),
        ("fibonacci_samples", ctypes.c_int32),
        ("scf_steps", ctypes.c_int32),
        ("use_ransac", ctypes.c_int32),
        ("ceres", OracleOpts),
        ("ransac", OracleRansacOpts),
    ]


# Options / C-ABI names of the RANSAC settings -> fields of OracleRansacOpts
_RANSAC_NAMES = {"max_ransac_iterations": "max_iterations", "ransac_sample_size": "sample_size",
                 "ransac_seed": "seed", "ransac_threshold": "threshold", "ransac_probability": "probability",
                 "ransac_max_variation": "max_variation", "ransac_sequential": "sequential"}


INFO_DTYPE = np.dtype(
    [
        ("status", np.int32),
        ("iterations", np.int32),
        ("num_successful_steps", np.int32),
        ("num_unsuccessful_steps", np.int32),
        ("initial_cost", np.float64),
        ("final_cost", np.float64),
        ("final_radius", np.float64),
        ("final_gradient_max_norm", np.float64),
    ]
)

_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/pnec_oracle.c (gcc) into oracle/_build/; returns the path."""
    srcs = [os.path.join(_HERE, f) for f in ("pnec_oracle.c", "pnec_oracle_frame.c")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs
    ):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_opts_default.argtypes = [ctypes.POINTER(OracleOpts)]
        L.oracle_opts_default.restype = None
        L.oracle_eval.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, dp, dp, dp, dp,
                                  ctypes.c_double, dp, dp, dp, dp]
        L.oracle_eval.restype = ctypes.c_int
        L.oracle_solve.argtypes = [ctypes.POINTER(OracleOpts), ctypes.c_int64, dp, dp, dp, dp, dp,
                                   dp, ctypes.POINTER(OracleInfo)]
        L.oracle_solve.restype = ctypes.c_int
        L.oracle_solve_batch.argtypes = [ctypes.POINTER(OracleOpts), ctypes.c_int64,
                                         ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), dp, dp,
                                         dp, dp, dp, dp, ctypes.c_void_p, ctypes.c_int]
        L.oracle_solve_batch.restype = ctypes.c_int
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_angles_from_vec.argtypes = [dp, dp, dp]
        L.oracle_angles_from_vec.restype = None
        L.oracle_rotational_difference_deg.argtypes = [dp, dp]
        L.oracle_rotational_difference_deg.restype = ctypes.c_double
        L.oracle_translational_difference_deg.argtypes = [dp, dp, ctypes.c_int]
        L.oracle_translational_difference_deg.restype = ctypes.c_double
        L.oracle_cost_function.argtypes = [ctypes.c_int64, dp, dp, dp, dp]
        L.oracle_cost_function.restype = ctypes.c_double
        L.oracle_energy.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, dp, dp,
                                    ctypes.c_double, dp, dp]
        L.oracle_energy.restype = ctypes.c_double
        L.oracle_unscented_transform_batch.argtypes = [ctypes.c_int64, dp, dp, dp, ctypes.c_double,
                                                       ctypes.c_int, dp]
        L.oracle_unscented_transform_batch.restype = None
        L.oracle_keypoints_unproject_batch.argtypes = [ctypes.c_int64, dp, dp, dp, dp, dp]
        L.oracle_keypoints_unproject_batch.restype = None
        L.oracle_fibonacci_sphere.argtypes = [ctypes.c_int, dp]
        L.oracle_fibonacci_sphere.restype = None
        L.oracle_scf_objective.argtypes = [ctypes.c_int64, dp, dp, dp]
        L.oracle_scf_objective.restype = ctypes.c_double
        L.oracle_scf_translation.argtypes = [ctypes.c_int64, dp, dp, dp, dp, ctypes.c_double, ctypes.c_int,
                                             ctypes.c_int, dp, dp]
        L.oracle_scf_translation.restype = ctypes.c_int
        L.oracle_nec_translation.argtypes = [ctypes.c_int64, dp, dp, dp, dp, dp]
        L.oracle_nec_translation.restype = None
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_es_smallest_ev.argtypes = [ctypes.c_int64, dp, dp, dp, dp, dp, dp]
        L.oracle_es_smallest_ev.restype = ctypes.c_double
        L.oracle_eigensolver.argtypes = [ctypes.c_int64, dp, dp, dp, dp, dp, ctypes.POINTER(OracleEsInfo)]
        L.oracle_eigensolver.restype = ctypes.c_int
        L.oracle_es_lm.argtypes = [ctypes.c_int64, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp, ip, ip]
        L.oracle_es_lm.restype = ctypes.c_int
        L.oracle_nec_eigensolver_pose.argtypes = [ctypes.c_int64, dp, dp, dp, dp, ctypes.POINTER(OracleEsInfo)]
        L.oracle_nec_eigensolver_pose.restype = ctypes.c_int
        L.oracle_weights.argtypes = [ctypes.c_int64, dp, dp, dp, ctypes.c_double, dp]
        L.oracle_weights.restype = None
        L.oracle_weighted_eigensolver.argtypes = [ctypes.c_int64, dp, dp, dp, dp, ctypes.c_double, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, dp]
        L.oracle_weighted_eigensolver.restype = ctypes.c_int
        L.oracle_frame_opts_default.argtypes = [ctypes.POINTER(OracleFrameOpts)]
        L.oracle_frame_opts_default.restype = None
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_frame_solve.argtypes = [ctypes.POINTER(OracleFrameOpts), ctypes.c_int64, ctypes.c_int64, dp, dp, dp,
                                         dp, dp, dp, u8p, ip, ip]
        L.oracle_frame_solve.restype = ctypes.c_int
        L.oracle_frame_solve_batch.argtypes = [ctypes.POINTER(OracleFrameOpts), ctypes.c_int64, ctypes.c_int64,
                                               ctypes.POINTER(ctypes.c_int64), dp, dp, dp, dp, dp, dp, ctypes.c_int,
                                               ctypes.c_int64, u8p, ip, ip]
        L.oracle_frame_solve_batch.restype = ctypes.c_int
        L.oracle_ransac_compute_model.argtypes = [ctypes.POINTER(OracleRansacOpts), ctypes.c_int64, ctypes.c_int64,
                                                  dp, dp, dp, dp, u8p, ip, ip]
        L.oracle_ransac_compute_model.restype = ctypes.c_int
        L.oracle_ransac_score.argtypes = [dp, dp, dp]
        L.oracle_ransac_score.restype = ctypes.c_double
        L.oracle_ransac_opts_default.argtypes = [ctypes.POINTER(OracleRansacOpts)]
        L.oracle_ransac_opts_default.restype = None
        L.oracle_ransac_eigensolver.argtypes = [ctypes.POINTER(OracleRansacOpts), ctypes.c_int64, ctypes.c_int64,
                                                dp, dp, dp, dp, ctypes.POINTER(ctypes.c_uint8), ip, ip]
        L.oracle_ransac_eigensolver.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a, shape_tail=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_tail is not None:
        a = a.reshape((-1,) + tuple(shape_tail))
    return a


def default_opts(variant: int = TARGET, regularization: float = 1e-13,
                 jacobian_mode: int = JAC_NUMERIC_CENTRAL, **overrides) -> OracleOpts:
    o = OracleOpts()
    lib().oracle_opts_default(ctypes.byref(o))
    o.variant = variant
    o.regularization = regularization
    o.jacobian_mode = jacobian_mode
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


@dataclass
class EvalResult:
    cost: float
    gradient: np.ndarray  # (5,)
    jtj: np.ndarray  # (15,) upper triangle, row-major packed


def covs_to_abi(covs):
    """(n,3,3) row-indexed numpy matrices -> the ABI's column-major double[n][9]."""
    if covs is None:
        return None
    covs = np.asarray(covs, dtype=np.float64).reshape(-1, 3, 3)
    return np.ascontiguousarray(covs.transpose(0, 2, 1)).reshape(-1, 9)


def evaluate(variant, f1, f2, cov_t, cov_h, reg, pose7, jacobian_mode=JAC_NUMERIC_CENTRAL):
    """cov_* are in the ABI layout: [n][9] column-major."""
    f1, f2 = _c(f1, (3,)), _c(f2, (3,))
    cov_t, cov_h = _c(cov_t, (9,)), _c(cov_h, (9,))
    pose7 = _c(pose7)
    cost = ctypes.c_double()
    g = np.zeros(5)
    H = np.zeros(15)
    rc = lib().oracle_eval(variant, jacobian_mode, f1.shape[0], _dp(f1), _dp(f2), _dp(cov_t),
                           _dp(cov_h), reg, _dp(pose7), ctypes.byref(cost), _dp(g), _dp(H))
    if rc != 0:
        raise RuntimeError("oracle_eval failed")
    return EvalResult(cost.value, g, H)


def solve(f1, f2, cov_t, cov_h, init_pose7, opts: OracleOpts):
    f1, f2 = _c(f1, (3,)), _c(f2, (3,))
    cov_t, cov_h = _c(cov_t, (9,)), _c(cov_h, (9,))
    init_pose7 = _c(init_pose7)
    out = np.zeros(7)
    info = OracleInfo()
    rc = lib().oracle_solve(ctypes.byref(opts), f1.shape[0], _dp(f1), _dp(f2), _dp(cov_t),
                            _dp(cov_h), _dp(init_pose7), _dp(out), ctypes.byref(info))
    if rc != 0:
        raise RuntimeError("oracle_solve failed")
    return out, info


def solve_batch(f1, f2, cov_t, cov_h, init_poses, opts: OracleOpts, offsets=None,
                n_per_problem=None, num_threads: int = 1):
    """Returns (poses [B,7], info structured array [B])."""
    f1, f2 = _c(f1, (3,)), _c(f2, (3,))
    cov_t, cov_h = _c(cov_t, (9,)), _c(cov_h, (9,))
    init_poses = _c(init_poses, (7,))
    B = init_poses.shape[0]
    if offsets is not None:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        assert offsets.shape[0] == B + 1
        op = offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        n_per_problem = 0
    else:
        op = None
        if n_per_problem is None:
            n_per_problem = f1.shape[0] // max(B, 1)
    out = np.zeros((B, 7))
    infos = np.zeros(B, dtype=INFO_DTYPE)
    assert INFO_DTYPE.itemsize == ctypes.sizeof(OracleInfo)
    rc = lib().oracle_solve_batch(ctypes.byref(opts), B, n_per_problem, op, _dp(f1), _dp(f2),
                                  _dp(cov_t), _dp(cov_h), _dp(init_poses), _dp(out),
                                  infos.ctypes.data_as(ctypes.c_void_p), num_threads)
    if rc != 0:
        raise RuntimeError("oracle_solve_batch failed")
    return out, infos


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def angles_from_vec(v):
    v = _c(v)
    th, ph = ctypes.c_double(), ctypes.c_double()
    lib().oracle_angles_from_vec(_dp(v), ctypes.byref(th), ctypes.byref(ph))
    return th.value, ph.value


def rotational_difference_deg(pose_a, pose_b) -> float:
    a, b = _c(pose_a), _c(pose_b)
    return float(lib().oracle_rotational_difference_deg(_dp(a), _dp(b)))


def translational_difference_deg(t1, t2, both_directions=True) -> float:
    a, b = _c(t1), _c(t2)
    return float(lib().oracle_translational_difference_deg(_dp(a), _dp(b), int(both_directions)))


def cost_function(f1, f2, cov, pose7) -> float:
    f1, f2, cov, pose7 = _c(f1, (3,)), _c(f2, (3,)), _c(cov, (9,)), _c(pose7)
    return float(lib().oracle_cost_function(f1.shape[0], _dp(f1), _dp(f2), _dp(cov), _dp(pose7)))


def energy(variant, f1, f2, cov_t, cov_h, reg, q_xyzw, t) -> float:
    f1, f2 = _c(f1, (3,)), _c(f2, (3,))
    cov_t, cov_h = _c(cov_t, (9,)), _c(cov_h, (9,))
    q, t = _c(q_xyzw), _c(t)
    return float(lib().oracle_energy(variant, f1.shape[0], _dp(f1), _dp(f2), _dp(cov_t),
                                     _dp(cov_h), reg, _dp(q), _dp(t)))


OMNIDIRECTIONAL, PINHOLE = 0, 1


def unscented_transform(mus, covs, K_inv=None, kappa=1.0, camera_model=PINHOLE):
    """mus (n,3); covs (n,9) column-major; K_inv (9,) column-major -> (n,9) column-major."""
    mus, covs = _c(mus, (3,)), _c(covs, (9,))
    K = _c(np.eye(3).reshape(9) if K_inv is None else K_inv)
    out = np.zeros((mus.shape[0], 9))
    lib().oracle_unscented_transform_batch(mus.shape[0], _dp(mus), _dp(covs), _dp(K), float(kappa),
                                           int(camera_model), _dp(out))
    return out


def keypoints_unproject(points, covs2, K_inv):
    """KeyPoint::Unproject for n keypoints -> (bvs (n,3), covs (n,9) column-major)."""
    points, covs2, K = _c(points, (2,)), _c(covs2, (4,)), _c(K_inv)
    bvs, covs = np.zeros((points.shape[0], 3)), np.zeros((points.shape[0], 9))
    lib().oracle_keypoints_unproject_batch(points.shape[0], _dp(points), _dp(covs2), _dp(K), _dp(bvs), _dp(covs))
    return bvs, covs


def fibonacci_sphere(samples: int) -> np.ndarray:
    """pnec::optimization::fibonacci_sphere (scf.cc:53-72) -> (samples, 3)."""
    pts = np.zeros((samples, 3))
    lib().oracle_fibonacci_sphere(int(samples), _dp(pts))
    return pts


def scf_objective(A, B, t) -> float:
    """pnec::optimization::obj_fun (scf.cc:43-51); A, B (n,3,3) row-indexed."""
    A, B, t = _c(A, (3, 3)), _c(B, (3, 3)), _c(t)
    return float(lib().oracle_scf_objective(A.shape[0], _dp(A), _dp(B), _dp(t)))


def scf_translation(f1, f2, cov, pose7, reg=1e-13, samples=500, steps=10):
    """SCF stage of PNEC::WeightedEigensolver for one pair -> (t (3,), objective)."""
    f1, f2, cov, pose7 = _c(f1, (3,)), _c(f2, (3,)), _c(cov, (9,)), _c(pose7)
    t = np.zeros(3)
    cost = ctypes.c_double()
    rc = lib().oracle_scf_translation(f1.shape[0], _dp(f1), _dp(f2), _dp(cov), _dp(pose7), float(reg),
                                      int(samples), int(steps), _dp(t), ctypes.byref(cost))
    if rc != 0:
        raise RuntimeError("oracle_scf_translation failed")
    return t, cost.value


def nec_translation(f1, f2, pose7):
    """TranslationFromM(ComposeM(...)) (common.cc:127-181) -> (t (3,), M6 (6,))."""
    f1, f2, pose7 = _c(f1, (3,)), _c(f2, (3,)), _c(pose7)
    t, M = np.zeros(3), np.zeros(6)
    lib().oracle_nec_translation(f1.shape[0], _dp(f1), _dp(f2), _dp(pose7), _dp(t), _dp(M))
    return t, M


# ----------------------------------------------------------------- front stages of PNEC::Solve


def es_smallest_ev(f1, f2, cayley, weights=None):
    """lambda_min of opengv's reduced M(c), its gradient (3,) and M (3,3) at a Cayley vector."""
    f1, f2, c, w = _c(f1, (3,)), _c(f2, (3,)), _c(cayley), _c(weights)
    jac, M = np.zeros(3), np.zeros(9)
    ev = lib().oracle_es_smallest_ev(f1.shape[0], _dp(f1), _dp(f2), _dp(w), _dp(c), _dp(jac), _dp(M))
    return float(ev), jac, M.reshape(3, 3)


def eigensolver(f1, f2, init_pose7, weights=None):
    """opengv::relative_pose::eigensolver restated -> (unit quaternion xyzw (4,), OracleEsInfo)."""
    f1, f2, p, w = _c(f1, (3,)), _c(f2, (3,)), _c(init_pose7), _c(weights)
    q = np.zeros(4)
    info = OracleEsInfo()
    lib().oracle_eigensolver(f1.shape[0], _dp(f1), _dp(f2), _dp(w), _dp(p), _dp(q), ctypes.byref(info))
    return q, info


def es_lm(f1, f2, x0, weights=None, ftol=5e-5, xtol=10 * np.finfo(float).eps, gtol=0.0, factor=100.0,
          maxfev=100, fev_per_jacobian=4):
    """The restated MINPACK lmdif alone on the eigensolver's residuals -> (x (3,), info, nfev)."""
    f1, f2, x0, w = _c(f1, (3,)), _c(f2, (3,)), _c(x0), _c(weights)
    x = np.zeros(3)
    info, nfev = ctypes.c_int32(), ctypes.c_int32()
    lib().oracle_es_lm(f1.shape[0], _dp(f1), _dp(f2), _dp(w), _dp(x0), ftol, xtol, gtol, factor, int(maxfev),
                       int(fev_per_jacobian), _dp(x), ctypes.byref(info), ctypes.byref(nfev))
    return x, info.value, nfev.value


def nec_eigensolver_pose(f1, f2, init_pose7):
    """PNEC::Eigensolver without RANSAC (pnec.cc:273-279) -> pose7."""
    f1, f2, p = _c(f1, (3,)), _c(f2, (3,)), _c(init_pose7)
    out = np.zeros(7)
    info = OracleEsInfo()
    lib().oracle_nec_eigensolver_pose(f1.shape[0], _dp(f1), _dp(f2), _dp(p), _dp(out), ctypes.byref(info))
    return out, info


def weights(f1, cov, pose7, reg=1e-13):
    """pnec::common::Weight(host_frame=false) * 1e-8 (common.cc:183-208, pnec.cc:296-300)."""
    f1, cov, p = _c(f1, (3,)), _c(cov, (9,)), _c(pose7)
    w = np.zeros(f1.shape[0])
    lib().oracle_weights(f1.shape[0], _dp(f1), _dp(cov), _dp(p), float(reg), _dp(w))
    return w


def weighted_eigensolver(f1, f2, cov, initial_pose7, reg=1e-13, weighted_iterations=10, samples=500, steps=10):
    """PNEC::WeightedEigensolver (pnec.cc:283-348) -> pose7."""
    f1, f2, cov, p = _c(f1, (3,)), _c(f2, (3,)), _c(cov, (9,)), _c(initial_pose7)
    out = np.zeros(7)
    rc = lib().oracle_weighted_eigensolver(f1.shape[0], _dp(f1), _dp(f2), _dp(cov), _dp(p), float(reg),
                                           int(weighted_iterations), int(samples), int(steps), _dp(out))
    if rc != 0:
        raise RuntimeError("oracle_weighted_eigensolver failed")
    return out


def default_frame_opts(**overrides) -> OracleFrameOpts:
    """Options() defaults (use_ransac = 1, pnec_config.h:58).  Keys: the fields of OracleFrameOpts, of its
    ceres options, and the RANSAC settings under their Options / C-ABI names (max_ransac_iterations,
    ransac_sample_size, ransac_seed, ransac_threshold, ransac_probability, ransac_max_variation,
    ransac_sequential)."""
    o = OracleFrameOpts()
    lib().oracle_frame_opts_default(ctypes.byref(o))
    for k, v in overrides.items():
        if k in _RANSAC_NAMES:
            setattr(o.ransac, _RANSAC_NAMES[k], v)
        elif hasattr(o, k):
            setattr(o, k, v)
        elif hasattr(o.ceres, k):
            setattr(o.ceres, k, v)
        else:
            raise AttributeError(k)
    return o


def frame_solve_batch(f1, f2, cov, init_poses, opts: OracleFrameOpts, offsets=None, n_per_problem=None,
                      num_threads: int = 1, pair_index_base: int = 0, return_ransac: bool = False):
    """PNEC::Solve per frame pair -> (poses [B,7], eigensolver poses [B,7]); with return_ransac also
    (inlier mask [total] bool, num_inliers [B], ransac iterations [B])."""
    f1, f2 = _c(f1, (3,)), _c(f2, (3,))
    cov = None if cov is None else _c(cov, (9,))
    init_poses = _c(init_poses, (7,))
    B = init_poses.shape[0]
    if offsets is not None:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        op = offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        n_per_problem = 0
    else:
        op = None
        if n_per_problem is None:
            n_per_problem = f1.shape[0] // max(B, 1)
    out, es = np.zeros((B, 7)), np.zeros((B, 7))
    mask = np.zeros(max(f1.shape[0], 1), dtype=np.uint8)
    ni, it = np.zeros(max(B, 1), dtype=np.int32), np.zeros(max(B, 1), dtype=np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib().oracle_frame_solve_batch(ctypes.byref(opts), B, n_per_problem, op, _dp(f1), _dp(f2), _dp(cov),
                                        _dp(init_poses), _dp(out), _dp(es), num_threads, int(pair_index_base),
                                        mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                        ni.ctypes.data_as(i32p), it.ctypes.data_as(i32p))
    if rc != 0:
        raise RuntimeError("oracle_frame_solve_batch failed")
    if return_ransac:
        return out, es, mask[:f1.shape[0]].astype(bool), ni[:B], it[:B]
    return out, es


def _ransac_opts(overrides) -> OracleRansacOpts:
    o = OracleRansacOpts()
    lib().oracle_ransac_opts_default(ctypes.byref(o))
    for k, v in overrides.items():
        k = _RANSAC_NAMES.get(k, k)
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def ransac_compute_model(f1, f2, init_pose7, pair_index=0, **overrides):
    """opengv::sac::Ransac<EigensolverSacProblem>::computeModel + selectWithinDistance, restated
    -> (best model pose7, inlier mask (n,) bool, iterations)."""
    f1, f2, p = _c(f1, (3,)), _c(f2, (3,)), _c(init_pose7)
    o = _ransac_opts(overrides)
    out = np.zeros(7)
    mask = np.zeros(max(f1.shape[0], 1), dtype=np.uint8)
    ni, it = ctypes.c_int32(), ctypes.c_int32()
    rc = lib().oracle_ransac_compute_model(ctypes.byref(o), int(pair_index), f1.shape[0], _dp(f1), _dp(f2), _dp(p),
                                           _dp(out), mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                           ctypes.byref(ni), ctypes.byref(it))
    if rc not in (0, -2):
        raise RuntimeError(f"oracle_ransac_compute_model failed ({rc})")
    return out, mask[:f1.shape[0]].astype(bool), it.value


def ransac_score(pose7, f1, f2) -> float:
    """EigensolverSacProblem's reprojection score of one correspondence under a model."""
    return float(lib().oracle_ransac_score(_dp(_c(pose7)), _dp(_c(f1)), _dp(_c(f2))))


def ransac_eigensolver(f1, f2, init_pose7, pair_index=0, **overrides):
    """PNEC::Eigensolver with use_ransac_ (pnec.cc:239-272), restated
    -> (pose7, inlier mask (n,) bool, iterations)."""
    f1, f2, p = _c(f1, (3,)), _c(f2, (3,)), _c(init_pose7)
    o = _ransac_opts(overrides)
    out = np.zeros(7)
    mask = np.zeros(max(f1.shape[0], 1), dtype=np.uint8)
    ni, it = ctypes.c_int32(), ctypes.c_int32()
    rc = lib().oracle_ransac_eigensolver(ctypes.byref(o), int(pair_index), f1.shape[0], _dp(f1), _dp(f2), _dp(p),
                                         _dp(out), mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                         ctypes.byref(ni), ctypes.byref(it))
    if rc != 0:
        raise RuntimeError(f"oracle_ransac_eigensolver failed ({rc})")
    mask = mask[:f1.shape[0]].astype(bool)
    assert int(mask.sum()) == ni.value
    return out, mask, it.value
