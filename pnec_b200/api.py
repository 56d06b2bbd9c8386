"""Host-side binding of the C-ABI (include/pnec_b200.h) — ctypes, no torch types
cross the boundary: tensors are passed as raw device pointers + the current CUDA
stream handle.  There is NO CPU fallback: if libpnec_b200.so is missing or no
Blackwell GPU is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNEC_B200_LIB") or os.path.join(_HERE, "lib", "libpnec_b200.so")

NEC, TARGET, HOST, SYMMETRIC = 0, 1, 2, 3
CAMERA_OMNIDIRECTIONAL, CAMERA_PINHOLE = 0, 1
VARIANT_NAMES = {"nec": NEC, "target": TARGET, "host": HOST, "symmetric": SYMMETRIC}
MEM_HOST, MEM_DEVICE = 0, 1

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

# every symbol include/pnec_b200.h declares
EXPORTED_SYMBOLS = (
    "pnec_version",
    "pnec_last_error",
    "pnec_status_string",
    "pnec_solver_opts_default",
    "pnec_create",
    "pnec_destroy",
    "pnec_solve_batch",
    "pnec_eval_batch",
    "pnec_cost_function_batch",
    "pnec_unscented_transform_batch",
    "pnec_keypoints_unproject_batch",
    "pnec_scf_translation_batch",
    "pnec_ransac_batch",
    "pnec_keypoints_to_batch",
    "pnec_solve_from_keypoints_batch",
    "pnec_frame_solve_from_keypoints_batch",
    "pnec_nec_translation_batch",
    "pnec_eigensolver_batch",
    "pnec_frame_opts_default",
    "pnec_frame_solve_batch",
    "pnec_launch_count",
)


class PnecError(RuntimeError):
    pass


class SolverOpts(ctypes.Structure):
    """pnec_solver_opts == ceres::Solver::Options defaults the reference runs with
    (src/optimization/pnec_ceres.cc:43-48) + Options::regularization_
    (include/rel_pose_estimation/pnec_config.h:50)."""

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
    ]


class _Batch(ctypes.Structure):
    _fields_ = [
        ("num_problems", ctypes.c_int64),
        ("n_per_problem", ctypes.c_int64),
        ("offsets", ctypes.c_void_p),
        ("memspace", ctypes.c_int32),
        ("cov_layout", ctypes.c_int32),
        ("bvs_host", ctypes.c_void_p),
        ("bvs_target", ctypes.c_void_p),
        ("covs_target", ctypes.c_void_p),
        ("covs_host", ctypes.c_void_p),
        ("poses", ctypes.c_void_p),
    ]


class _SolveOut(ctypes.Structure):
    _fields_ = [
        ("poses", ctypes.c_void_p),
        ("status", ctypes.c_void_p),
        ("iterations", ctypes.c_void_p),
        ("cost", ctypes.c_void_p),
        ("initial_cost", ctypes.c_void_p),
    ]


class _EvalOut(ctypes.Structure):
    _fields_ = [
        ("cost", ctypes.c_void_p),
        ("gradient", ctypes.c_void_p),
        ("jtj", ctypes.c_void_p),
    ]


class FrameOpts(ctypes.Structure):
    """pnec_frame_opts == pnec::rel_pose_estimation::Options as PNEC::Solve reads it
    (include/rel_pose_estimation/pnec_config.h:46-65)."""

    _fields_ = [
        ("use_nec", ctypes.c_int32),
        ("use_ceres", ctypes.c_int32),
        ("weighted_iterations", ctypes.c_int32),
        ("use_ransac", ctypes.c_int32),
        ("fibonacci_samples", ctypes.c_int32),
        ("scf_steps", ctypes.c_int32),
        ("ceres", SolverOpts),
        ("max_ransac_iterations", ctypes.c_int32),
        ("ransac_sample_size", ctypes.c_int32),
        ("ransac_threshold", ctypes.c_double),
        ("ransac_probability", ctypes.c_double),
        ("ransac_max_variation", ctypes.c_double),
        ("ransac_seed", ctypes.c_uint64),
        ("ransac_pair_index_base", ctypes.c_int64),
    ]


class _FrameOut(ctypes.Structure):
    _fields_ = [
        ("poses", ctypes.c_void_p),
        ("es_poses", ctypes.c_void_p),
        ("status", ctypes.c_void_p),
        ("iterations", ctypes.c_void_p),
        ("cost", ctypes.c_void_p),
        ("num_inliers", ctypes.c_void_p),
        ("inlier_index", ctypes.c_void_p),
        ("ransac_iterations", ctypes.c_void_p),
        ("stage_ms", ctypes.c_void_p),
    ]


class _KeypointBatch(ctypes.Structure):
    _fields_ = [
        ("num_problems", ctypes.c_int64),
        ("n_per_problem", ctypes.c_int64),
        ("offsets", ctypes.c_void_p),
        ("memspace", ctypes.c_int32),
        ("packed_covs", ctypes.c_int32),
        ("num_host_keypoints", ctypes.c_int64),
        ("num_target_keypoints", ctypes.c_int64),
        ("host_points", ctypes.c_void_p),
        ("target_points", ctypes.c_void_p),
        ("host_covs2", ctypes.c_void_p),
        ("target_covs2", ctypes.c_void_p),
        ("host_index", ctypes.c_void_p),
        ("target_index", ctypes.c_void_p),
        ("K_inv", ctypes.c_void_p),
        ("poses", ctypes.c_void_p),
    ]


_lib = None


def load_library() -> ctypes.CDLL:
    """Loads libpnec_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnecError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  pnec_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.pnec_version.restype = ctypes.c_int
    L.pnec_last_error.restype = ctypes.c_char_p
    L.pnec_status_string.restype = ctypes.c_char_p
    L.pnec_status_string.argtypes = [ctypes.c_int32]
    L.pnec_solver_opts_default.argtypes = [ctypes.POINTER(SolverOpts)]
    L.pnec_solver_opts_default.restype = None
    L.pnec_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.pnec_create.restype = ctypes.c_int
    L.pnec_destroy.argtypes = [ctypes.c_void_p]
    L.pnec_destroy.restype = None
    L.pnec_solve_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch),
                                   ctypes.POINTER(SolverOpts), ctypes.POINTER(_SolveOut),
                                   ctypes.c_void_p]
    L.pnec_solve_batch.restype = ctypes.c_int
    L.pnec_eval_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_int32,
                                  ctypes.c_double, ctypes.POINTER(_EvalOut), ctypes.c_void_p]
    L.pnec_eval_batch.restype = ctypes.c_int
    L.pnec_cost_function_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch),
                                           ctypes.c_void_p, ctypes.c_void_p]
    L.pnec_cost_function_batch.restype = ctypes.c_int
    L.pnec_unscented_transform_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_double, ctypes.c_int32, ctypes.c_void_p,
                                                 ctypes.c_void_p]
    L.pnec_unscented_transform_batch.restype = ctypes.c_int
    L.pnec_keypoints_unproject_batch.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.pnec_keypoints_unproject_batch.restype = ctypes.c_int
    L.pnec_scf_translation_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_double,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]
    L.pnec_scf_translation_batch.restype = ctypes.c_int
    L.pnec_nec_translation_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]
    L.pnec_nec_translation_batch.restype = ctypes.c_int
    L.pnec_eigensolver_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_void_p,
                                         ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
    L.pnec_eigensolver_batch.restype = ctypes.c_int
    L.pnec_frame_opts_default.argtypes = [ctypes.POINTER(FrameOpts)]
    L.pnec_frame_opts_default.restype = None
    L.pnec_frame_solve_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.POINTER(FrameOpts),
                                         ctypes.POINTER(_FrameOut), ctypes.c_void_p]
    L.pnec_frame_solve_batch.restype = ctypes.c_int
    L.pnec_ransac_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.POINTER(FrameOpts), ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
    L.pnec_ransac_batch.restype = ctypes.c_int
    L.pnec_keypoints_to_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_KeypointBatch), ctypes.POINTER(_Batch),
                                          ctypes.c_void_p]
    L.pnec_keypoints_to_batch.restype = ctypes.c_int
    L.pnec_solve_from_keypoints_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_KeypointBatch),
                                                  ctypes.POINTER(SolverOpts), ctypes.POINTER(_SolveOut), ctypes.c_void_p]
    L.pnec_solve_from_keypoints_batch.restype = ctypes.c_int
    L.pnec_frame_solve_from_keypoints_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_KeypointBatch),
                                                        ctypes.POINTER(FrameOpts), ctypes.POINTER(_FrameOut),
                                                        ctypes.c_void_p]
    L.pnec_frame_solve_from_keypoints_batch.restype = ctypes.c_int
    L.pnec_launch_count.argtypes = [ctypes.c_void_p]
    L.pnec_launch_count.restype = ctypes.c_int64
    _lib = L
    return L


def default_opts(variant: int = TARGET, regularization: float = 1e-13, **overrides) -> SolverOpts:
    o = SolverOpts()
    load_library().pnec_solver_opts_default(ctypes.byref(o))
    o.variant = int(variant)
    o.regularization = float(regularization)
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(f"pnec_solver_opts has no field {k!r}")
        setattr(o, k, v)
    return o


def default_frame_opts(**overrides) -> FrameOpts:
    """pnec_frame_opts_default + overrides; unknown names are looked up in the nested `ceres` options."""
    o = FrameOpts()
    load_library().pnec_frame_opts_default(ctypes.byref(o))
    for k, v in overrides.items():
        if k != "ceres" and hasattr(o, k):
            setattr(o, k, v)
        elif hasattr(o.ceres, k):
            setattr(o.ceres, k, v)
        else:
            raise AttributeError(f"pnec_frame_opts has no field {k!r}")
    return o


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


@dataclass
class SolveResult:
    poses: object  # (B,7) numpy (host call) or torch.cuda tensor (device call)
    status: object  # (B,) int32
    iterations: object  # (B,) int32
    cost: object  # (B,) final 1/2 sum r^2
    initial_cost: object  # (B,)


@dataclass
class FrameResult:
    poses: object  # (B,7) result of PNEC::Solve
    es_poses: object  # (B,7) result of PNEC::Eigensolver (ES_solution)
    status: object  # (B,) int32, refinement status
    iterations: object  # (B,) int32
    cost: object  # (B,)
    num_inliers: object = None  # (B,) int32: size of PNEC::Solve's `inliers` (0 without RANSAC)
    ransac_iterations: object = None  # (B,) int32
    inlier_index: object = None  # (total,) int32: pair b's ascending inlier indices start at its offset
    stage_ms: object = None  # (3,) float32: nec_es, it_es, ceres (FrameTiming), when asked for


@dataclass
class EvalResult:
    cost: object  # (B,)
    gradient: object  # (B,5)
    jtj: object  # (B,15) packed upper triangle


class Handle:
    """pnec_handle wrapper.  One per device/stream user; re-entrant per handle."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        rc = self._lib.pnec_create(int(device), ctypes.byref(self._h))
        if rc != 0:
            raise PnecError(f"pnec_create failed ({rc}): {self._lib.pnec_last_error().decode()}")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.pnec_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.pnec_launch_count(self._h))

    # ------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise PnecError(f"{what} failed ({rc}): {self._lib.pnec_last_error().decode()}")

    @staticmethod
    def _prep_host(a, tail):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a.reshape((-1,) + tail)

    def _batch(self, f1, f2, ct, ch, poses, offsets, n_per_problem, keep):
        device = _is_torch(f1)
        b = _Batch()
        if device:
            import torch

            def dev(t, tail):
                if t is None:
                    return None
                if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                    raise PnecError("device batches need contiguous float64 CUDA tensors")
                if t.device.index != self.device:
                    raise PnecError("tensor is on a different device than the handle")
                return t

            packed = ct is not None and ct.dim() == 2 and ct.shape[1] == 6
            f1, f2, ct, ch, poses = (dev(x, None) for x in (f1, f2, ct, ch, poses))
            ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
            total = f1.numel() // 3
            B = poses.numel() // 7
            b.memspace = MEM_DEVICE
        else:
            f1 = self._prep_host(f1, (3,))
            f2 = self._prep_host(f2, (3,))
            # covariances: (total, 9) / (total, 3, 3) column-major 3x3, or (total, 6) packed symmetric
            packed = ct is not None and np.asarray(ct).ndim == 2 and np.asarray(ct).shape[1] == 6
            ct = self._prep_host(ct, (6,) if packed else (9,))
            ch = self._prep_host(ch, (6,) if packed else (9,))
            poses = self._prep_host(poses, (7,))
            ptr = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)
            total = f1.shape[0]
            B = poses.shape[0]
            b.memspace = MEM_HOST
        keep.extend([f1, f2, ct, ch, poses])
        size = (lambda t: t.numel()) if device else (lambda a: a.size)
        if f2 is not None and size(f2) != total * 3:
            raise PnecError("bvs_host and bvs_target differ in size")
        # the C side cannot see array lengths: an undersized covariance or pose array would be read
        # out of bounds there
        per = 6 if packed else 9
        for name, arr in (("covs_target", ct), ("covs_host", ch)):
            if arr is not None and size(arr) != total * per:
                raise PnecError(f"{name} must hold {per} doubles per correspondence ({total * per}), got {size(arr)}")
        b.cov_layout = 1 if packed else 0
        if poses is not None and size(poses) != B * 7:
            raise PnecError("poses must hold 7 doubles per frame pair")
        b.num_problems = B
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            if offsets.shape != (B + 1,):
                raise PnecError("offsets must have num_problems + 1 entries")
            if int(offsets[-1]) > total:
                raise PnecError("offsets run past the end of the correspondence arrays")
            keep.append(offsets)
            b.offsets = ctypes.c_void_p(offsets.ctypes.data)
            b.n_per_problem = 0
        else:
            if n_per_problem is None:
                n_per_problem = total // B if B else 0
            if n_per_problem * B > total:
                raise PnecError("n_per_problem * num_problems exceeds the correspondence arrays")
            b.offsets = None
            b.n_per_problem = int(n_per_problem)
        b.bvs_host, b.bvs_target = ptr(f1), ptr(f2)
        b.covs_target, b.covs_host, b.poses = ptr(ct), ptr(ch), ptr(poses)
        return b, device, B

    def _stream(self, device: bool):
        if not device:
            return None
        import torch

        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ----------------------------------------------------------------- API
    def solve_batch(self, bvs_host, bvs_target, covs_target, covs_host, init_poses,
                    opts: Optional[SolverOpts] = None, *, offsets=None, n_per_problem=None,
                    out: Optional[SolveResult] = None) -> SolveResult:
        """Batched on-device LM refinement (pnec_solve_batch).

        numpy inputs -> HOST call (copies inside, synchronous); torch CUDA tensors ->
        DEVICE call, enqueued on torch's current stream, results are CUDA tensors.
        """
        opts = opts or default_opts()
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, covs_host, init_poses,
                                   offsets, n_per_problem, keep)
        o = _SolveOut()
        if device:
            import torch

            dev = torch.device("cuda", self.device)
            if out is None:
                out = SolveResult(
                    torch.empty((B, 7), dtype=torch.float64, device=dev),
                    torch.empty((B,), dtype=torch.int32, device=dev),
                    torch.empty((B,), dtype=torch.int32, device=dev),
                    torch.empty((B,), dtype=torch.float64, device=dev),
                    torch.empty((B,), dtype=torch.float64, device=dev),
                )
            o.poses, o.status = out.poses.data_ptr(), out.status.data_ptr()
            o.iterations, o.cost = out.iterations.data_ptr(), out.cost.data_ptr()
            o.initial_cost = out.initial_cost.data_ptr()
        else:
            if out is None:
                out = SolveResult(np.empty((B, 7)), np.empty(B, np.int32), np.empty(B, np.int32),
                                  np.empty(B), np.empty(B))
            o.poses, o.status = out.poses.ctypes.data, out.status.ctypes.data
            o.iterations, o.cost = out.iterations.ctypes.data, out.cost.ctypes.data
            o.initial_cost = out.initial_cost.ctypes.data
        rc = self._lib.pnec_solve_batch(self._h, ctypes.byref(b), ctypes.byref(opts),
                                        ctypes.byref(o), self._stream(device))
        self._check(rc, "pnec_solve_batch")
        return out

    def eval_batch(self, bvs_host, bvs_target, covs_target, covs_host, poses, variant=TARGET,
                   regularization=1e-13, *, offsets=None, n_per_problem=None,
                   out: Optional[EvalResult] = None) -> EvalResult:
        """One fused residual + Jacobian + JtJ/Jtr/cost pass per problem (pnec_eval_batch)."""
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, covs_host, poses, offsets,
                                   n_per_problem, keep)
        o = _EvalOut()
        if device:
            import torch

            dev = torch.device("cuda", self.device)
            if out is None:
                out = EvalResult(torch.empty((B,), dtype=torch.float64, device=dev),
                                 torch.empty((B, 5), dtype=torch.float64, device=dev),
                                 torch.empty((B, 15), dtype=torch.float64, device=dev))
            o.cost, o.gradient, o.jtj = out.cost.data_ptr(), out.gradient.data_ptr(), out.jtj.data_ptr()
        else:
            if out is None:
                out = EvalResult(np.empty(B), np.empty((B, 5)), np.empty((B, 15)))
            o.cost, o.gradient, o.jtj = out.cost.ctypes.data, out.gradient.ctypes.data, out.jtj.ctypes.data
        rc = self._lib.pnec_eval_batch(self._h, ctypes.byref(b), int(variant), float(regularization),
                                       ctypes.byref(o), self._stream(device))
        self._check(rc, "pnec_eval_batch")
        return out

    def cost_function_batch(self, bvs_host, bvs_target, covs_target, poses, *, offsets=None,
                            n_per_problem=None):
        """pnec::common::CostFunction (src/common/common.cc:237-259) per problem."""
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, None, poses, offsets,
                                   n_per_problem, keep)
        if device:
            import torch

            out = torch.empty((B,), dtype=torch.float64, device=torch.device("cuda", self.device))
            p = ctypes.c_void_p(out.data_ptr())
        else:
            out = np.empty(B)
            p = ctypes.c_void_p(out.ctypes.data)
        rc = self._lib.pnec_cost_function_batch(self._h, ctypes.byref(b), p, self._stream(device))
        self._check(rc, "pnec_cost_function_batch")
        return out


    def unscented_transform(self, mus, covs, K_inv=None, kappa=1.0, camera_model=CAMERA_PINHOLE):
        """pnec::common::UnscentedTransform (src/common/common.cc:467-550) for n points.
        mus (n,3), covs (n,9) column-major, K_inv (9,) column-major host array -> (n,9)."""
        device = _is_torch(mus)
        K = np.ascontiguousarray(np.eye(3).reshape(9) if K_inv is None else K_inv, dtype=np.float64).reshape(9)
        if device:
            import torch

            n = mus.numel() // 3
            if not (mus.is_cuda and covs.is_cuda and mus.is_contiguous() and covs.is_contiguous()
                    and mus.dtype == torch.float64 and covs.dtype == torch.float64):
                raise PnecError("device calls need contiguous float64 CUDA tensors")
            out = torch.empty((n, 9), dtype=torch.float64, device=mus.device)
            pm, pc, po = mus.data_ptr(), covs.data_ptr(), out.data_ptr()
        else:
            mus = self._prep_host(mus, (3,))
            covs = self._prep_host(covs, (9,))
            n = mus.shape[0]
            out = np.empty((n, 9))
            pm, pc, po = mus.ctypes.data, covs.ctypes.data, out.ctypes.data
        ncov = (covs.numel() if device else covs.size) // 9
        if ncov != n:
            raise PnecError("mus and covs differ in length")
        rc = self._lib.pnec_unscented_transform_batch(
            self._h, n, MEM_DEVICE if device else MEM_HOST, ctypes.c_void_p(pm), ctypes.c_void_p(pc),
            ctypes.c_void_p(K.ctypes.data), float(kappa), int(camera_model), ctypes.c_void_p(po),
            self._stream(device))
        self._check(rc, "pnec_unscented_transform_batch")
        return out


    def keypoints_unproject(self, points, covs2, K_inv):
        """KeyPoint::Unproject (src/frames/keypoints.cc:49-62) for n keypoints: points (n,2),
        covs2 (n,4) column-major 2x2, K_inv (9,) column-major -> (bvs (n,3), covs (n,9))."""
        device = _is_torch(points)
        K = np.ascontiguousarray(K_inv, dtype=np.float64).reshape(9)
        if device:
            import torch

            n = points.numel() // 2
            if not (points.is_cuda and covs2.is_cuda and points.is_contiguous() and covs2.is_contiguous()
                    and points.dtype == torch.float64 and covs2.dtype == torch.float64):
                raise PnecError("device calls need contiguous float64 CUDA tensors")
            bvs = torch.empty((n, 3), dtype=torch.float64, device=points.device)
            covs = torch.empty((n, 9), dtype=torch.float64, device=points.device)
            ptrs = (points.data_ptr(), covs2.data_ptr(), bvs.data_ptr(), covs.data_ptr())
            ncov = covs2.numel() // 4
        else:
            points = self._prep_host(points, (2,))
            covs2 = self._prep_host(covs2, (4,))
            n = points.shape[0]
            bvs, covs = np.empty((n, 3)), np.empty((n, 9))
            ptrs = (points.ctypes.data, covs2.ctypes.data, bvs.ctypes.data, covs.ctypes.data)
            ncov = covs2.shape[0]
        if ncov != n:
            raise PnecError("points and covs2 differ in length")
        rc = self._lib.pnec_keypoints_unproject_batch(
            self._h, n, MEM_DEVICE if device else MEM_HOST, ctypes.c_void_p(ptrs[0]),
            ctypes.c_void_p(ptrs[1]), ctypes.c_void_p(K.ctypes.data), ctypes.c_void_p(ptrs[2]),
            ctypes.c_void_p(ptrs[3]), self._stream(device))
        self._check(rc, "pnec_keypoints_unproject_batch")
        return bvs, covs


    def _out(self, device, shape):
        if device:
            import torch

            t = torch.empty(shape, dtype=torch.float64, device=torch.device("cuda", self.device))
            return t, ctypes.c_void_p(t.data_ptr())
        a = np.empty(shape)
        return a, ctypes.c_void_p(a.ctypes.data)

    def scf_translation_batch(self, bvs_host, bvs_target, covs_target, poses, regularization=1e-13,
                              fibonacci_samples=500, scf_steps=10, *, offsets=None, n_per_problem=None):
        """Translation given rotation by Fibonacci scan + SCF (pnec.cc:317-343, scf.cc) ->
        (translations (B,3), objective (B,)).  `poses`: rotation quaternion + start translation."""
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, None, poses, offsets, n_per_problem, keep)
        t, pt = self._out(device, (B, 3))
        c, pc = self._out(device, (B,))
        rc = self._lib.pnec_scf_translation_batch(self._h, ctypes.byref(b), float(regularization),
                                                  int(fibonacci_samples), int(scf_steps), pt, pc,
                                                  self._stream(device))
        self._check(rc, "pnec_scf_translation_batch")
        return t, c

    def nec_translation_batch(self, bvs_host, bvs_target, poses, *, offsets=None, n_per_problem=None):
        """TranslationFromM(ComposeM(...)) (common.cc:127-181) -> (translations (B,3), M (B,6))."""
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, None, None, poses, offsets, n_per_problem, keep)
        t, pt = self._out(device, (B, 3))
        m, pm = self._out(device, (B, 6))
        rc = self._lib.pnec_nec_translation_batch(self._h, ctypes.byref(b), pt, pm, self._stream(device))
        self._check(rc, "pnec_nec_translation_batch")
        return t, m


    def _out_i32(self, device, shape):
        if device:
            import torch

            t = torch.zeros(shape, dtype=torch.int32, device=torch.device("cuda", self.device))
            return t, ctypes.c_void_p(t.data_ptr())
        a = np.zeros(shape, np.int32)
        return a, ctypes.c_void_p(a.ctypes.data)

    def eigensolver_batch(self, bvs_host, bvs_target, init_poses, *, covs_target=None, weight_poses=None,
                          regularization=1e-13, offsets=None, n_per_problem=None):
        """opengv::relative_pose::eigensolver for B frame pairs (pnec.cc:274 / :313) ->
        (poses (B,7): result rotation + the input translation, lm_info (B,), smallest_ev (B,)).
        `weight_poses` (B,7) switches on the weights of PNEC::WeightedEigensolver (needs covs_target)."""
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, None, init_poses, offsets, n_per_problem, keep)
        pw = None
        if weight_poses is not None:
            if device:
                if not (_is_torch(weight_poses) and weight_poses.is_cuda and weight_poses.is_contiguous()):
                    raise PnecError("weight_poses must be a contiguous CUDA tensor for device batches")
                pw = ctypes.c_void_p(weight_poses.data_ptr())
            else:
                weight_poses = self._prep_host(weight_poses, (7,))
                pw = ctypes.c_void_p(weight_poses.ctypes.data)
            keep.append(weight_poses)
        poses, pp = self._out(device, (B, 7))
        info, pi = self._out_i32(device, (B,))
        ev, pe = self._out(device, (B,))
        rc = self._lib.pnec_eigensolver_batch(self._h, ctypes.byref(b), pw, float(regularization), pp, pi, pe,
                                              self._stream(device))
        self._check(rc, "pnec_eigensolver_batch")
        return poses, info, ev

    def ransac_batch(self, bvs_host, bvs_target, init_poses, opts: Optional[FrameOpts] = None, *,
                     pair_index_base: int = 0, offsets=None, n_per_problem=None):
        """opengv's Ransac<EigensolverSacProblem>::computeModel + selectWithinDistance for B frame pairs
        (pnec.cc:239-251) -> (winning models (B,7), num_inliers (B,), iterations (B,), inlier_index (total,))."""
        opts = opts or default_frame_opts()
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, None, None, init_poses, offsets, n_per_problem, keep)
        total = int(offsets[-1]) if offsets is not None else B * int(b.n_per_problem)
        models, pm = self._out(device, (B, 7))
        ni, pn = self._out_i32(device, (B,))
        it, pi = self._out_i32(device, (B,))
        idx, px = self._out_i32(device, (max(total, 1),))
        rc = self._lib.pnec_ransac_batch(self._h, ctypes.byref(b), ctypes.byref(opts), int(pair_index_base), pm, pn,
                                         pi, px, self._stream(device))
        self._check(rc, "pnec_ransac_batch")
        return models, ni, it, idx[:total]

    def frame_solve_batch(self, bvs_host, bvs_target, covs_target, init_poses,
                          opts: Optional[FrameOpts] = None, *, offsets=None, n_per_problem=None,
                          stage_timing: bool = False) -> FrameResult:
        """PNEC::Solve (pnec.cc:77-124) for B frame pairs, every stage on the device.  With
        `stage_timing` the result carries stage_ms = (nec_es, it_es, ceres) in milliseconds
        (FrameTiming of the timed Solve overloads) and the call synchronises."""
        opts = opts or default_frame_opts()
        keep = []
        b, device, B = self._batch(bvs_host, bvs_target, covs_target, None, init_poses, offsets, n_per_problem, keep)
        total = int(offsets[-1]) if offsets is not None else B * int(b.n_per_problem)
        poses, pp = self._out(device, (B, 7))
        es, pes = self._out(device, (B, 7))
        status, pst = self._out_i32(device, (B,))
        iters, pit = self._out_i32(device, (B,))
        cost, pc = self._out(device, (B,))
        ni, pni = self._out_i32(device, (B,))
        rit, prit = self._out_i32(device, (B,))
        idx, pidx = self._out_i32(device, (max(total, 1),)) if opts.use_ransac else (None, None)
        stage = np.zeros(3, np.float32) if stage_timing else None
        o = _FrameOut(pp, pes, pst, pit, pc, pni, pidx, prit,
                      ctypes.c_void_p(stage.ctypes.data) if stage_timing else None)
        rc = self._lib.pnec_frame_solve_batch(self._h, ctypes.byref(b), ctypes.byref(opts), ctypes.byref(o),
                                              self._stream(device))
        self._check(rc, "pnec_frame_solve_batch")
        res = FrameResult(poses, es, status, iters, cost)
        res.num_inliers, res.ransac_iterations = ni, rit
        res.inlier_index = None if idx is None else idx[:total]
        res.stage_ms = stage
        return res


    # ------------------------------------------------------------ from keypoints
    def _keypoint_batch(self, host_points, target_points, target_covs2, init_poses, K_inv, host_covs2, host_index,
                        target_index, packed, offsets, n_per_problem, keep):
        device = _is_torch(host_points)
        kb = _KeypointBatch()
        cd = 3 if packed else 4
        if device:
            import torch

            def chk(t, dtype):
                if t is None:
                    return None
                if not (t.is_cuda and t.dtype == dtype and t.is_contiguous() and t.device.index == self.device):
                    raise PnecError("device keypoint batches need contiguous CUDA tensors on the handle's device "
                                    "(float64 tables, int32 indices)")
                return t

            hp, tp, hc, tc, poses = (chk(x, torch.float64) for x in (host_points, target_points, host_covs2,
                                                                      target_covs2, init_poses))
            hi, ti = chk(host_index, torch.int32), chk(target_index, torch.int32)
            size = lambda t: t.numel()
            ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
            kb.memspace = MEM_DEVICE
        else:
            f64 = lambda a, tail: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape((-1,) + tail)
            i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
            hp, tp = f64(host_points, (2,)), f64(target_points, (2,))
            hc, tc = f64(host_covs2, (cd,)), f64(target_covs2, (cd,))
            poses = f64(init_poses, (7,))
            hi, ti = i32(host_index), i32(target_index)
            size = lambda a: a.size
            ptr = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)
            kb.memspace = MEM_HOST
        kinv = np.ascontiguousarray(K_inv, dtype=np.float64).reshape(9)  # column-major, as keypoints_unproject
        keep.extend([hp, tp, hc, tc, poses, hi, ti, kinv])
        kh, kt = size(hp) // 2, size(tp) // 2
        B = size(poses) // 7
        if hc is not None and size(hc) != kh * cd:
            raise PnecError("host_covs2 does not match host_points")
        if tc is not None and size(tc) != kt * cd:
            raise PnecError("target_covs2 does not match target_points")
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            if offsets.shape != (B + 1,):
                raise PnecError("offsets must have num_problems + 1 entries")
            keep.append(offsets)
            kb.offsets = ctypes.c_void_p(offsets.ctypes.data)
            kb.n_per_problem = 0
            total = int(offsets[-1])
        else:
            if n_per_problem is None:
                ref = size(hi) if hi is not None else kh
                n_per_problem = ref // B if B else 0
            kb.offsets = None
            kb.n_per_problem = int(n_per_problem)
            total = B * int(n_per_problem)
        for name, idx in (("host_index", hi), ("target_index", ti)):
            if idx is not None and size(idx) != total:
                raise PnecError(f"{name} must have one entry per correspondence")
        kb.num_problems = B
        kb.packed_covs = 1 if packed else 0
        kb.num_host_keypoints, kb.num_target_keypoints = kh, kt
        kb.host_points, kb.target_points = ptr(hp), ptr(tp)
        kb.host_covs2, kb.target_covs2 = ptr(hc), ptr(tc)
        kb.host_index, kb.target_index = ptr(hi), ptr(ti)
        kb.K_inv = ctypes.c_void_p(kinv.ctypes.data)
        kb.poses = ptr(poses)
        return kb, device, B, total

    def solve_from_keypoints(self, host_points, target_points, target_covs2, init_poses, K_inv,
                             opts: Optional[SolverOpts] = None, *, host_covs2=None, host_index=None,
                             target_index=None, packed=False, offsets=None, n_per_problem=None,
                             out: Optional[SolveResult] = None) -> SolveResult:
        """pnec_solve_batch fed from keypoints (pixel + 2x2 image covariance per keypoint, the fields
        of the reference's KeyPoint; Frame2Frame::GetFeatures on the device)."""
        opts = opts or default_opts(TARGET)
        keep = []
        kb, device, B, _ = self._keypoint_batch(host_points, target_points, target_covs2, init_poses, K_inv,
                                                host_covs2, host_index, target_index, packed, offsets,
                                                n_per_problem, keep)
        if out is None:
            poses, _ = self._out(device, (B, 7))
            status, _ = self._out_i32(device, (B,))
            iters, _ = self._out_i32(device, (B,))
            cost, _ = self._out(device, (B,))
            init_cost, _ = self._out(device, (B,))
            out = SolveResult(poses, status, iters, cost, init_cost)
        p = (lambda t: ctypes.c_void_p(t.data_ptr())) if device else (lambda a: ctypes.c_void_p(a.ctypes.data))
        o = _SolveOut(p(out.poses), p(out.status), p(out.iterations), p(out.cost), p(out.initial_cost))
        rc = self._lib.pnec_solve_from_keypoints_batch(self._h, ctypes.byref(kb), ctypes.byref(opts), ctypes.byref(o),
                                                       self._stream(device))
        self._check(rc, "pnec_solve_from_keypoints_batch")
        return out

    def frame_solve_from_keypoints(self, host_points, target_points, target_covs2, init_poses, K_inv,
                                   opts: Optional[FrameOpts] = None, *, host_index=None, target_index=None,
                                   packed=False, offsets=None, n_per_problem=None) -> FrameResult:
        """pnec_frame_solve_batch (PNEC::Solve) fed from keypoints."""
        opts = opts or default_frame_opts()
        keep = []
        kb, device, B, total = self._keypoint_batch(host_points, target_points, target_covs2, init_poses, K_inv, None,
                                                    host_index, target_index, packed, offsets, n_per_problem, keep)
        poses, pp = self._out(device, (B, 7))
        es, pes = self._out(device, (B, 7))
        status, pst = self._out_i32(device, (B,))
        iters, pit = self._out_i32(device, (B,))
        cost, pc = self._out(device, (B,))
        ni, pni = self._out_i32(device, (B,))
        rit, prit = self._out_i32(device, (B,))
        idx, pidx = self._out_i32(device, (max(total, 1),)) if opts.use_ransac else (None, None)
        o = _FrameOut(pp, pes, pst, pit, pc, pni, pidx, prit, None)
        rc = self._lib.pnec_frame_solve_from_keypoints_batch(self._h, ctypes.byref(kb), ctypes.byref(opts),
                                                             ctypes.byref(o), self._stream(device))
        self._check(rc, "pnec_frame_solve_from_keypoints_batch")
        res = FrameResult(poses, es, status, iters, cost)
        res.num_inliers, res.ransac_iterations = ni, rit
        res.inlier_index = None if idx is None else idx[:total]
        return res


_default_handles = {}


def default_handle(device: int = 0) -> Handle:
    h = _default_handles.get(device)
    if h is None:
        h = _default_handles[device] = Handle(device)
    return h
