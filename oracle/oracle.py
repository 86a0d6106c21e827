"""ctypes binding of the C oracle (oracle/clc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (camlasercalibratool_b200/) never does.  PARITY UNPINNED -- see clc_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libclc_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile next to this file (gcc only, no GPU needed)."""
    src = [os.path.join(_HERE, n) for n in ("clc_oracle.c", "clc_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libclc_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Problem(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int64),
        ("frame_pose", C.POINTER(C.c_double)),
        ("offsets", C.POINTER(C.c_int64)),
        ("points", C.POINTER(C.c_double)),
        ("edge_points", C.POINTER(C.c_double)),
        ("use_loss", C.c_int),
        ("cauchy_a", C.c_double),
    ]


class Options(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int),
        ("jacobi_scaling", C.c_int),
        ("linear_solver", C.c_int),
        ("num_threads", C.c_int),
    ]


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int),
        ("step_is_valid", C.c_int),
        ("step_is_successful", C.c_int),
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double),
        ("step_norm", C.c_double),
        ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("termination", C.c_int),
        ("num_iterations", C.c_int),
        ("num_successful_steps", C.c_int),
        ("num_unsuccessful_steps", C.c_int),
        ("num_residual_evaluations", C.c_int),
        ("num_jacobian_evaluations", C.c_int),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
    ]


class _GenDesc(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int64),
        ("beams", C.c_int64),
        ("seed", C.c_uint64),
        ("sigma", C.c_double),
        ("exact_m", C.c_int),
        ("with_edges", C.c_int),
    ]


TERMINATION = {
    1: "CONVERGENCE_FUNCTION",
    2: "CONVERGENCE_PARAMETER",
    3: "CONVERGENCE_GRADIENT",
    4: "CONVERGENCE_MIN_RADIUS",
    5: "NO_CONVERGENCE",
    6: "FAILURE",
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.oracle_default_options.argtypes = [C.POINTER(Options)]
        L.oracle_quat_to_rot.argtypes = [dp, dp]
        L.oracle_rot_to_quat.argtypes = [dp, dp]
        L.oracle_T_to_pose7.argtypes = [dp, dp]
        L.oracle_pose7_to_T.argtypes = [dp, dp]
        L.oracle_pose_plus.argtypes = [dp, dp, dp]
        L.oracle_frame_plane.argtypes = [dp, dp]
        L.oracle_edge_planes.argtypes = [dp, dp, dp]
        L.oracle_factor_evaluate.argtypes = [dp, dp, C.c_double, dp, dp, dp]
        L.oracle_num_residuals.argtypes = [C.POINTER(_Problem)]
        L.oracle_num_residuals.restype = C.c_int64
        L.oracle_evaluate.argtypes = [C.POINTER(_Problem), dp, dp, dp, dp, dp, C.c_int]
        L.oracle_evaluate_normal.argtypes = [C.POINTER(_Problem), dp, dp, dp, dp, C.c_int]
        L.oracle_solve.argtypes = [C.POINTER(_Problem), dp, C.POINTER(Options), C.POINTER(Summary),
                                   C.POINTER(Iteration), C.c_int]
        L.oracle_information.argtypes = [C.POINTER(_Problem), dp, dp, dp, dp, dp]
        L.oracle_closed_form.argtypes = [C.POINTER(_Problem), dp, C.POINTER(C.c_int), dp, dp]
        L.oracle_sym_singular_values.argtypes = [dp, C.c_int, dp]
        L.oracle_scan_to_points.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_double, C.c_double, C.c_double, dp]
        L.oracle_auto_get_line_pts.argtypes = [dp, C.c_int64, ip, ip]
        L.oracle_line_fit.argtypes = [dp, C.c_int64, dp, C.c_int, C.POINTER(Summary), C.POINTER(Iteration), C.c_int]
        L.oracle_gen_ground_truth.argtypes = [dp, dp]
        L.oracle_gen_frames.argtypes = [C.POINTER(_GenDesc), dp, ip]
        L.oracle_gen_frames.restype = C.c_int64
        L.oracle_gen_points.argtypes = [C.POINTER(_GenDesc), dp, ip, dp, dp]
        L.oracle_philox4x32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class Problem:
    """Flat (marshalled) view of a std::vector<Oberserve>: what the C ABI of the product also takes."""

    frame_pose: np.ndarray            # [N,7]  qx qy qz qw tx ty tz
    offsets: np.ndarray               # [N+1]  int64 CSR
    points: np.ndarray                # [P,3]
    edge_points: np.ndarray | None = None  # [N,6] points.front(), points.back()
    use_loss: bool = True
    cauchy_a: float = 0.05
    _c: _Problem = field(default=None, repr=False)

    def __post_init__(self):
        self.frame_pose = _f64(self.frame_pose, (-1, 7))
        self.offsets = np.ascontiguousarray(self.offsets, dtype=np.int64)
        self.points = _f64(self.points, (-1, 3))
        if self.edge_points is not None:
            self.edge_points = _f64(self.edge_points, (-1, 6))
        assert self.offsets.shape[0] == self.frame_pose.shape[0] + 1
        assert self.offsets[-1] == self.points.shape[0]
        c = _Problem()
        c.n_frames = self.frame_pose.shape[0]
        c.frame_pose = _dp(self.frame_pose)
        c.offsets = self.offsets.ctypes.data_as(C.POINTER(C.c_int64))
        c.points = _dp(self.points)
        c.edge_points = _dp(self.edge_points)
        c.use_loss = 1 if self.use_loss else 0
        c.cauchy_a = float(self.cauchy_a)
        self._c = c

    @property
    def n_frames(self):
        return self.frame_pose.shape[0]

    @property
    def n_points(self):
        return self.points.shape[0]

    def num_residuals(self):
        return int(lib().oracle_num_residuals(C.byref(self._c)))


def default_options(**kw) -> Options:
    o = Options()
    lib().oracle_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def quat_to_rot(q):
    q = _f64(q)
    R = np.empty(9)
    lib().oracle_quat_to_rot(_dp(q), _dp(R))
    return R.reshape(3, 3)


def rot_to_quat(R):
    R = _f64(R).reshape(9)
    q = np.empty(4)
    lib().oracle_rot_to_quat(_dp(R), _dp(q))
    return q


def T_to_pose7(T):
    T = _f64(T).reshape(16)
    p = np.empty(7)
    lib().oracle_T_to_pose7(_dp(T), _dp(p))
    return p


def pose7_to_T(p):
    p = _f64(p)
    T = np.empty(16)
    lib().oracle_pose7_to_T(_dp(p), _dp(T))
    return T.reshape(4, 4)


def pose_plus(x, d):
    x, d = _f64(x), _f64(d)
    xp = np.empty(7)
    lib().oracle_pose_plus(_dp(x), _dp(d), _dp(xp))
    return xp


def frame_plane(fp):
    fp = _f64(fp)
    pl = np.empty(4)
    lib().oracle_frame_plane(_dp(fp), _dp(pl))
    return pl


def edge_planes(fp):
    fp = _f64(fp)
    a, b = np.empty(4), np.empty(4)
    lib().oracle_edge_planes(_dp(fp), _dp(a), _dp(b))
    return a, b


def factor_evaluate(plane, pt, scale, pose7):
    plane, pt, pose7 = _f64(plane), _f64(pt), _f64(pose7)
    r = C.c_double()
    j = np.empty(7)
    lib().oracle_factor_evaluate(_dp(plane), _dp(pt), float(scale), _dp(pose7), C.byref(r), _dp(j))
    return r.value, j


def evaluate(p: Problem, pose7, jac=True, num_threads=1):
    """Ceres-shaped evaluation: cost, residuals[R], jacobian[R,6], gradient[6]."""
    pose7 = _f64(pose7)
    R = p.num_residuals()
    cost = C.c_double()
    res = np.empty(R)
    J = np.empty((R, 6)) if jac else None
    g = np.empty(6) if jac else None
    lib().oracle_evaluate(C.byref(p._c), _dp(pose7), C.byref(cost), _dp(res), _dp(J), _dp(g), num_threads)
    return cost.value, res, J, g


def evaluate_normal(p: Problem, pose7, jac=True, num_threads=1):
    """Streaming evaluation: cost, H[6,6], g[6]."""
    pose7 = _f64(pose7)
    cost = C.c_double()
    H = np.empty((6, 6)) if jac else None
    g = np.empty(6) if jac else None
    lib().oracle_evaluate_normal(C.byref(p._c), _dp(pose7), C.byref(cost), _dp(H), _dp(g), num_threads)
    return cost.value, H, g


def solve(p: Problem, pose7, options: Options | None = None, trace_cap=256):
    """The Ceres trust-region LM solve.  Returns (pose7_out, Summary, [Iteration...])."""
    x = _f64(pose7).copy()
    o = options if options is not None else default_options()
    s = Summary()
    tr = (Iteration * trace_cap)()
    lib().oracle_solve(C.byref(p._c), _dp(x), C.byref(o), C.byref(s), tr, trace_cap)
    return x, s, [tr[i] for i in range(min(s.num_iterations, trace_cap))]


def information(p: Problem, pose7):
    pose7 = _f64(pose7)
    H, b, sv = np.empty((6, 6)), np.empty(6), np.empty(6)
    chi = C.c_double()
    lib().oracle_information(C.byref(p._c), _dp(pose7), _dp(H), _dp(b), C.byref(chi), _dp(sv))
    return H, b, chi.value, sv


def closed_form(p: Problem):
    T, AtA, Atb = np.empty(16), np.empty((9, 9)), np.empty(9)
    un = C.c_int()
    lib().oracle_closed_form(C.byref(p._c), _dp(T), C.byref(un), _dp(AtA), _dp(Atb))
    return T.reshape(4, 4), bool(un.value), AtA, Atb


def scan_to_points(ranges, angle_min, angle_increment, range_min):
    r = np.ascontiguousarray(ranges, dtype=np.float32)
    pts = np.empty((r.shape[0], 3))
    lib().oracle_scan_to_points(r.ctypes.data_as(C.POINTER(C.c_float)), r.shape[0], angle_min, angle_increment, range_min, _dp(pts))
    return pts


def auto_get_line_pts(points):
    """(start, end) inclusive indices of the chosen segment, or None."""
    pts = _f64(points, (-1, 3))
    s, e = C.c_int64(), C.c_int64()
    if not lib().oracle_auto_get_line_pts(_dp(pts), pts.shape[0], C.byref(s), C.byref(e)):
        return None
    return s.value, e.value


def line_fit(points, line0=(0.0, 0.0), max_num_iterations=10, trace_cap=64):
    """LineFittingCeres restatement: returns (line[2], Summary, [Iteration...])."""
    pts = _f64(points, (-1, 3))
    line = _f64(line0).copy()
    s = Summary()
    tr = (Iteration * trace_cap)()
    lib().oracle_line_fit(_dp(pts), pts.shape[0], _dp(line), int(max_num_iterations), C.byref(s), tr, trace_cap)
    return line, s, [tr[i] for i in range(min(s.num_iterations, trace_cap))]


def sym_singular_values(A):
    A = _f64(A)
    n = A.shape[0]
    sv = np.empty(n)
    lib().oracle_sym_singular_values(_dp(A), n, _dp(sv))
    return sv


def ground_truth():
    """(T_lc 4x4, T_cl as pose7) of the synthetic generator (calibr_simulation.cpp:15-20)."""
    T, p = np.empty(16), np.empty(7)
    lib().oracle_gen_ground_truth(_dp(T), _dp(p))
    return T.reshape(4, 4), p


def generate(n_frames, beams, seed=1, sigma=0.0, exact_m=False, with_edges=False, use_loss=True) -> Problem:
    g = _GenDesc(int(n_frames), int(beams), int(seed), float(sigma), int(bool(exact_m)), int(bool(with_edges)))
    fp = np.empty((n_frames, 7))
    off = np.empty(n_frames + 1, dtype=np.int64)
    total = lib().oracle_gen_frames(C.byref(g), _dp(fp), off.ctypes.data_as(C.POINTER(C.c_int64)))
    pts = np.empty((total, 3))
    ep = np.empty((n_frames, 6)) if with_edges else None
    lib().oracle_gen_points(C.byref(g), _dp(fp), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(pts), _dp(ep))
    return Problem(fp, off, pts, ep, use_loss=use_loss)


def philox(seed, lo, hi):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox4x32(seed, lo, hi, out)
    return [int(v) for v in out]


def pose_error(pose_a, pose_b):
    """(rotation angle in rad, translation distance in m) between two pose7 vectors."""
    Ra, Rb = quat_to_rot(_f64(pose_a)[3:]), quat_to_rot(_f64(pose_b)[3:])
    # normalise in case a quaternion is not exactly unit
    Ra = Ra / np.cbrt(np.linalg.det(Ra))
    Rb = Rb / np.cbrt(np.linalg.det(Rb))
    dR = Ra.T @ Rb
    # robust small-angle formula: |log(dR)| from the skew part and the trace
    s = 0.5 * np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    c = 0.5 * (np.trace(dR) - 1.0)
    ang = float(np.arctan2(s, c))
    dt = float(np.linalg.norm(_f64(pose_a)[:3] - _f64(pose_b)[:3]))
    return ang, dt
