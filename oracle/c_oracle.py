"""TEST INFRASTRUCTURE -- ctypes wrapper of the C oracle (oracle/drm_oracle.c -> oracle/_ref/libdrm_oracle.so).

Same role and restrictions as oracle/drm_oracle.py: a CPU checker / CPU baseline, never a product path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libdrm_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(_SO)
    return _lib


class CRobot:
    """Flat arrays of a drm_oracle.Robot in the layout drm_oracle.c expects."""

    def __init__(self, robot, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.suffix = "f32" if self.dtype == np.float32 else "f64"
        self.n_links, self.n_dofs = len(robot.names), robot.n_dofs
        f = lambda t: np.ascontiguousarray(t.detach().numpy().astype(self.dtype))  # noqa: E731
        self.arrays = dict(
            parent=np.ascontiguousarray(np.array(robot.parent, dtype=np.int32)),
            dof=np.ascontiguousarray(np.array(robot.dof, dtype=np.int32)),
            axis=f(robot.axis), trans=f(robot.trans), rpy=f(robot.rpy), mass=f(robot.mass), com=f(robot.com),
            inertia=f(robot.inertia.reshape(-1, 9)), damping=f(robot.damping))
        real_p = ctypes.POINTER(ctypes.c_float if self.dtype == np.float32 else ctypes.c_double)

        class Struct(ctypes.Structure):
            _fields_ = [("n_links", ctypes.c_int), ("n_dofs", ctypes.c_int),
                        ("parent", ctypes.POINTER(ctypes.c_int)), ("dof", ctypes.POINTER(ctypes.c_int)),
                        ("axis", real_p), ("trans", real_p), ("rpy", real_p), ("mass", real_p), ("com", real_p),
                        ("inertia", real_p), ("damping", real_p)]

        a = self.arrays
        self.struct = Struct(self.n_links, self.n_dofs, a["parent"].ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                             a["dof"].ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                             *[a[k].ctypes.data_as(real_p) for k in ("axis", "trans", "rpy", "mass", "com", "inertia", "damping")])
        self.real_p = real_p

    def _p(self, arr):
        return None if arr is None else arr.ctypes.data_as(self.real_p)

    def fk_jacobian(self, ee, q, n_threads=0):
        q = np.ascontiguousarray(q, dtype=self.dtype)
        B, n = q.shape
        pos, quat = np.empty((B, 3), self.dtype), np.empty((B, 4), self.dtype)
        jl, ja = np.empty((B, 3, n), self.dtype), np.empty((B, 3, n), self.dtype)
        fn = getattr(_load(), f"drm_oracle_fk_jacobian_{self.suffix}")
        fn.restype = None
        fn(ctypes.byref(self.struct), ctypes.c_int(ee), self._p(q), ctypes.c_long(B), self._p(pos), self._p(quat),
           self._p(jl), self._p(ja), ctypes.c_int(n_threads))
        return pos, quat, jl, ja

    def inverse_dynamics(self, q, qd, qdd, gravity=True, damping=True, n_threads=0):
        q, qd, qdd = (np.ascontiguousarray(t, dtype=self.dtype) for t in (q, qd, qdd))
        B, n = q.shape
        tau = np.zeros((B, n), self.dtype)
        fn = getattr(_load(), f"drm_oracle_inverse_dynamics_{self.suffix}")
        fn.restype = None
        fn(ctypes.byref(self.struct), self._p(q), self._p(qd), self._p(qdd), ctypes.c_long(B), ctypes.c_int(gravity),
           ctypes.c_int(damping), self._p(tau), ctypes.c_int(n_threads))
        return tau

    def forward_dynamics(self, q, qd, f, gravity=True, damping=False, n_threads=0):
        q, qd, f = (np.ascontiguousarray(t, dtype=self.dtype) for t in (q, qd, f))
        B, n = q.shape
        qdd = np.zeros((B, n), self.dtype)
        fn = getattr(_load(), f"drm_oracle_forward_dynamics_{self.suffix}")
        fn.restype = None
        fn(ctypes.byref(self.struct), self._p(q), self._p(qd), self._p(f), ctypes.c_long(B), ctypes.c_int(gravity),
           ctypes.c_int(damping), self._p(qdd), ctypes.c_int(n_threads))
        return qdd
