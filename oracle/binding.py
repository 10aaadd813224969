"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/liblh2oracle.so (built by oracle/Makefile)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblh2oracle.so")
_lib = None


class OrcMesh(ctypes.Structure):
    _fields_ = [("verts4", ctypes.c_void_p), ("triCount", ctypes.c_int)]


class OrcInstance(ctypes.Structure):
    _fields_ = [("mesh", ctypes.c_int), ("xform", ctypes.c_float * 12)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _scene(meshes, instances):
    """meshes: list of float32[3T,4]; instances: list of (meshIdx, 4x4 or None)."""
    keep = [np.ascontiguousarray(m, np.float32).reshape(-1, 4) for m in meshes]
    cm = (OrcMesh * len(keep))()
    for i, m in enumerate(keep):
        cm[i].verts4, cm[i].triCount = m.ctypes.data, m.shape[0] // 3
    ci = (OrcInstance * max(len(instances), 1))()
    for i, (mi, xf) in enumerate(instances):
        ci[i].mesh = mi
        x = np.eye(4, dtype=np.float32) if xf is None else np.asarray(xf, np.float32).reshape(4, 4)
        for k in range(12):
            ci[i].xform[k] = float(x.flat[k])
    return keep, cm, ci


def closest_hits(meshes, instances, O4, D4, threads=None):
    keep, cm, ci = _scene(meshes, instances)
    O4 = np.ascontiguousarray(O4, np.float32).reshape(-1, 4)
    D4 = np.ascontiguousarray(D4, np.float32).reshape(-1, 4)
    n = O4.shape[0]
    hits = np.empty((n, 4), np.uint32)
    lib().orc_closest_hits(cm, len(keep), ci, len(instances), ctypes.c_void_p(O4.ctypes.data), ctypes.c_void_p(D4.ctypes.data),
                           n, ctypes.c_void_p(hits.ctypes.data), threads or os.cpu_count())
    return hits


def occluded(meshes, instances, O4, D4, threads=None):
    keep, cm, ci = _scene(meshes, instances)
    O4 = np.ascontiguousarray(O4, np.float32).reshape(-1, 4)
    D4 = np.ascontiguousarray(D4, np.float32).reshape(-1, 4)
    n = O4.shape[0]
    occ = np.empty(n, np.uint8)
    lib().orc_occluded(cm, len(keep), ci, len(instances), ctypes.c_void_p(O4.ctypes.data), ctypes.c_void_p(D4.ctypes.data),
                       n, ctypes.c_void_p(occ.ctypes.data), threads or os.cpu_count())
    return occ
