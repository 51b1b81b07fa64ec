"""ctypes binding of oracle/cpu_port.cpp (ORACLE side: tests / smoke / bench cpu_baseline only)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libhsr_cpu_port.so"
SRC = [HERE / "cpu_port.cpp", HERE.parent / "hsr_env_b200" / "csrc" / "hsr_core.h",
       HERE.parent / "hsr_env_b200" / "csrc" / "hsr_model.h"]


def build(force=False):
    LIB.parent.mkdir(exist_ok=True)
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= s.stat().st_mtime for s in SRC):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(LIB), str(SRC[0]), "-lpthread"]
    subprocess.check_call(cmd)
    return LIB


_dp = ctypes.POINTER(ctypes.c_double)


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


class CpuPort:
    def __init__(self, model):
        lib = ctypes.CDLL(str(build()))
        lib.hsrp_create.restype = ctypes.c_void_p
        lib.hsrp_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        lib.hsrp_destroy.argtypes = [ctypes.c_void_p]
        lib.hsrp_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        lib.hsrp_set_caps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.hsrp_set_goals.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int]
        lib.hsrp_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [_dp] * 8 + [
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(ctypes.c_int), _dp,
            ctypes.POINTER(ctypes.c_longlong)]
        _ip = ctypes.POINTER(ctypes.c_int)
        lib.hsrp_set_goal_list.argtypes = [ctypes.c_void_p, ctypes.c_int, _ip, _ip, _dp, _dp, _dp, ctypes.c_int]
        lib.hsrp_set_starts.argtypes = [ctypes.c_void_p, ctypes.c_int, _ip, _ip, _dp, _dp]
        lib.hsrp_set_scan_noise.argtypes = [ctypes.c_ulonglong]
        lib.hsrp_reset.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, _dp, _dp]
        self.lib = lib
        blob = model.to_blob()
        self.h = lib.hsrp_create(blob, len(blob))
        if not self.h:
            raise RuntimeError("hsrp_create failed")
        self.model = model
        self._dims()

    def _dims(self):
        d = (ctypes.c_int * 8)()
        self.lib.hsrp_dims(self.h, d)
        self.nq, self.nv, self.nu, self.nbody, self.ncon_max, self.nefc_max, self.debug_size, self.ws_bytes_f32 = list(d)

    def set_caps(self, ncon_max, nefc_max):
        self.lib.hsrp_set_caps(self.h, ncon_max, nefc_max)
        self._dims()

    def set_goals(self, goal_lohi=None, block_lohi=None, geofence=0.0, min_sep=0.0, qidx=(0, 2)):
        has = goal_lohi is not None
        g = np.ascontiguousarray(np.zeros(6) if goal_lohi is None else np.asarray(goal_lohi, float).reshape(6))
        b = np.ascontiguousarray(np.zeros(8) if block_lohi is None else np.asarray(block_lohi, float).reshape(8))
        self.lib.hsrp_set_goals(self.h, int(has), int(block_lohi is not None), _p(g), _p(b), float(geofence), float(min_sep), qidx[0], qidx[1])

    def set_goal_list(self, a, b, dist, point_lohi=None, fixed=None):
        """mirror of hsrb_set_goal_list: endpoint codes >= 0 body id, -1 the goal point, -2-k fixed point k"""
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32); dist = np.ascontiguousarray(dist, float)
        pl = None if point_lohi is None else np.ascontiguousarray(point_lohi, float).reshape(6)
        fx = np.zeros((0, 3)) if fixed is None else np.ascontiguousarray(fixed, float).reshape(-1, 3)
        self.lib.hsrp_set_goal_list(self.h, len(a), _p(a, ctypes.c_int), _p(b, ctypes.c_int), _p(dist), _p(pl) if pl is not None else None,
                                    _p(fx) if len(fx) else None, len(fx))

    def set_starts(self, adr, width, lo, hi):
        adr = np.ascontiguousarray(adr, np.int32); width = np.ascontiguousarray(width, np.int32)
        lo = np.ascontiguousarray(lo, float).reshape(-1, 7); hi = np.ascontiguousarray(hi, float).reshape(-1, 7)
        self.lib.hsrp_set_starts(self.h, len(adr), _p(adr, ctypes.c_int), _p(width, ctypes.c_int), _p(lo), _p(hi))

    def step(self, qpos, qvel, warm, ctrl, mocap=None, nsub=1, use_float=False, nthreads=1, debug=False):
        qpos = np.ascontiguousarray(np.atleast_2d(qpos), float)
        n = qpos.shape[0]
        qvel = np.ascontiguousarray(np.atleast_2d(qvel), float)
        warm = np.ascontiguousarray(np.atleast_2d(warm), float)
        ctrl = np.ascontiguousarray(np.atleast_2d(ctrl), float).reshape(n, max(self.nu, 0))
        mocap = np.zeros((n, 3)) if mocap is None else np.ascontiguousarray(np.atleast_2d(mocap), float)
        qo, vo, wo = np.zeros_like(qpos), np.zeros_like(qvel), np.zeros_like(warm)
        taken = np.zeros(n, np.int32); succ = np.zeros(n, np.uint8); flags = np.zeros(n, np.int32)
        dbg = np.zeros((n, self.debug_size)) if debug else None
        cnt = np.zeros((n, 4), np.int64)
        self.lib.hsrp_step(self.h, int(use_float), n, nsub, nthreads, _p(qpos), _p(qvel), _p(warm), _p(ctrl), _p(mocap),
                           _p(qo), _p(vo), _p(wo), _p(taken, ctypes.c_int), _p(succ, ctypes.c_ubyte),
                           _p(flags, ctypes.c_int), _p(dbg), _p(cnt, ctypes.c_longlong))
        out = dict(qpos=qo, qvel=vo, warm=wo, taken=taken, success=succ, flags=flags, counters=cnt)
        if debug:
            out["debug"] = [unpack_debug(self, dbg[i]) for i in range(n)]
        return out

    def set_scan_noise(self, seed):
        """fp32-rounding-sized noise on the hull-vertex scans (single-threaded calls of this thread); 0 switches it off"""
        self.lib.hsrp_set_scan_noise(int(seed))

    def reset(self, seed, env_id, episode):
        q = np.zeros(self.nq); mo = np.zeros(3)
        self.lib.hsrp_reset(self.h, seed, env_id, episode, _p(q), _p(mo))
        return q, mo

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.hsrp_destroy(self.h)
            self.h = None


def unpack_debug(dims, v):
    """Inverse of hsr::debug_dump (hsr_env_b200/csrc/hsr_core.h)."""
    nv, nb, nc, ne, nq = dims.nv, dims.nbody, dims.ncon_max, dims.nefc_max, dims.nq
    k = 0

    def take(n, shape=None):
        nonlocal k
        a = v[k:k + n]
        k += n
        return a.reshape(shape) if shape else a

    ncon, nefc, nlimit, iters = [int(x) for x in take(4)]
    out = dict(ncon=ncon, nefc=nefc, nlimit=nlimit, iters=iters)
    out["xpos"] = take(nb * 3, (nb, 3))
    out["M"] = take(nv * nv, (nv, nv))
    out["qfrc_smooth"] = take(nv); out["qacc_smooth"] = take(nv); out["qacc"] = take(nv)
    con = take(nc * 14, (nc, 14))[:ncon]
    out["con_pair"] = con[:, 0].astype(int); out["con_dist"] = con[:, 1]; out["con_pos"] = con[:, 2:5]
    out["con_frame"] = con[:, 5:14].reshape(-1, 3, 3)
    efc = take(ne * (nv + 3), (ne, nv + 3))[:nefc]
    out["efc_J"] = efc[:, :nv]; out["efc_D"] = efc[:, nv]; out["efc_aref"] = efc[:, nv + 1]; out["efc_force"] = efc[:, nv + 2]
    out["qpos"] = take(nq); out["qvel"] = take(nv)
    return out
