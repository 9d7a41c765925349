"""TEST INFRASTRUCTURE ONLY: builds tests/hostsim/libfjsp_hostsim.so (the kernel source
compiled by g++ as a one-lane program) and wraps it with the same surface as the CUDA
vector environment, for the CPU parity tests."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "deep_reinforcement_learning_for_fjsp_b200", "csrc")
LIB = os.path.join(HERE, "libfjsp_hostsim.so")
_lib = None


def build():
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(CSRC, f) for f in
                                                  ("fjsp_core.cuh", "fjsp_host.h", "fjsp_layout.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-w",
                               "-o", LIB, os.path.join(HERE, "hostsim.cpp")])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.fjsp_hostsim_last_error.restype = ctypes.c_char_p
        L.fjsp_hostsim_create.argtypes = [vp, vp, i, vp, i, i, i, vp]
        L.fjsp_hostsim_destroy.argtypes = [vp]
        L.fjsp_hostsim_reset.argtypes = [vp, vp]
        L.fjsp_hostsim_step.argtypes = [vp, i, vp, vp, i, d, d, d, i, vp, vp, vp, vp, vp]
        L.fjsp_hostsim_info.argtypes = [vp, vp]
        L.fjsp_hostsim_query.argtypes = [vp, vp]
        L.fjsp_hostsim_pyset_order.argtypes = [vp, i, vp]
        L.fjsp_hostsim_selectable.argtypes = [ctypes.c_uint, vp, i, vp]
        _lib = L
    return _lib


INFO_KEYS = ["step_time", "step_count", "completion", "delay_sum", "energy", "lp_solves", "lp_iters", "error",
             "done", "next_order", "episodes", "delay_unprocessed"]
VARIANTS = {"SO_DFJSP": 0, "MO_DFJSP": 1, "MO_DFJSP_breakdown": 2, "SO_FJSSP": 3}


class HostSimVec:
    def __init__(self, blobs, env_instance, variant, sum_mode=1):
        L = lib()
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        offs = np.cumsum([0] + [len(b) for b in blobs[:-1]]).astype(np.int64)
        flat = np.ascontiguousarray(np.concatenate(blobs), np.int32)
        ei = np.ascontiguousarray(env_instance, np.int32)
        self.B = len(ei)
        h = ctypes.c_void_p()
        rc = L.fjsp_hostsim_create(flat.ctypes.data, offs.ctypes.data, len(blobs), ei.ctypes.data, self.B,
                                   self.variant, sum_mode, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(L.fjsp_hostsim_last_error().decode())
        self.h = h
        self.nstate = 20 if self.variant in (0, 3) else 30

    def __del__(self):
        if getattr(self, "h", None):
            lib().fjsp_hostsim_destroy(self.h)
            self.h = None

    def reset(self):
        st = np.zeros((self.B, self.nstate))
        lib().fjsp_hostsim_reset(self.h, st.ctypes.data)
        return st

    def step(self, actions, rnd=None, reward_policy=1, completion=1.0, tardiness=1.0, energy=1.0, autoreset=True):
        actions = np.ascontiguousarray(actions, np.int32)
        if actions.ndim == 2:
            actions = actions[None]
        T = actions.shape[0]
        rnd = np.zeros((T, self.B, 2), np.uint32) if rnd is None else np.ascontiguousarray(rnd, np.uint32).reshape(T, self.B, 2)
        st = np.zeros((T, self.B, self.nstate))
        rw = np.zeros((T, self.B))
        dn = np.zeros((T, self.B), np.int32)
        rec = np.zeros((T, self.B, 8), np.int32)
        lib().fjsp_hostsim_step(self.h, T, actions.ctypes.data, rnd.ctypes.data, reward_policy, completion, tardiness,
                                energy, int(autoreset), st.ctypes.data, None, rw.ctypes.data, dn.ctypes.data,
                                rec.ctypes.data)
        return st, rw, dn, rec

    reset_host = reset

    def step_host(self, actions, rnd=None, reward_policy=1, completion=1.0, tardiness=1.0, energy=1.0,
                  autoreset=True, want_rec=True, state_dtype=None):
        return self.step(actions, rnd, reward_policy, completion, tardiness, energy, autoreset)

    def info(self):
        a = np.zeros((self.B, 12), np.int64)
        lib().fjsp_hostsim_info(self.h, a.ctypes.data)
        return {k: a[:, i] for i, k in enumerate(INFO_KEYS)}
