"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_build/libfjsp_oracle.so.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never from the product package."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libfjsp_oracle.so")
VARIANTS = {"SO_DFJSP": 0, "MO_DFJSP": 1, "MO_DFJSP_breakdown": 2, "SO_FJSSP": 3}
NSTATE = {0: 20, 1: 30, 2: 30, 3: 20}
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("fjsp_lp.c", "fjsp_oracle.c", "fjsp_oracle_batch.c", "pyemu.h", "Makefile")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.fjsp_oracle_create.restype = ctypes.c_void_p
        L.fjsp_oracle_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.fjsp_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.fjsp_oracle_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.fjsp_oracle_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.fjsp_oracle_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.fjsp_oracle_machine_end.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.fjsp_oracle_batch_rollout.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.fjsp_oracle_max_threads.restype = ctypes.c_int
        L.fjsp_lp_solve_sparse.restype = ctypes.c_int
        L.fjsp_lp_solve_sparse.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p]
        L.fjsp_pysum.restype = ctypes.c_double
        L.fjsp_pysum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.fjsp_pyset_order.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.fjsp_pyset_intersection_list.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.c_void_p]
        _lib = L
    return _lib


class OracleEnv:
    """One environment with the reference's reset()/step(action) contract."""

    def __init__(self, blob, variant, sum_mode=1):
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.blob = np.ascontiguousarray(blob, dtype=np.int32)
        self.h = lib().fjsp_oracle_create(self.blob.ctypes.data, self.variant, sum_mode)
        if not self.h:
            raise ValueError("bad instance blob")
        self.nstate = NSTATE[self.variant]
        self.M = int(self.blob[2])

    def __del__(self):
        if getattr(self, "h", None):
            lib().fjsp_oracle_destroy(self.h)
            self.h = None

    def reset(self):
        st = np.zeros(self.nstate)
        rc = lib().fjsp_oracle_reset(self.h, st.ctypes.data)
        assert rc == 0, rc
        return st

    def step(self, action, rnd=(0, 0), reward_policy=1, completion=1.0, tardiness=1.0, energy=1.0):
        st = np.zeros(self.nstate)
        rw = ctypes.c_double(0)
        dn = ctypes.c_int(0)
        rec = np.zeros(8, np.int32)
        rc = lib().fjsp_oracle_step(self.h, int(action[0]), int(action[1]), int(rnd[0]), int(rnd[1]),
                                    reward_policy, completion, tardiness, energy,
                                    st.ctypes.data, ctypes.byref(rw), ctypes.byref(dn), rec.ctypes.data)
        assert rc == 0, "oracle error flags %d" % rc
        return st, rw.value, bool(dn.value), rec

    def info(self):
        a = np.zeros(10, np.int64)
        lib().fjsp_oracle_info(self.h, a.ctypes.data)
        keys = ["step_time", "step_count", "completion", "delay_sum", "energy", "lp_solves", "lp_iters",
                "error", "done", "next_order"]
        return dict(zip(keys, (int(x) for x in a)))

    def machine_end(self):
        a = np.zeros(self.M, np.int32)
        lib().fjsp_oracle_machine_end(self.h, a.ctypes.data)
        return a


def batch_rollout(envs, actions, rnd, reward_policy=1, want_state=True, want_rec=True, threads=0, allow_errors=False):
    """actions, rnd: [T, B, 2].  Returns dict of [T, B, ...] arrays.  Auto-resets finished envs.
    allow_errors: do not raise when an environment raised its error flag (the reference would raise an
    exception there); the caller reads the flags with info() -- that environment's outputs are undefined from
    the failing step on."""
    T, B = actions.shape[:2]
    nstate = envs[0].nstate
    actions = np.ascontiguousarray(actions, np.int32)
    rnd = np.ascontiguousarray(rnd, np.uint32)
    handles = (ctypes.c_void_p * B)(*[e.h for e in envs])
    state = np.zeros((T, B, nstate)) if want_state else None
    reward = np.zeros((T, B))
    done = np.zeros((T, B), np.int32)
    rec = np.zeros((T, B, 8), np.int32) if want_rec else None
    rc = lib().fjsp_oracle_batch_rollout(handles, B, T, actions.ctypes.data, rnd.ctypes.data, reward_policy, nstate,
                                         state.ctypes.data if want_state else None, reward.ctypes.data,
                                         done.ctypes.data, rec.ctypes.data if want_rec else None, threads)
    assert rc == 0 or allow_errors, "oracle error flags %d" % rc
    return dict(state=state, reward=reward, done=done, rec=rec, error_flags=rc)
