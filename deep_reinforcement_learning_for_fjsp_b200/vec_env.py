"""Host-side mirror of the reference environments on top of the C ABI.

FJSPVecEnv   : B environment copies stepped by one kernel launch; tensors stay on the GPU
               (torch is used for device memory and streams only).
FJSPEnv      : ONE environment with the reference's exact surface -- reset() -> state,
               step(action) -> (state, reward, done), actions_size, state_size,
               action_tuple, done, ... -- so the agents in the reference's agents/ can
               consume it unchanged (environments/SO_DFJSP.py:13-52, MO_DFJSP.py:12-56).
rollout      : T steps per launch, the device-side replacement of
               utilities/Parallel_Experience_Generator.py:28-66.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .instance import FJSPInstance

VARIANTS = {"SO_DFJSP": 0, "MO_DFJSP": 1, "MO_DFJSP_breakdown": 2, "SO_FJSSP": 3}
ACTIONS_SIZE = {0: [6, 5], 1: [12, 10], 2: [12, 10], 3: [6, 5]}
INFO_KEYS = ["step_time", "step_count", "completion_time", "delay_time_sum", "energy_consumption", "lp_solves",
             "lp_iterations", "error", "done", "next_order", "episodes", "delay_time_sum_unprocessed"]


class MyError(Exception):
    """Same name as the reference's utilities.Utility_Class.MyError."""


class FJSPVecEnv:
    def __init__(self, instances: Sequence[FJSPInstance], env_instance: Optional[Sequence[int]] = None,
                 variant: str = "MO_DFJSP", device: int = 0, sum_mode: int = 1, blobs=None):
        import torch
        self.torch = torch
        L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("FJSPVecEnv needs a CUDA device; there is no CPU fallback")
        if variant not in VARIANTS:
            raise MyError("unknown environment variant %r (device path: %s)" % (variant, sorted(VARIANTS)))
        self.variant_name, self.variant = variant, VARIANTS[variant]
        blobs = [inst.to_blob() for inst in instances] if blobs is None else list(blobs)
        if env_instance is None:
            env_instance = np.arange(len(blobs))
        ei = np.ascontiguousarray(env_instance, np.int32)
        offs = np.cumsum([0] + [len(b) for b in blobs[:-1]]).astype(np.int64)
        flat = np.ascontiguousarray(np.concatenate(blobs), np.int32)
        self.n_envs, self.device = len(ei), device
        self.dev = torch.device("cuda", device)
        h = ctypes.c_void_p()
        _lib.check(L.fjsp_vec_create(flat.ctypes.data, offs.ctypes.data, len(blobs), ei.ctypes.data, self.n_envs,
                                     self.variant, sum_mode, device, ctypes.byref(h)))
        self._h, self._L = h, L
        self.actions_size = ACTIONS_SIZE[self.variant]
        self.state_size = 20 if self.variant in (0, 3) else 30
        self.observation_space = self.state_size // 2
        self.action_tuple = tuple((a1, a2) for a1 in range(self.actions_size[0]) for a2 in range(self.actions_size[1]))
        self.action_types = "DISCRETE"

    def close(self):
        if getattr(self, "_h", None):
            self._L.fjsp_vec_destroy(self._h)
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------ device API
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def reset(self, dtype=None):
        """reset() of every copy -> [n_envs, state_size] tensor on the GPU."""
        t = self.torch
        dtype = dtype or t.float64
        st = t.empty((self.n_envs, self.state_size), dtype=dtype, device=self.dev)
        p64 = st.data_ptr() if dtype == t.float64 else None
        p32 = st.data_ptr() if dtype == t.float32 else None
        _lib.check(self._L.fjsp_vec_reset(self._h, self._stream(), p64, p32))
        return st

    def rollout(self, actions, rnd=None, reward_policy=1, completion=1.0, tardiness=1.0, energy=1.0,
                autoreset=True, want_state=True, want_rec=False, state_dtype=None, out=None):
        """actions [T, n_envs, 2] int32 CUDA tensor (task rule, machine rule; 0-based as in the
        reference); rnd [T, n_envs, 2] draws for the random rules (int32/uint32 bit patterns).
        Returns dict(state [T,B,S], reward [T,B], done [T,B], rec [T,B,8])."""
        t = self.torch
        assert actions.is_cuda and actions.dtype == t.int32 and actions.is_contiguous()
        T = actions.shape[0]
        assert tuple(actions.shape) == (T, self.n_envs, 2)
        if rnd is not None:
            assert rnd.is_cuda and rnd.is_contiguous() and rnd.element_size() == 4 and rnd.numel() == actions.numel()
        sd = state_dtype or t.float64
        out = out or {}
        if want_state and "state" not in out:
            out["state"] = t.empty((T, self.n_envs, self.state_size), dtype=sd, device=self.dev)
        if "reward" not in out:
            out["reward"] = t.empty((T, self.n_envs), dtype=t.float64, device=self.dev)
        if "done" not in out:
            out["done"] = t.empty((T, self.n_envs), dtype=t.int32, device=self.dev)
        if want_rec and "rec" not in out:
            out["rec"] = t.empty((T, self.n_envs, 8), dtype=t.int32, device=self.dev)
        st = out.get("state") if want_state else None
        p64 = st.data_ptr() if st is not None and st.dtype == t.float64 else None
        p32 = st.data_ptr() if st is not None and st.dtype == t.float32 else None
        _lib.check(self._L.fjsp_vec_step(
            self._h, self._stream(), T, actions.data_ptr(), rnd.data_ptr() if rnd is not None else None,
            int(reward_policy), float(completion), float(tardiness), float(energy), int(autoreset),
            p64, p32, out["reward"].data_ptr(), out["done"].data_ptr(),
            out["rec"].data_ptr() if want_rec else None))
        return out

    def step(self, actions, rnd=None, **kw):
        """One step(action) for every copy: actions [n_envs, 2] -> (state, reward, done) tensors."""
        o = self.rollout(actions.reshape(1, self.n_envs, 2), None if rnd is None else rnd.reshape(1, self.n_envs, 2), **kw)
        return o["state"][0], o["reward"][0], o["done"][0]

    # ------------------------------------------------------------------ host API (numpy in / out)
    def step_host(self, actions, rnd=None, reward_policy=1, completion=1.0, tardiness=1.0, energy=1.0,
                  autoreset=True, want_rec=True, state_dtype=np.float64):
        actions = np.ascontiguousarray(actions, np.int32)
        if actions.ndim == 2:
            actions = actions[None]
        T = actions.shape[0]
        assert actions.shape == (T, self.n_envs, 2)
        rnd = None if rnd is None else np.ascontiguousarray(rnd, np.uint32).reshape(T, self.n_envs, 2)
        st = np.empty((T, self.n_envs, self.state_size), state_dtype)
        rw = np.empty((T, self.n_envs), np.float64)
        dn = np.empty((T, self.n_envs), np.int32)
        rec = np.empty((T, self.n_envs, 8), np.int32) if want_rec else None
        _lib.check(self._L.fjsp_vec_step_host(
            self._h, T, actions.ctypes.data, rnd.ctypes.data if rnd is not None else None, int(reward_policy),
            float(completion), float(tardiness), float(energy), int(autoreset),
            st.ctypes.data if state_dtype == np.float64 else None,
            st.ctypes.data if state_dtype == np.float32 else None,
            rw.ctypes.data, dn.ctypes.data, rec.ctypes.data if want_rec else None))
        return st, rw, dn, rec

    def reset_host(self):
        st = np.empty((self.n_envs, self.state_size), np.float64)
        _lib.check(self._L.fjsp_vec_reset_host(self._h, st.ctypes.data, None))
        return st

    def info(self):
        a = np.zeros((self.n_envs, 12), np.int64)
        _lib.check(self._L.fjsp_vec_info(self._h, a.ctypes.data))
        return {k: a[:, i] for i, k in enumerate(INFO_KEYS)}

    def slots(self):
        """Diagnostic: environment of every warp slot of the last step launch (-1 = empty),
        shaped [virtual CTAs, env warps per CTA] (the LP-aware packing rewrites it before every launch)."""
        n = self._L.fjsp_vec_slots(self._h, None, 0)
        if n < 0:
            _lib.check(n)
        a = np.zeros(n, np.int32)
        if self._L.fjsp_vec_slots(self._h, a.ctypes.data, n) < 0:
            _lib.check(-1)
        return a.reshape(-1, self.query()["env_warps"])

    def query(self):
        a = np.zeros(12, np.int64)
        _lib.check(self._L.fjsp_vec_query(self._h, a.ctypes.data))
        keys = ["n_envs", "state_size", "env_record_bytes", "instance_record_bytes", "grid", "block",
                "lp_scratch_bytes_per_slab", "launches", "env_warps", "lp_server_ctas", "n_slots", "step_smem_bytes"]
        return dict(zip(keys, (int(x) for x in a)))


class FJSPEnv:
    """One environment, the reference's surface (SO_DFJSP_Environment / MO_DFJSP_Environment)."""

    def __init__(self, use_instance=True, variant="SO_DFJSP", device=0, seed=0, instance=None, **kwargs):
        if instance is None:
            if use_instance:
                instance = FJSPInstance.generate(seed, kwargs.get("DDT", 1.0), kwargs.get("M", 10), kwargs.get("S", 3),
                                                 "DA3C")
            else:
                instance = FJSPInstance.from_csv(kwargs["path"], kwargs["file_name"],
                                                 "SO" if variant in ("SO_DFJSP", "SO_FJSSP") else "MO")
        self.instance = instance
        self.vec = FJSPVecEnv([instance], [0], variant, device)
        for k in ("actions_size", "state_size", "observation_space", "action_tuple", "action_types"):
            setattr(self, k, getattr(self.vec, k))
        self.machine_count, self.kind_count, self.order_count = instance.machine_count, instance.kind_count, instance.order_count
        self.state = self.next_state = self.reward = None
        self.done = False
        self.reward_sum = 0
        self.step_count = 0
        self._rng = np.random.default_rng(seed)

    def reset(self):
        self.state = self.vec.reset_host()[0]
        self.done, self.reward_sum, self.step_count = False, 0, 0
        return self.state

    def step(self, action, reward_policy=None, completion=None, tardiness=None, energy_consumption=None):
        if len(action) == 1:
            action = self.action_tuple[action[0]]
        if not (0 <= action[0] < self.actions_size[0]):
            raise MyError("报错：未定义该工序动作规则")
        if not (0 <= action[1] < self.actions_size[1]):
            raise MyError("报错：未定义该机器分配规则。")
        if self.vec.variant in (1, 2) and reward_policy not in (0, 1, 2, 3):
            raise MyError("未定义该回报函数")
        rnd = self._rng.integers(0, 2**32, (1, 1, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, _ = self.vec.step_host(np.array(action, np.int32).reshape(1, 1, 2), rnd,
                                           reward_policy or 0, completion or 1.0, tardiness or 0.0,
                                           energy_consumption or 1.0, autoreset=False, want_rec=False)
        self.state = self.next_state = st[0, 0]
        self.reward, self.done = float(rw[0, 0]), bool(dn[0, 0])
        self.reward_sum += self.reward
        self.step_count += 1
        return self.state, self.reward, self.done

    @property
    def completion_time(self):
        return int(self.vec.info()["completion_time"][0])

    @property
    def delay_time_sum(self):
        return int(self.vec.info()["delay_time_sum"][0])

    @property
    def energy_consumption(self):
        return int(self.vec.info()["energy_consumption"][0])
