import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
from deep_reinforcement_learning_for_fjsp_b200 import _lib
cfg = dict(bench.CONFIGS["mo_4096"])
blobs, env_inst = bench.config_blobs(cfg, 2026, 0)
vec = FJSPVecEnv(None, env_inst, "MO_DFJSP", device=0, blobs=blobs)
vec.reset()
B, T = 4096, 32
MODE = sys.argv[1] if len(sys.argv) > 1 else ''
if 'burn' in MODE:
    dev = torch.device('cuda', 0)
    r0 = np.random.default_rng(3)
    for i in range(64):
        a, r = bench.make_actions(r0, T, B, 'MO_DFJSP')
        vec.rollout(torch.from_numpy(a).to(dev), torch.from_numpy(r.view(np.int32)).to(dev), reward_policy=1, state_dtype=torch.float32)
    torch.cuda.synchronize()
if 'flush' in MODE:
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    if 'nofill' not in MODE:
        flush.fill_(1)
    torch.cuda.synchronize()
    if 'free' in MODE:
        del flush
        torch.cuda.empty_cache()
if 'small' in MODE:
    small = torch.empty(1024, dtype=torch.uint8, device='cuda'); small.fill_(1); torch.cuda.synchronize()

rng = np.random.default_rng(1)
ha = [torch.from_numpy(bench.make_actions(rng, T, B, "MO_DFJSP")[0]).pin_memory() for _ in range(2)]
hr = [torch.from_numpy(bench.make_actions(rng, T, B, "MO_DFJSP")[1].view(np.int32)).pin_memory() for _ in range(2)]
hs = [torch.empty((T, B, 30), dtype=torch.float32).pin_memory() for _ in range(2)]
hrw = [torch.empty((T, B), dtype=torch.float64).pin_memory() for _ in range(2)]
hdn = [torch.empty((T, B), dtype=torch.int32).pin_memory() for _ in range(2)]
L = vec._L
def begin(k):
    _lib.check(L.fjsp_vec_step_host_begin(vec._h, T, ha[k].data_ptr(), hr[k].data_ptr(), 1, 1.0, 1.0, 1.0, 1, None, None if 'nostate' in MODE else hs[k].data_ptr(), hrw[k].data_ptr(), hdn[k].data_ptr(), None))
for i in range(30):
    _lib.check(L.fjsp_vec_step_host(vec._h, T, ha[i%2].data_ptr(), hr[i%2].data_ptr(), 1, 1.0, 1.0, 1.0, 1, None, hs[i%2].data_ptr(), hrw[i%2].data_ptr(), hdn[i%2].data_ptr(), None))
torch.cuda.synchronize()
tb, tw = [], []
t0 = time.perf_counter(); begin(0); tb.append(time.perf_counter() - t0)
for i in range(1, 12):
    t0 = time.perf_counter(); begin(i % 2); tb.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); _lib.check(L.fjsp_vec_step_host_wait(vec._h)); tw.append(time.perf_counter() - t0)
t0 = time.perf_counter(); _lib.check(L.fjsp_vec_step_host_wait(vec._h)); tw.append(time.perf_counter() - t0)
print(MODE, 'per call ms', round((sum(tb) + sum(tw)) / len(tb) * 1e3, 3))
print("begin ms", [round(x * 1e3, 3) for x in tb])
print("wait  ms", [round(x * 1e3, 3) for x in tw])

# ---- the step launch alone: outputs in device memory vs written by the kernel into page-locked host memory
dev = torch.device('cuda', 0)
da = torch.from_numpy(bench.make_actions(rng, T, B, 'MO_DFJSP')[0]).to(dev)
dr = torch.from_numpy(bench.make_actions(rng, T, B, 'MO_DFJSP')[1].view(np.int32)).to(dev)
ds = torch.empty((T, B, 30), dtype=torch.float32, device=dev)
drw = torch.empty((T, B), dtype=torch.float64, device=dev)
ddn = torch.empty((T, B), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(dev)
import ctypes
for name, (ps, prw, pdn) in (("device outputs", (ds.data_ptr(), drw.data_ptr(), ddn.data_ptr())),
                             ("host-mapped outputs", (hs[0].data_ptr(), hrw[0].data_ptr(), hdn[0].data_ptr())),
                             ("host-mapped state only", (hs[0].data_ptr(), drw.data_ptr(), ddn.data_ptr()))):
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        _lib.check(L.fjsp_vec_step(vec._h, ctypes.c_void_p(st.cuda_stream), T, da.data_ptr(), dr.data_ptr(), 1, 1.0, 1.0, 1.0, 1, None, ps, prw, pdn, None))
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    print("step launch (flag + pack + step kernel),", name, "ms:", ts)
