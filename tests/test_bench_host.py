"""Host-side pieces of bench.py that run without a GPU: the five BASELINE workloads build, their action
distributions match the variants' rule counts, the reference arm prints a line of the contract's shape, and the
NUMA binding never raises and never widens the CPU set."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("name", sorted(bench.CONFIGS))
def test_config_workloads_build(name):
    cfg = dict(bench.CONFIGS[name])
    cfg["envs"] = min(cfg["envs"], 24)
    if "distinct" in cfg:
        cfg["distinct"] = min(cfg["distinct"], 8)
    blobs, env_inst = bench.config_blobs(cfg, 2026, 0)
    assert len(env_inst) == cfg["envs"] and 0 <= env_inst.min() and env_inst.max() < len(blobs)
    a, r = bench.make_actions(np.random.default_rng(0), 5, cfg["envs"], cfg["variant"])
    nt, nm = (12, 10) if cfg["variant"].startswith("MO") else (6, 5)
    assert a.shape == (5, cfg["envs"], 2) and a[..., 0].max() < nt and a[..., 1].max() < nm and a.min() >= 0
    assert r.shape == (5, cfg["envs"], 2) and r.dtype == np.uint32
    # another rank plays another shard of generated instances (Brandimarte: the same ten files)
    blobs1, _ = bench.config_blobs(cfg, 2026, 1)
    if not cfg.get("brandimarte"):
        assert any(not np.array_equal(x, y) for x, y in zip(blobs, blobs1))


def test_numa_binding_is_safe():
    before = os.sched_getaffinity(0)
    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: types.SimpleNamespace(
        pci_domain_id=0, pci_bus_id=0xfe, pci_device_id=0x1f)))
    info = bench.bind_to_gpu_numa_node(fake, 0)
    try:
        assert isinstance(info, dict) and "bound" in info
        assert os.sched_getaffinity(0) <= before
        if not info["bound"]:
            assert info.get("why")
    finally:
        os.sched_setaffinity(0, before)

    def boom(i):
        raise RuntimeError("no device")
    info = bench.bind_to_gpu_numa_node(types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=boom)), 0)
    assert info["bound"] is False and os.sched_getaffinity(0) == before


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "so_single",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == bench.CONFIGS["so_single"]["workload"]
