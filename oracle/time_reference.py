"""TEST INFRASTRUCTURE ONLY -- times the UNMODIFIED Python reference environments (imported from
/root/reference through oracle/refshim; CPLEX is absent, so its LP is served by the shim: oracle/fjsp_lp.c or
scipy's HiGHS) on one host core, for BASELINE.md.  Runs in the build container only.
    python oracle/time_reference.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
import ref_loader  # noqa: E402
from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance  # noqa: E402


def time_case(name, variant, inst, steps, seed=1):
    inst.ddt = float(int(inst.ddt))
    nt, nm = make_golden.NRULES[variant]
    rng = np.random.default_rng(seed)
    actions = np.stack([rng.integers(0, nt, steps + 8), rng.integers(0, nm, steps + 8)], 1).astype(np.int32)
    rnd = rng.integers(0, 2**32, (steps + 8, 2), dtype=np.uint64).astype(np.uint32)
    t0 = time.perf_counter()
    ref, T = make_golden.run_reference(variant, inst, actions, rnd, 1, 1, steps)
    dt = time.perf_counter() - t0
    print("%-34s %-18s KT=%3d M=%2d S=%d: %5d steps (incl. reset and %d LP solves) in %6.2f s = %7.1f steps/s" % (
        name, variant, len(inst.kind_task_tuple), inst.machine_count, inst.order_count, T, inst.order_count, dt, T / dt))
    return T / dt


def main():
    ref_data = os.path.join(ref_loader.REFERENCE_ROOT, "data")
    G = FJSPInstance.generate
    time_case("data/DA3C/DDT0.5_M10_S3", "SO_DFJSP", FJSPInstance.from_csv(os.path.join(ref_data, "DA3C"), "DDT0.5_M10_S3", "SO"), 600)
    time_case("data/HMPSAC/DDT1.0_M10_S3", "MO_DFJSP", FJSPInstance.from_csv(os.path.join(ref_data, "HMPSAC"), "DDT1.0_M10_S3", "MO"), 300)
    time_case("bench instance (generated, M10 S3)", "MO_DFJSP", G(2026 * 100003, 0.5, 10, 3, "DA3C"), 600)
    time_case("Brandimarte Mk01", "SO_DFJSP", FJSPInstance.from_csv(os.path.join(ref_data, "benchmark", "Brandimarte_Data"), "Mk01", "SO"), 55)
    time_case("generated M20 S5 (HMPSAC profile)", "MO_DFJSP", G(77, 1.0, 20, 5, "HMPSAC"), 300)


if __name__ == "__main__":
    main()
