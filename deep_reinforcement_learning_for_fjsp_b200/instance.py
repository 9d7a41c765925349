"""FJSP problem instances: the reference's CSV formats, its random generators, and
the flat int32 "instance blob" the C-ABI consumes.

Mirrors (reference file:line)
  environments/SO_DFJSP_instance_read.py:6-89   Data  (based/process/order csv)
  environments/MO_DFJSP_instance_read.py:6-113  Data  (+ power column, machine_data.csv
                                                       with idle power and breakdowns)
  environments/Instance_generate.py:19-94       Instance (DA3C profile)
  data/generated_HMPSAC.py:16-92                Instance (HMPSAC profile)

Blob layout (int32 words; include/fjsp_b200.h documents the same table):
  [0] magic 'FJSP' 0x464A5350   [1] total words   [2] M   [3] K   [4] KT   [5] S
  [6] NP eligible (machine, operation-type) pairs   [7] NBD breakdown intervals
  [8],[9] the DDT state feature as float64 bits (lo, hi)
  [10] Nmax = max jobs of one kind over all orders   [11] NJ total jobs   [12..15] 0
  ntask[K] rj_kind[KT] rj_stage[KT] nelig[KT] mt_order[KT*M] ptime[KT*M] power[KT*M]
  idle_power[M] arrive[S] due[S] count[S*K] bd_ptr[M+1] bd_start[NBD] bd_end[NBD]
  pair_order[NP]
`mt_order` keeps each operation type's machine tuple in FILE order and `pair_order`
is the iteration order of the reference's `var_list` set (class_FJSP.py:263); both
are needed to reproduce CPython's tie-breaking.
"""
from __future__ import annotations

import csv
import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = 0x464A5350
MAX_MACHINES = 32
MAX_OPTYPES = 256

RJ = Tuple[int, int]


def _ints(s: str) -> Tuple[int, ...]:
    return tuple(int(i) for i in re.findall(r"\d+", s))


@dataclass
class FJSPInstance:
    machine_count: int
    ntask: List[int]                               # operations per job kind
    machine_rj: Dict[RJ, Tuple[int, ...]]          # (r, j) -> machines, file order
    time_rjm: Dict[RJ, Dict[int, int]]             # (r, j) -> {m: processing time}
    arrive: List[int]
    due: List[int]
    count: List[Tuple[int, ...]]                   # count[s][r]
    ddt: float = 0.0                               # the value the reference exposes as state[0] (MO)
    power_rjm: Optional[Dict[RJ, Dict[int, int]]] = None
    idle_power: Optional[List[int]] = None
    breakdowns: Dict[int, List[Tuple[int, int]]] = field(default_factory=dict)
    name: str = ""

    # ------------------------------------------------------------------ shape
    @property
    def kind_count(self) -> int:
        return len(self.ntask)

    @property
    def order_count(self) -> int:
        return len(self.arrive)

    @property
    def kind_task_tuple(self) -> Tuple[RJ, ...]:
        return tuple((r, j) for r in range(self.kind_count) for j in range(self.ntask[r]))

    @property
    def total_operations(self) -> int:
        return sum(self.count[s][r] * self.ntask[r] for s in range(self.order_count)
                   for r in range(self.kind_count))

    def validate(self) -> None:
        M = self.machine_count
        if not (1 <= M <= MAX_MACHINES):
            raise ValueError(f"machine_count {M} outside 1..{MAX_MACHINES}")
        if len(self.kind_task_tuple) > MAX_OPTYPES:
            raise ValueError("too many operation types")
        for rj in self.kind_task_tuple:
            ms = self.machine_rj[rj]
            if len(ms) == 0 or len(set(ms)) != len(ms):
                raise ValueError(f"operation type {rj}: empty or duplicated machine tuple")
            for m in ms:
                if not (0 <= m < M) or self.time_rjm[rj][m] <= 0:
                    raise ValueError(f"operation type {rj}: bad machine/time")
        if any(len(c) != self.kind_count for c in self.count):
            raise ValueError("count rows must have one entry per kind")
        if any(x <= 0 for c in self.count for x in c):
            raise ValueError("every order must hold at least one job of every kind")

    # ------------------------------------------------------------------ csv in
    @classmethod
    def from_csv(cls, path: str, file_name: str, fmt: str = "SO") -> "FJSPInstance":
        """fmt='SO': SO_DFJSP_instance_read.Data; fmt='MO': MO_DFJSP_instance_read.Data."""
        base = os.path.join(path, file_name)

        def rows(name):
            with open(os.path.join(base, name), "r") as f:
                return [r for r in csv.reader(f)]

        b = rows("based_data.csv")
        kind_count, machine_count, order_count = (_ints(b[1][i])[0] for i in range(3))
        # reference quirk (…instance_read.py:36-39,53): DDT is read with the int regex,
        # so "0.5" becomes 0 and "1.5" becomes 1.
        ddt = float(_ints(b[1][3])[0]) if len(b[1]) > 3 else 0.0
        arrive, due, count = [0] * order_count, [0] * order_count, [()] * order_count
        for row in rows("order_data.csv")[1:]:
            s = _ints(row[0])[0]
            arrive[s], due[s], count[s] = _ints(row[1])[0], _ints(row[2])[0], _ints(row[3])
        ntask = [0] * kind_count
        machine_rj, time_rjm, power_rjm = {}, {}, {}
        for row in rows("process_data.csv")[1:]:
            r, j = _ints(row[0])[0], _ints(row[1])[0]
            ms, ts = _ints(row[2]), _ints(row[3])
            ntask[r] = max(ntask[r], j + 1)
            machine_rj[(r, j)] = ms
            time_rjm[(r, j)] = dict(zip(ms, ts))
            if fmt == "MO":
                power_rjm[(r, j)] = dict(zip(ms, _ints(row[4])))
        idle_power, breakdowns = None, {}
        if fmt == "MO":
            idle_power = [None] * machine_count
            breakdowns = {m: [] for m in range(machine_count)}
            for row in rows("machine_data.csv")[1:]:
                m = _ints(row[0])[0]
                if idle_power[m] is None:
                    idle_power[m] = _ints(row[1])[0]
                if len(row) > 2:
                    breakdowns[m].append((_ints(row[2])[0], _ints(row[3])[0]))
        inst = cls(machine_count, ntask, machine_rj, time_rjm, arrive, due, count, ddt,
                   power_rjm if fmt == "MO" else None, idle_power, breakdowns, file_name)
        inst.validate()
        return inst

    # ------------------------------------------------------------------ csv out
    def write_csv(self, path: str, file_name: str, fmt: str = "SO", ddt_text: Optional[str] = None) -> str:
        """Write the instance in the reference's csv layout (so the unmodified reference
        can load it with use_instance=False).  `ddt_text` is what goes in the DDT cell;
        the default reproduces self.ddt through the reference's integer regex."""
        base = os.path.join(path, file_name)
        os.makedirs(base, exist_ok=True)
        if ddt_text is None:
            ddt_text = str(int(self.ddt))
        with open(os.path.join(base, "based_data.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kind_count", "machine_count", "order_count", "DDT"])
            w.writerow([self.kind_count, self.machine_count, self.order_count, ddt_text])
        with open(os.path.join(base, "order_data.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["order", "time_arrive", "time_delivery", "kind_number"])
            for s in range(self.order_count):
                w.writerow([s, self.arrive[s], self.due[s], tuple(int(x) for x in self.count[s])])
        with open(os.path.join(base, "process_data.csv"), "w", newline="") as f:
            w = csv.writer(f)
            head = ["kind", "task", "machine_selectable", "process_time"]
            w.writerow(head + (["power"] if fmt == "MO" else []))
            for (r, j) in self.kind_task_tuple:
                ms = tuple(int(m) for m in self.machine_rj[(r, j)])
                row = [r, j, ms, tuple(int(self.time_rjm[(r, j)][m]) for m in ms)]
                if fmt == "MO":
                    row.append(tuple(int(self.power_rjm[(r, j)][m]) for m in ms))
                w.writerow(row)
        if fmt == "MO":
            with open(os.path.join(base, "machine_data.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["machine", "idle_power", "breakdown_start", "breakdown_end"])
                for m in range(self.machine_count):
                    bds = self.breakdowns.get(m, [])
                    if not bds:
                        w.writerow([m, self.idle_power[m]])
                    for (bs, be) in bds:
                        w.writerow([m, self.idle_power[m], bs, be])
        return base

    # ------------------------------------------------------------------ generators
    @classmethod
    def generate(cls, seed: int, DDT: float = 1.0, M: int = 10, S: int = 3, profile: str = "DA3C",
                 breakdowns: bool = False, scale: float = 1.0) -> "FJSPInstance":
        """Random instance with the reference generators' distributions.
        profile 'DA3C'  : Instance_generate.py (3-12 kinds, 3-5 ops, t 40-400, 5-50 jobs/kind/order)
        profile 'HMPSAC': data/generated_HMPSAC.py (5-15 kinds, 5-10 ops, t 1-20, 5-10 jobs)
        `scale` shrinks the job counts (for small test instances).  The draw order
        differs from the reference's Mersenne-Twister stream: instances are inputs,
        parity is judged on what the environments do with the SAME instance."""
        rng = np.random.default_rng(seed)
        if profile == "DA3C":
            kr, jr, tr, nr = (3, 12), (3, 5), (40, 400), (5, 50)
        elif profile == "HMPSAC":
            kr, jr, tr, nr = (5, 15), (5, 10), (1, 20), (5, 10)
        else:
            raise ValueError(profile)
        while True:
            K = int(rng.integers(kr[0], kr[1] + 1))
            ntask = [int(rng.integers(jr[0], jr[1] + 1)) for _ in range(K)]
            machine_rj, time_rjm, power_rjm = {}, {}, {}
            for r in range(K):
                for j in range(ntask[r]):
                    k = int(rng.integers(1, M + 1))
                    ms = tuple(int(m) for m in rng.permutation(M)[:k])
                    machine_rj[(r, j)] = ms
                    time_rjm[(r, j)] = {m: int(rng.integers(tr[0], tr[1] + 1)) for m in ms}
                    power_rjm[(r, j)] = {m: int(rng.integers(10, 201)) for m in ms}
            used = {m for ms in machine_rj.values() for m in ms}
            if len(used) == M:   # the reference divides by len(kind_task_tuple) of every machine
                break
        lo = max(1, int(round(nr[0] * scale)))
        hi = max(lo, int(round(nr[1] * scale)))
        count = [tuple(int(rng.integers(lo, hi + 1)) for _ in range(K)) for _ in range(S)]
        time_rj = {rj: sum(time_rjm[rj].values()) / len(time_rjm[rj]) for rj in machine_rj}
        gaps = [sum(time_rj[(r, j)] * count[s][r] for r in range(K) for j in range(ntask[r])) * DDT / (M * 2)
                for s in range(S)]
        intervals = [0.0] + [float(rng.uniform(100, 200)) * max(scale, 0.05) for _ in range(S - 1)]
        arrive = [int(sum(intervals[: s + 1])) for s in range(S)]
        due = [int(x) for x in sorted(arrive[s] + gaps[s] for s in range(S))]
        idle_power = [int(rng.integers(1, 10)) for _ in range(M)]
        bds: Dict[int, List[Tuple[int, int]]] = {m: [] for m in range(M)}
        if breakdowns:
            horizon = max(due) * 2 + 100
            for m in range(M):
                t = 0
                for _ in range(int(rng.integers(0, 4))):
                    t += int(rng.integers(max(2, horizon // 8), max(3, horizon // 2)))
                    length = int(rng.integers(1, max(2, horizon // 20)))
                    bds[m].append((t, t + length))
                    t += length
        inst = cls(M, ntask, machine_rj, time_rjm, arrive, due, count, float(DDT), power_rjm,
                   idle_power, bds, f"gen_{profile}_s{seed}_DDT{DDT}_M{M}_S{S}")
        inst.validate()
        return inst

    # ------------------------------------------------------------------ blob
    def pair_order(self) -> List[Tuple[int, RJ]]:
        """Iteration order of the reference's decision-variable set (class_FJSP.py:263),
        built exactly as the reference builds it so CPython lays the table out the same."""
        machine_tuple = tuple(range(self.machine_count))
        kt = self.kind_task_tuple
        kind_task_m_dict = {m: tuple(rj for rj in kt if m in self.machine_rj[rj]) for m in machine_tuple}
        var_list = {(m, (r, j)) for m in machine_tuple for (r, j) in kind_task_m_dict[m]}
        return list(var_list)

    def to_blob(self) -> np.ndarray:
        self.validate()
        M, K, S = self.machine_count, self.kind_count, self.order_count
        kt = self.kind_task_tuple
        KT = len(kt)
        index = {rj: q for q, rj in enumerate(kt)}
        mt = np.full((KT, M), -1, np.int32)
        pt = np.zeros((KT, M), np.int32)
        pw = np.zeros((KT, M), np.int32)
        nelig = np.zeros(KT, np.int32)
        for rj, q in index.items():
            ms = self.machine_rj[rj]
            nelig[q] = len(ms)
            for i, m in enumerate(ms):
                mt[q, i] = m
                pt[q, m] = self.time_rjm[rj][m]
                if self.power_rjm is not None:
                    pw[q, m] = self.power_rjm[rj][m]
        idle = np.array(self.idle_power if self.idle_power is not None else [0] * M, np.int32)
        bd_ptr, bd_s, bd_e = [0], [], []
        for m in range(M):
            for (a, b) in self.breakdowns.get(m, []):
                bd_s.append(a)
                bd_e.append(b)
            bd_ptr.append(len(bd_s))
        po = np.array([index[rj] * M + m for (m, rj) in self.pair_order()], np.int32)
        njobs = [sum(self.count[s][r] for s in range(S)) for r in range(K)]
        lo, hi = struct.unpack("<ii", struct.pack("<d", float(self.ddt)))
        header = np.zeros(16, np.int32)
        header[[0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]] = [MAGIC, M, K, KT, S, len(po), len(bd_s), lo, hi,
                                                      max(njobs), sum(njobs)]
        parts = [header, np.array(self.ntask, np.int32),
                 np.array([r for r, _ in kt], np.int32), np.array([j for _, j in kt], np.int32),
                 nelig, mt.ravel(), pt.ravel(), pw.ravel(), idle,
                 np.array(self.arrive, np.int32), np.array(self.due, np.int32),
                 np.array(self.count, np.int32).ravel(), np.array(bd_ptr, np.int32),
                 np.array(bd_s, np.int32), np.array(bd_e, np.int32), po]
        blob = np.concatenate([np.asarray(p, np.int32).ravel() for p in parts]).astype(np.int32)
        blob[1] = blob.size
        return blob
