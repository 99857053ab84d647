#!/usr/bin/env python
"""Benchmark of the AIMNet2 E+F hot path (BASELINE.json metric: atom-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg1]

One "step" = one energy+forces evaluation of one batch of synthetic input, neighbor construction included.
Default workload (N=1): cfg-2 of BASELINE.json — 1024 random 50-atom organic molecules (51 200 atoms), aimnet2 graph
with seeded random weights, Coulomb "simple" + DFT-D3, coordinates jittered every step.

Printed JSON line (rank 0):
  value      whole-job atom-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the C-ABI host-buffer entry (aimnet2_engine_eval_host): H2D of coord/numbers/
             charge/mol_idx from pinned memory + compute + D2H of energy/forces/charges, every step
  roofline   the dominant kernel class (per-atom MLP GEMMs): algorithmic FLOPs / summed GEMM launch time measured
             with CUDA events around every GEMM launch (engine timing level 2) in a separate instrumented pass
  cpu_baseline  the CPU oracle (a PyTorch-CPU restatement with the reference's computational shape) on a bounded
             sample of the same workload, all host threads
`--impl reference` times that CPU path alone (the Python reference tree cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "atom-steps/sec (E+F)"
UNIT = "atom-steps/s"
# algorithmic work per atom-step of the MLP stacks, forward + input-gradient backward
MLP_MACS_PER_ATOM = {1: 2_181_760, 2: 2_212_976}


def make_workload(name: str, seed: int):
    from aimnetcentral_b200.structures import allose_supercell, random_molecules

    if name == "cfg2":
        coord, numbers = random_molecules(1024, 50, seed=seed)
        B, n = coord.shape[:2]
        return dict(coord=coord.reshape(-1, 3), numbers=numbers.reshape(-1).astype(np.int32),
                    charge=np.zeros(B, np.float32), mol_idx=np.repeat(np.arange(B), n).astype(np.int32), cell=None,
                    desc="cfg-2: 1024 x 50-atom random organic molecules, aimnet2, E+F, Coulomb simple + DFT-D3",
                    stress=False)
    if name == "cfg3":
        z, x, cell = allose_supercell((7, 3, 5), jitter=0.02, seed=seed)
        return dict(coord=x, numbers=z.astype(np.int32), charge=np.zeros(1, np.float32), mol_idx=None, cell=cell,
                    desc="cfg-3: 10 080-atom allose supercell, PBC, DSF Coulomb + DFT-D3, E+F+stress", stress=True)
    if name == "cfg5":
        z, x, cell = allose_supercell((14, 6, 10), jitter=0.02, seed=seed)
        return dict(coord=x, numbers=z.astype(np.int32), charge=np.zeros(1, np.float32), mol_idx=None, cell=cell,
                    desc="cfg-5: 80 640-atom allose supercell, PBC, Ewald Coulomb (1e-6) + DFT-D3, E+F+stress (one "
                         "replica per GPU)", stress=True, coulomb="ewald")
    if name == "cfg1":
        g = np.load(os.path.join(ROOT, "tests", "golden", "taxol_q0.npz"))
        return dict(coord=g["in_coord"].astype(np.float32), numbers=g["in_numbers"].astype(np.int32),
                    charge=np.zeros(1, np.float32), mol_idx=None, cell=None,
                    desc="cfg-1: taxol, 113 atoms, single molecule, E+F, Coulomb simple + DFT-D3", stress=False)
    if name == "cfg4":
        coord, numbers = random_molecules(512, 80, seed=4321 + seed, box=8.5)
        B, n = coord.shape[:2]
        rng = np.random.default_rng(seed)
        charge = rng.integers(-1, 2, size=B).astype(np.float32)
        nelec = numbers.sum(axis=1) - charge.astype(np.int64)
        mult = np.where(nelec % 2 == 0, rng.choice([1, 3], size=B), 2).astype(np.float32)
        return dict(coord=coord.reshape(-1, 3), numbers=numbers.reshape(-1).astype(np.int32), charge=charge, mult=mult,
                    mol_idx=np.repeat(np.arange(B), n).astype(np.int32), cell=None, channels=2,
                    desc="cfg-4: aimnet2-nse graph (2 charge channels), 512 x 80-atom molecules, E+F+charges+spin charges",
                    stress=False)
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(workload: str, seed: int, budget_s: float = 20.0, steps: int | None = None):
    """Time the CPU oracle on a bounded sample of the workload. Returns (atom-steps/s, description, cores)."""
    import torch

    from aimnetcentral_b200 import ModelSpec, random_state_dict
    from oracle.calculator_oracle import oracle_calculate

    torch.set_num_threads(os.cpu_count() or 1)
    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    w = make_workload(workload, seed)
    if workload == "cfg2":
        nmol = 32
        sel = slice(0, nmol * 50)
        inp = dict(coord=w["coord"][sel], numbers=w["numbers"][sel], charge=w["charge"][:nmol], mol_idx=w["mol_idx"][sel])
        sample = f"{nmol} of the 1024 molecules ({nmol * 50} atoms) per step, Coulomb simple + D3, E+F (the reference's "\
                 "mode-1 path needs an N_total^2 scratch for the all-pairs list, so it is run in chunks; atom-steps/s is "\
                 "chunk-invariant)"
        kw = dict(stress=False)
    else:
        from aimnetcentral_b200.structures import allose_supercell

        z, x, cell = allose_supercell((2, 1, 1), jitter=0.02, seed=seed)
        inp = dict(coord=x, numbers=z, charge=np.zeros(1, np.float32), cell=cell)
        sample = "2x1x1 allose supercell (192 atoms), DSF + D3, E+F+stress (the torch D3 path materialises (N,M,5,5) "\
                 "temporaries, so the 10 080-atom box does not fit the time budget)"
        kw = dict(stress=True)
    natoms = len(inp["numbers"])
    rng = np.random.default_rng(seed)
    t_all = []
    oracle_calculate(sd, inp, **kw)  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        step_in = dict(inp, coord=(inp["coord"] + rng.normal(0, 0.01, inp["coord"].shape)).astype(np.float32))
        t1 = time.perf_counter()
        oracle_calculate(sd, step_in, **kw)
        t_all.append(time.perf_counter() - t1)
        n += 1
        if steps is not None:
            if n >= steps:
                break
        elif time.perf_counter() - t0 > budget_s or n >= 50:
            break
    dt = float(np.median(t_all))
    return natoms / dt, sample, torch.get_num_threads(), dt, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.seed)
    for _ in range(max(0, args.warmup - 1)):
        pass
    value, sample, cores, dt, n = cpu_baseline(args.workload, args.seed, steps=max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--gemm-backend", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from aimnetcentral_b200 import ModelSpec, random_state_dict
    from aimnetcentral_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    w = make_workload(args.workload, args.seed + rank)  # weak scaling: every rank owns its own batch
    spec = ModelSpec(num_charge_channels=w.get("channels", 1))
    sd = random_state_dict(0, spec)
    eng = Engine(sd, spec.C, dev)
    if args.gemm_backend is not None:
        eng.set_gemm_backend(args.gemm_backend)
    pbc = w["cell"] is not None
    eng.set_options(coulomb_method=w.get("coulomb", "dsf" if pbc else "simple"), dispersion=True)
    N = len(w["numbers"])
    B = len(w["charge"])
    rng = np.random.default_rng(args.seed + 17 * rank)
    n_sets = W + K
    # host (pinned) and device copies of every step's jittered coordinates
    coords_h = [torch.from_numpy((w["coord"] + rng.normal(0, 0.01, w["coord"].shape)).astype(np.float32)).pin_memory()
                for _ in range(n_sets)]
    coords_d = [c.to(dev) for c in coords_h]
    numbers_h = torch.from_numpy(w["numbers"]).pin_memory()
    charge_h = torch.from_numpy(w["charge"]).pin_memory()
    mol_h = torch.from_numpy(w["mol_idx"]).pin_memory() if w["mol_idx"] is not None else None
    numbers_d, charge_d = numbers_h.to(dev), charge_h.to(dev)
    mult_h = torch.from_numpy(w["mult"]).pin_memory() if w.get("mult") is not None else None
    mult_d = mult_h.to(dev) if mult_h is not None else None
    mol_d = mol_h.to(dev) if mol_h is not None else None
    cell_d = torch.from_numpy(w["cell"]).to(dev) if pbc else None
    gather_e = gather_f = None
    if world > 1:
        gather_e = torch.empty(world * B, dtype=torch.float64, device=dev)
        gather_f = torch.empty(world * N, 3, dtype=torch.float32, device=dev)

    def step(i):
        out = eng.eval(coords_d[i], numbers_d, charge_d, mol_idx=mol_d, mult=mult_d, cell=cell_d, forces=True,
                       stress=w["stress"])
        if world > 1:  # result gather of the batch split (NCCL over NVLink)
            dist.all_gather_into_tensor(gather_e, out["energy"])
            dist.all_gather_into_tensor(gather_f, out["forces"])
        return out

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(W):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        step(W + i)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.last_launches() * K
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * N * K / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + compute + D2H inside every step) ----
    out_h = {"energy": torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
             "charges": torch.empty(N, dtype=torch.float32).pin_memory().numpy(),
             "forces": torch.empty(N, 3, dtype=torch.float32).pin_memory().numpy()}
    if w["stress"]:
        out_h["stress"] = torch.empty(3, 3, dtype=torch.float32).pin_memory().numpy()
    if spec.C == 2:
        out_h["spin_charges"] = torch.empty(N, dtype=torch.float32).pin_memory().numpy()

    def step_host(i):
        eng.eval_host(coords_h[i], numbers_h, charge_h, mol_idx=mol_h, mult=mult_h, cell=w["cell"], forces=True,
                      stress=w["stress"], out=out_h)

    for i in range(2):
        step_host(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(W + i)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * K / float(t.item())
    h2d = coords_h[0].numel() * 4 + numbers_h.numel() * 4 + charge_h.numel() * 4 + (mol_h.numel() * 4 if mol_h is not None else 0) + (36 if pbc else 0)
    d2h = sum(int(v.nbytes) for v in out_h.values())
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel class (instrumented pass, not part of the timed region) ----
    eng.enable_timing(2)
    gemm_ms, tot_ms, phases = [], [], []
    step(0)
    torch.cuda.synchronize(dev)
    for i in range(5):
        step(i)
        torch.cuda.synchronize(dev)
        tm = eng.last_timing()
        gemm_ms.append(tm["gemm_ms"])
        tot_ms.append(tm["total_ms"])
        phases.append(tm)
    eng.enable_timing(0)
    gms = float(np.median(gemm_ms))
    n_gemm = int(phases[-1]["gemm_launches"])
    flops = 2.0 * 2.0 * MLP_MACS_PER_ATOM[spec.C] * N  # fwd + dgrad
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    # three error-compensating MMAs per product; kind::f16 runs at the bf16 rate, kind::tf32 at half of it
    peak_tf = bf16 / 3.0 if eng.gemm_backend in (2, 3) else bf16 / 2.0 / 3.0
    achieved = flops / (gms * 1e-3) / 1e12
    roofline = {"kernel": "gemm_nt (per-atom MLP stacks, %d launches/step)" % n_gemm, "bound": "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                # DRAM bytes of the largest GEMM launch (51 200 x 512 x 704, GELU + pre-split output) from the committed
                # `ncu --set full` capture profiles/r1_prof_gemm_tc16_summary.csv: 159 MB read + 163 MB written, against
                # 144 MB (A hi+lo) + 210 MB (y hi+lo, gelu') algorithmic -- no re-reads; null for the other workloads
                "traffic": 321.8e6 if (args.workload == "cfg2" and eng.gemm_backend == 2) else None,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained") +
                               (" / 3 (3xFP16 split on the kind::f16 pipe)" if eng.gemm_backend in (2, 3) else
                                " / 2 (tf32) / 3 (3xTF32 split)"),
                "gemm_ms_per_step": gms, "gemm_share_of_step": gms / (ms_max / K),
                "phase_ms": {k: float(np.median([p[k] for p in phases])) for k in
                             ("neighbors_ms", "forward_ms", "pair_terms_ms", "backward_ms", "total_ms")}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "atoms_per_gpu": N, "molecules_per_gpu": B,
                       "weights": "seeded random, aimnet2 architecture (2.2M params)",
                       "l2": "per-step working set (activations + saved tensors, >1 GB) exceeds the 126 MB L2; "
                             "coordinates change every step",
                       "gemm_backend": {2: "tcgen05-3xfp16-rowchunk-scaled", 3: "tcgen05-3xfp16-rowchunk-scaled-pipelined-epilogue (experimental)",
                                        1: "tcgen05-3xtf32"}.get(eng.gemm_backend, "simt-fp32"),
                       "multi_gpu": "independent batches per rank + all_gather of energy/forces (NCCL)" if world > 1 else "single"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline and args.workload in ("cfg2", "cfg3"):
            v, sample, cores, dtc, n = cpu_baseline(args.workload, args.seed, budget_s=15.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "s_per_step": dtc, "steps": n}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
