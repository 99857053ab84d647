#!/usr/bin/env python
"""Benchmark of the AIMNet2 E+F hot path (BASELINE.json metric: atom-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg1..cfg5]
                    [--scaling weak|strong] [--no-extra] [--no-cpu-baseline] [--gemm-backend B]

One "step" = one energy+forces evaluation of one batch of synthetic input, neighbor construction included.
Default workload: cfg-2 of BASELINE.json — 1024 random 50-atom organic molecules (51 200 atoms), aimnet2 graph with
seeded random weights, Coulomb "simple" + DFT-D3, coordinates jittered every step.

Printed JSON line (rank 0):
  value       whole-job atom-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks.  N = 1: the C-ABI call with
              device pointers (aimnet2_engine_eval).  N > 1: the product's multi-GPU module (ShardedCalculator around
              AIMNet2Calculator): contiguous atom-balanced molecule shards, one fused NCCL all_gather of all outputs per step
              inside the timed region; weak scaling (1024 molecules per GPU) unless --scaling strong.
  e2e         the same metric through the C-ABI host-buffer entry (aimnet2_engine_eval_host): H2D of coord / numbers /
              charge / mol_idx from pinned memory + compute + D2H of energy / forces / charges, every step
  api         the same through AIMNet2Calculator.__call__ with numpy inputs and a copy of every output tensor into pinned
              host buffers (the front door, as an MD / screening driver uses it)
  roofline    the kernel class that takes the largest share of the step, `roofline_classes` both classes (per-atom MLP
              GEMMs on the tensor pipe, AEV / conv_sv on the fp32 FMA pipe) from CUDA events around every launch of the
              class in a separate instrumented pass, `step_roofline` the t_roof / t_measured of SURVEY.md section 8(d)
  cpu_baseline  the UNMODIFIED reference modules (oracle/_ref, vendored by oracle/make_ref.py) on the host cores on a
              bounded sample of the same workload (kind "reference"), else the oracle port (kind "port")
  extra_workloads  (N = 1) short runs of cfg-1 / cfg-3 / cfg-4 / cfg-5 in the same process, each with its own clocks
`--impl reference` times the reference's own CPU path alone on this arm's config / metric / steps / warmup.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from aimnetcentral_b200.structures import benchmark_workload as make_workload  # noqa: E402

METRIC = "atom-steps/sec (E+F)"
UNIT = "atom-steps/s"
# algorithmic work per atom-step
MLP_MACS_PER_ATOM = {1: 2_181_760, 2: 2_212_976}          # fwd; x2 for the input-gradient backward, x2 FLOP per MAC
CONV_FLOP_PER_PAIR = {1: 19_200, 2: 19_968}               # 3 passes x 2 x (3 A G 4 + 2 C G 4): fwd + grad_a + grad_g
AGH_FLOP_PER_ATOM = 0.17e6
REF_SAMPLE_MOLS = 64                                      # CPU sample of the molecule workloads (>= 64)


def workload_config(w: dict, n_atoms: int, n_mol: int) -> dict:
    """Identical in both arms (the driver compares the arms' `config`)."""
    return {"workload": w["desc"], "atoms_per_gpu": int(n_atoms), "molecules_per_gpu": int(n_mol),
            "weights": "seeded random, aimnet2 architecture (2.2M params)",
            "l2": "per-step working set (activations + saved tensors, > 1 GB) exceeds the 126 MB L2; coordinates change "
                  "every step"}


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
        return self

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


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref) or, when it is not vendored, the oracle port
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample(workload: str, seed: int):
    """Bounded sample of the workload for the host cores: (workload, inputs, kwargs, description)."""
    w = make_workload(workload, seed)
    if w["mol_idx"] is not None:   # molecule batches: the first REF_SAMPLE_MOLS molecules (molecules are independent)
        B = len(w["charge"])
        nmol = min(REF_SAMPLE_MOLS, B)
        a1 = int(np.searchsorted(w["mol_idx"], nmol))
        inp = dict(coord=w["coord"][:a1], numbers=w["numbers"][:a1], charge=w["charge"][:nmol],
                   mol_idx=w["mol_idx"][:a1].astype(np.int64))
        if w.get("mult") is not None:
            inp["mult"] = w["mult"][:nmol]
        sample = (f"{nmol} of the {B} molecules ({a1} atoms) per step; the reference's mode-1 path needs an N_total^2 "
                  "scratch for the all-pairs list, so the batch is run as a chunk (atom-steps/s is chunk-invariant)")
        return w, inp, dict(stress=False), sample
    if workload == "cfg1":
        return w, dict(coord=w["coord"], numbers=w["numbers"], charge=w["charge"]), dict(stress=False), "the whole system"
    from aimnetcentral_b200.structures import allose_supercell

    z, x, cell = allose_supercell((2, 1, 1), jitter=0.02, seed=seed)
    inp = dict(coord=x, numbers=z, charge=np.zeros(1, np.float32), cell=cell)
    sample = ("2x1x1 allose supercell (192 atoms), DSF + D3, E+F+stress (the torch D3 path materialises (N,M,5,5) "
              "temporaries, so the full box does not fit the time budget)")
    return w, inp, dict(stress=True), sample


def cpu_runner(workload: str, seed: int):
    """Returns (step(coord) callable, inputs, sample description, kind, cores)."""
    import torch

    from aimnetcentral_b200 import ModelSpec, random_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w, inp, kw, sample = cpu_sample(workload, seed)
    spec = ModelSpec(num_charge_channels=w.get("channels", 1))
    sd = random_state_dict(0, spec)
    kind = "port"
    try:
        from oracle import ref_harness as rh

        if not rh.reference_available():
            raise RuntimeError("no reference tree (oracle/_ref is made by `python -m oracle.make_ref`)")
        calc = rh.build_reference_calculator(sd, spec)
        if "cell" in inp:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                calc.set_lrcoulomb_method("dsf")
        kind = "reference"

        def step(coord):
            return rh.run_reference(calc, dict(inp, coord=coord), forces=True, stress=kw["stress"])
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] reference modules unavailable ({ex}); timing the oracle port", file=sys.stderr)
        from oracle.calculator_oracle import oracle_calculate

        def step(coord):
            return oracle_calculate(sd, dict(inp, coord=coord), num_charge_channels=spec.C, **kw)
    return step, inp, sample, kind, torch.get_num_threads()


def time_cpu(workload: str, seed: int, warmup: int, steps: int | None, budget_s: float):
    step, inp, sample, kind, cores = cpu_runner(workload, seed)
    rng = np.random.default_rng(seed)
    jit = lambda: (inp["coord"] + rng.normal(0, 0.01, inp["coord"].shape)).astype(np.float32)   # noqa: E731
    for _ in range(max(1, warmup)):
        step(jit())
    ts = []
    t0 = time.perf_counter()
    while True:
        c = jit()
        t1 = time.perf_counter()
        step(c)
        ts.append(time.perf_counter() - t1)
        if steps is not None:
            if len(ts) >= steps:
                break
        elif time.perf_counter() - t0 > budget_s or len(ts) >= 50:
            break
    dt = float(np.mean(ts))
    n = len(inp["numbers"])
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "s_per_step": dt, "steps": len(ts),
            "sample_atoms": n}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.seed)
    W, K = max(3, args.warmup), max(1, args.steps)
    # bounded: every step is the CPU sample of the workload
    cb = time_cpu(args.workload, args.seed, warmup=W, steps=K, budget_s=240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, len(w["numbers"]), len(w["charge"])),
        "device": "cpu", "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_atoms")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def mean_neighbors(w, dev, cutoff):
    """Mean row length of the full neighbor list at `cutoff` (for the algorithmic-work figures)."""
    import torch

    from aimnetcentral_b200 import ops

    x = torch.as_tensor(w["coord"], device=dev)
    kw = {}
    if w["cell"] is not None:
        cell = torch.as_tensor(w["cell"], device=dev).reshape(1, 3, 3)
        x = ops.wrap_positions(x, cell[0])
        kw = dict(cell=cell, pbc=torch.ones(1, 3, dtype=torch.bool, device=dev))
    if w["mol_idx"] is not None:
        kw["batch_idx"] = torch.as_tensor(w["mol_idx"], device=dev)
    cap = 256 if cutoff <= 6 else 2400
    _, cnt, *_ = ops.neighbor_list(x, cutoff, max_neighbors=cap, **kw)
    return float(cnt.float().mean().item())


class Runner:
    """One workload on one device: engine, calculator, host / device copies of every step's coordinates."""

    def __init__(self, name, seed, dev, n_sets, gemm_backend=None, jitter_seed=0):
        import torch

        from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict

        self.torch = torch
        self.dev = dev
        self.w = w = make_workload(name, seed)
        self.spec = ModelSpec(num_charge_channels=w.get("channels", 1))
        self.sd = random_state_dict(0, self.spec)
        self.calc = AIMNet2Calculator((self.sd, self.spec), device=str(dev))
        self.eng = self.calc.engine
        if gemm_backend is not None:
            self.eng.set_gemm_backend(gemm_backend)
        self.pbc = w["cell"] is not None
        self.calc.set_lrcoulomb_method(w.get("coulomb", "dsf" if self.pbc else "simple"))
        self.N, self.B = len(w["numbers"]), len(w["charge"])
        rng = np.random.default_rng(seed + 17 * jitter_seed)
        self.coords_np = [(w["coord"] + rng.normal(0, 0.01, w["coord"].shape)).astype(np.float32) for _ in range(n_sets)]
        self.coords_h = [torch.from_numpy(c).pin_memory() for c in self.coords_np]
        self.coords_d = [c.to(dev) for c in self.coords_h]
        pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory()   # noqa: E731
        self.numbers_h, self.charge_h = pin(w["numbers"]), pin(w["charge"])
        self.mol_h, self.mult_h = pin(w["mol_idx"]), pin(w.get("mult"))
        to = lambda t: None if t is None else t.to(dev)   # noqa: E731
        self.numbers_d, self.charge_d, self.mol_d, self.mult_d = to(self.numbers_h), to(self.charge_h), to(self.mol_h), to(self.mult_h)
        self.cell_d = torch.from_numpy(w["cell"]).to(dev) if self.pbc else None
        self.out_h = {"energy": torch.empty(self.B, dtype=torch.float64).pin_memory().numpy(),
                      "charges": torch.empty(self.N, dtype=torch.float32).pin_memory().numpy(),
                      "forces": torch.empty(self.N, 3, dtype=torch.float32).pin_memory().numpy()}
        if w["stress"]:
            self.out_h["stress"] = torch.empty(3, 3, dtype=torch.float32).pin_memory().numpy()
        if self.spec.C == 2:
            self.out_h["spin_charges"] = torch.empty(self.N, dtype=torch.float32).pin_memory().numpy()

    def step_device(self, i):
        return self.eng.eval(self.coords_d[i], self.numbers_d, self.charge_d, mol_idx=self.mol_d, mult=self.mult_d,
                             cell=self.cell_d, host_cell=self.w["cell"], forces=True, stress=self.w["stress"])

    def step_host(self, i):
        self.eng.eval_host(self.coords_h[i], self.numbers_h, self.charge_h, mol_idx=self.mol_h, mult=self.mult_h,
                           cell=self.w["cell"], forces=True, stress=self.w["stress"], out=self.out_h)

    def api_inputs(self, i):
        d = {"coord": self.coords_np[i], "numbers": self.w["numbers"], "charge": self.w["charge"]}
        if self.w["mol_idx"] is not None:
            d["mol_idx"] = self.w["mol_idx"]
        if self.w.get("mult") is not None:
            d["mult"] = self.w["mult"]
        if self.pbc:
            d["cell"] = self.w["cell"]
        return d

    def step_api(self, i):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = self.calc(self.api_inputs(i), forces=True, stress=self.w["stress"])
        # what an MD / screening driver does with the returned device tensors: one asynchronous copy per output into
        # page-locked host buffers, one synchronisation
        if not hasattr(self, "_api_pinned"):
            self._api_pinned = {k: self.torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
        for k, v in out.items():
            self._api_pinned[k].copy_(v, non_blocking=True)
        self.torch.cuda.current_stream(self.dev).synchronize()
        return self._api_pinned

    def h2d_bytes(self):
        n = self.coords_h[0].numel() * 4 + self.numbers_h.numel() * 4 + self.charge_h.numel() * 4
        n += self.mol_h.numel() * 4 if self.mol_h is not None else 0
        n += self.mult_h.numel() * 4 if self.mult_h is not None else 0
        return int(n + (36 if self.pbc else 0))

    def d2h_bytes(self):
        return int(sum(int(v.nbytes) for v in self.out_h.values()))

    def time_events(self, fn, W, K):
        torch = self.torch
        for i in range(W):
            fn(i)
        torch.cuda.synchronize(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(W + i)
        e1.record()
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1)

    def time_wall(self, fn, W, K):
        for i in range(W):
            fn(i)
        self.torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        for i in range(K):
            fn(W + i)
        self.torch.cuda.synchronize(self.dev)
        return (time.perf_counter() - t0) * 1e3


def rooflines(r: Runner, step_ms: float, timed_region_s: float, traffic: dict):
    """Instrumented pass (CUDA events around every launch of the two dominant kernel classes; not the timed region)."""
    torch, eng = r.torch, r.eng
    eng.enable_timing(2)
    rows = []
    r.step_device(0)
    torch.cuda.synchronize(r.dev)
    for i in range(5):
        r.step_device(i % len(r.coords_d))
        torch.cuda.synchronize(r.dev)
        rows.append(eng.last_timing())
    eng.enable_timing(0)
    med = {k: float(np.median([x[k] for x in rows])) for k in rows[0]}
    peaks = load_peaks()
    burst = timed_region_s < 1.0   # burst peak for a short timed region, the sustained one inside a long step loop
    bf16 = peaks.get("bf16_tflops" if burst else "bf16_tflops_sustained", 1662.9 if burst else 1398.5)
    hbm = peaks.get("hbm_gbs", 6550.7)
    tc16 = eng.gemm_backend in (2, 3, 4, 5)
    peak_tensor = bf16 / 3.0 if tc16 else bf16 / 2.0 / 3.0   # three error-compensating MMAs per product (tf32: half rate)
    peak_src = (("MEASURED_PEAKS.json" if peaks else "fallback") +
                (" bf16_tflops (burst: timed region %.2f s)" % timed_region_s if burst else " bf16_tflops_sustained") +
                (" / 3 (3xFP16 split on the kind::f16 pipe)" if tc16 else " / 2 (tf32) / 3 (3xTF32 split)"))
    C, N = r.spec.C, r.N
    m_sr = mean_neighbors(r.w, r.dev, 5.0)
    m_lr = mean_neighbors(r.w, r.dev, 15.0) if r.pbc else 0.0   # non-periodic: pair walkers iterate molecule segments, no list
    f_mlp = 2.0 * 2.0 * MLP_MACS_PER_ATOM[C] * N
    f_conv = CONV_FLOP_PER_PAIR[C] * m_sr * N
    gemm_ms, conv_ms = med["gemm_ms"], med["conv_ms"]
    sms, sm_ghz = 148, peaks.get("sm_max_mhz", 1965.0) / 1e3
    peak_fma = sms * 128 * 2 * sm_ghz / 1e3   # TFLOP/s fp32: SMs x 128 lanes x 2 FLOP x clock (computed, not measured)
    g = {"kernel": "gemm_nt: per-atom MLP stacks, fwd + input-gradient bwd (%d launches/step incl. presplit)" % int(med["gemm_launches"]),
         "bound": "tensor", "achieved": f_mlp / (gemm_ms * 1e-3) / 1e12, "peak": peak_tensor, "unit": "TFLOP/s",
         "peak_source": peak_src, "ms_per_step": gemm_ms, "share_of_step": gemm_ms / step_ms,
         "algorithmic": "%.3f MFLOP/atom-step x %d atoms" % (f_mlp / N / 1e6, N), "traffic": traffic.get("gemm")}
    c = {"kernel": "conv_sv: AEV recomputed in flight + gather / contraction, fwd + analytic bwd (%d calls/step)" % int(med["conv_calls"]),
         "bound": "simt-fp32-fma", "achieved": f_conv / (conv_ms * 1e-3) / 1e12, "peak": peak_fma, "unit": "TFLOP/s",
         "peak_source": "computed: %d SMs x 128 lanes x 2 x %.3f GHz (MEASURED_PEAKS.json has no fp32 figure)" % (sms, sm_ghz),
         "ms_per_step": conv_ms, "share_of_step": conv_ms / step_ms,
         "algorithmic": "%d FLOP/pair x mean %.1f short-range neighbors x %d atoms" % (CONV_FLOP_PER_PAIR[C], m_sr, N),
         "traffic": traffic.get("conv")}
    for d in (g, c):
        d["frac"] = d["achieved"] / d["peak"]
    # t_roof = max(N F_total / P_tensor_eff, N Bytes / HBM)
    f_total = f_mlp + f_conv + AGH_FLOP_PER_ATOM * N
    bytes_total = N * (32 + 48 * m_sr + 32 * m_lr + 69_000)
    t_tensor, t_hbm = f_total / (peak_tensor * 1e12) * 1e3, bytes_total / (hbm * 1e9) * 1e3
    step = {"t_roof_ms": max(t_tensor, t_hbm), "t_tensor_ms": t_tensor, "t_hbm_ms": t_hbm, "t_measured_ms": step_ms,
            "frac": max(t_tensor, t_hbm) / step_ms, "flop_per_atom_step": f_total / N, "bytes_per_atom_step": bytes_total / N,
            "mean_sr_neighbors": m_sr, "mean_lr_neighbors": m_lr,
            "definition": "SURVEY.md 8(d): max(N F_total / P_tensor_eff, N Bytes / measured HBM GB/s) / measured step"}
    phases = {k: med[k] for k in ("neighbors_ms", "forward_ms", "pair_terms_ms", "backward_ms", "total_ms")}
    dominant = dict(g if gemm_ms >= conv_ms else c)
    dominant["dominant_class"] = "gemm" if gemm_ms >= conv_ms else "conv"
    dominant["gemm_ms_per_step"], dominant["conv_ms_per_step"], dominant["phase_ms"] = gemm_ms, conv_ms, phases
    return dominant, [g, c], step


def load_traffic(eng_backend: int, workload: str):
    """DRAM bytes per launch of the largest launch of each class from the committed `ncu --set full` capture of THIS build
    (profiles/ncu_traffic.json, keyed by the sha256 of the kernel sources); null when the library has changed since."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        from aimnetcentral_b200 import build as b

        if t.get("build_sha256") != b._digest() or t.get("workload") != workload or t.get("gemm_backend") != eng_backend:
            return {}
        return {"gemm": t.get("gemm"), "conv": t.get("conv")}
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--gemm-backend", type=int, default=None)
    ap.add_argument("--conv-impl", type=int, default=None, help="0 list kernels always, 1 dense shared-memory forward for small molecules (default), 2 dense forward + backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from aimnetcentral_b200.sharded import ShardedCalculator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), max(1, args.steps)
    n_sets = W + K
    r = Runner(args.workload, args.seed + rank, dev, n_sets, args.gemm_backend, jitter_seed=rank)
    if args.conv_impl is not None:
        r.eng.set_conv_impl(args.conv_impl)
    N, B, w = r.N, r.B, r.w
    sampler = ClockSampler(local).start() if rank == 0 else None

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    multi = None
    if world == 1:
        barrier()
        ms = r.time_events(r.step_device, W, K)
        atoms_total = N
    else:
        # The product's multi-GPU module.  The GLOBAL batch is resident on every rank (weak: world x 1024 molecules, built
        # from the same seeds on every rank; strong: the N = 1 batch); ShardedCalculator evaluates this rank's contiguous
        # atom-balanced shard and gathers every output with one fused NCCL all_gather per step, inside the timed region.
        if w["mol_idx"] is None:
            reps, mode = world, "replicas"   # independent periodic replicas, one box per GPU
            base = Runner(args.workload, args.seed, dev, n_sets)
            glob = [dict(coord=base.coords_d[i].unsqueeze(0).repeat(reps, 1, 1), numbers=base.numbers_d.unsqueeze(0).repeat(reps, 1),
                         charge=base.charge_d.repeat(reps), cell=base.cell_d.unsqueeze(0).repeat(reps, 1, 1)) for i in range(n_sets)]
            del base
        else:
            reps, mode = (world if args.scaling == "weak" else 1), args.scaling
            parts = [Runner(args.workload, args.seed + k, dev, n_sets, jitter_seed=k) for k in range(reps)]
            mol = torch.cat([x.mol_d + k * B for k, x in enumerate(parts)])
            numbers, charge = torch.cat([x.numbers_d for x in parts]), torch.cat([x.charge_d for x in parts])
            mult = torch.cat([x.mult_d for x in parts]) if r.mult_d is not None else None
            glob = []
            for i in range(n_sets):
                d = dict(coord=torch.cat([x.coords_d[i] for x in parts]), numbers=numbers, charge=charge, mol_idx=mol)
                if mult is not None:
                    d["mult"] = mult
                glob.append(d)
            del parts
        atoms_total = reps * N
        sharded = ShardedCalculator(r.calc)

        def step_sharded(i):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return sharded(glob[i], forces=True, stress=w["stress"])

        for i in range(W):
            step_sharded(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            out = step_sharded(W + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        multi = {"module": "aimnetcentral_b200.sharded.ShardedCalculator(AIMNet2Calculator)", "mode": mode,
                 "global_atoms": atoms_total, "global_forces_shape": list(out["forces"].shape),
                 "gather": "one fused all_gather_into_tensor of every output per step (NCCL), inside the timed region"}
    launches = r.eng.last_launches() * K
    ms_max = maxr(ms)
    value = atoms_total * K / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + compute + D2H inside every step); every rank its own batch ----
    barrier()
    e2e_ms = maxr(r.time_wall(r.step_host, 2, K))
    e2e_value = world * N * K / (e2e_ms * 1e-3)
    # ---- api: the front door (AIMNet2Calculator.__call__, numpy in, .cpu() out) ----
    api_ms = maxr(r.time_wall(r.step_api, 2, K))
    api_value = world * N * K / (api_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    step_ms = ms_max / K if world == 1 else r.time_events(lambda i: r.step_device(i % n_sets), 1, 3) / 3
    dominant, classes, step_roof = rooflines(r, step_ms, ms_max * 1e-3, load_traffic(r.eng.gemm_backend, args.workload))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, N, B),
            "impl_detail": {"gemm_backend": {2: "tcgen05-3xfp16-rowchunk-scaled", 3: "tcgen05-3xfp16-rowchunk-scaled-pipelined-epilogue",
                                             1: "tcgen05-3xtf32"}.get(r.eng.gemm_backend, "simt-fp32"),
                            "conv": r.eng.conv_mode(),
                            "value_path": "aimnet2_engine_eval (C ABI, device pointers)" if world == 1 else "ShardedCalculator",
                            "multi_gpu": multi or "single"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r.h2d_bytes(), "d2h_bytes_per_step": r.d2h_bytes(),
                    "path": "aimnet2_engine_eval_host (C ABI, pinned host buffers)"},
            "api": {"value": api_value, "unit": UNIT, "ratio_to_e2e": api_value / e2e_value,
                    "path": "AIMNet2Calculator.__call__(numpy inputs) + copy of every output into pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": dominant, "roofline_classes": classes, "step_roofline": step_roof,
        }
        if world == 1 and not args.no_extra and args.workload == "cfg2":
            extra = {}
            del r
            for name, k in (("cfg1", 50), ("cfg3", 10), ("cfg4", 10), ("cfg5", 5)):
                torch.cuda.empty_cache()
                s = ClockSampler(local).start()
                rx = Runner(name, args.seed, dev, 3 + k)
                ms_x = rx.time_events(rx.step_device, 3, k)
                ms_h = rx.time_wall(rx.step_host, 2, k)
                extra[name] = {"workload": rx.w["desc"], "atoms": rx.N, "steps": k, "warmup": 3, "ms_per_step": ms_x / k,
                               "value": rx.N * k / (ms_x * 1e-3), "e2e": rx.N * k / (ms_h * 1e-3), "unit": UNIT,
                               "gpu_launches_per_step": rx.eng.last_launches()}
                if name == "cfg1":   # launch-bound: the same step replayed as one CUDA graph (AIMNet2Calculator(cuda_graph=True))
                    rx.eng.enable_cuda_graph(True)
                    ms_g = rx.time_events(rx.step_device, 3, k)
                    ms_gh = rx.time_wall(rx.step_host, 3, k)
                    extra[name]["cuda_graph"] = {"ms_per_step": ms_g / k, "value": rx.N * k / (ms_g * 1e-3), "e2e": rx.N * k / (ms_gh * 1e-3),
                                                 "stats": rx.eng.graph_stats()}
                extra[name]["clocks"] = s.stop()
                del rx
            line["extra_workloads"] = extra
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = time_cpu(args.workload, args.seed, warmup=1, steps=None, budget_s=15.0)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
