"""Fault isolation for the rare wrong result of round 1: the very first evaluation of a fresh engine in a
fresh process, compared BITWISE with a second evaluation of the same input (every kernel is run-to-run reproducible, so
any difference is an event), buffer by buffer over the whole device workspace (aimnet2_engine_debug_layout /
_read_workspace), and against the CPU oracle at the north-star tolerance.

    python tools/first_touch.py TAG [--no-oracle] [--save FILE.npz | --expect FILE.npz] [--quiet]

  --save    store this process's outputs (bitwise reference for --expect runs)
  --expect  compare this process's FIRST evaluation bitwise with the stored outputs (fresh-process loop, tools/fresh_loop.sh)
Appends one JSON line per run to gpurun_out/first_touch_log.jsonl; on an event the differing buffers go to
gpurun_out/first_touch_TAG_event.npz and the report names the first diverging buffer and the atoms (rows) involved.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# order in which the buffers are produced during one evaluation (first diverging one = the faulty kernel's output)
ORDER = ["nb_sr", "cnt_sr", "a0", "x", "x16_hi", "x16_lo", "x16_inv", "gp00", "gp01", "y0", "sumq0", "sumf0", "a1", "q0",
         "T_a0", "gp10", "gp11", "gp12", "y1", "sumq1", "sumf1", "a2", "q1", "T_a1", "T_q1", "gp20", "gp21", "gp22", "gp23", "aim",
         "aim_inv", "T_a2", "T_q2", "gp_h1", "h1", "h1_inv", "gp_h2", "h2", "e_nn", "e_sr", "e_lr", "e_d3", "gq", "cn", "dEdCN", "dz32",
         "dzA", "dzB", "dzA_inv", "dzB_inv", "dx", "dS_a", "dS_q", "grad_a", "grad_q", "da_tot", "dq", "dq_base", "s1"]


def gpu_state():
    try:
        q = "name,clocks.sm,clocks.mem,pstate,temperature.gpu,power.draw,ecc.errors.corrected.volatile.total,ecc.errors.uncorrected.volatile.total"
        return subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True,
                              timeout=20).stdout.strip()
    except Exception as ex:  # noqa: BLE001
        return f"nvidia-smi failed: {ex}"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next(line.split(":", 1)[1].strip() for line in f if line.startswith("model name"))
    except Exception:  # noqa: BLE001
        return "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--save")
    ap.add_argument("--expect")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--mols", type=int, default=64)
    args = ap.parse_args()
    t_start = time.time()
    state0 = gpu_state()
    import torch

    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    coord, numbers = random_molecules(args.mols, 50, seed=99)
    inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(args.mols, np.float32)}
    N = args.mols * 50
    calc = AIMNet2Calculator((sd, spec), device="cuda:0")
    rec = {"tag": args.tag, "pid": os.getpid(), "gpu_before": state0, "host": cpu_model(), "uptime_s": float(open("/proc/uptime").read().split()[0]),
           "atoms": N}

    def run(c):
        out = {k: v.cpu().numpy().copy() for k, v in c(dict(inp), forces=True).items()}
        return out

    full = not args.expect
    out1 = run(calc)
    snap1 = calc.engine.debug_snapshot() if full else None
    events = []
    if args.expect:
        exp = np.load(args.expect)
        bad = [k for k in exp.files if not np.array_equal(exp[k], out1[k])]
        if bad:
            events.append(f"first evaluation differs from the stored reference outputs in {bad}")
            snap1 = calc.engine.debug_snapshot()
    if full or events:
        out2 = run(calc)
        snap2 = calc.engine.debug_snapshot()
        bad_out = [k for k in out1 if not np.array_equal(out1[k], out2[k])]
        bad_buf = [k for k in snap1 if k in snap2 and not np.array_equal(snap1[k], snap2[k])]
        if bad_out or bad_buf:
            events.append(f"evaluation 1 and 2 of one engine differ: outputs {bad_out}, buffers {bad_buf}")
            first = [k for k in ORDER if k in bad_buf]
            rep = {}
            for k in bad_buf:
                a, b = snap1[k], snap2[k]
                per = len(a) // N if len(a) % N == 0 and len(a) >= N else 0
                if per:
                    rows = np.nonzero((a.reshape(N, per) != b.reshape(N, per)).any(axis=1))[0]
                    cols = np.nonzero((a.reshape(N, per) != b.reshape(N, per)).any(axis=0))[0]
                    rep[k] = {"rows": len(rows), "row_min": int(rows.min()), "row_max": int(rows.max()), "rows_first": rows[:24].tolist(),
                              "byte_col_min": int(cols.min()), "byte_col_max": int(cols.max()), "bytes_per_row": per}
                else:
                    idx = np.nonzero(a != b)[0]
                    rep[k] = {"bytes": len(idx), "first": idx[:16].tolist()}
            rec["first_diverging_in_order"] = first[:6]
            rec["buffers"] = rep
            keep = {f"{k}_1": snap1[k] for k in bad_buf[:12]}
            keep.update({f"{k}_2": snap2[k] for k in bad_buf[:12]})
            keep.update({f"out1_{k}": v for k, v in out1.items()})
            keep.update({f"out2_{k}": v for k, v in out2.items()})
            np.savez_compressed(f"gpurun_out/first_touch_{args.tag}_event.npz", **keep)
        # a second fresh engine in the same process
        calc_b = AIMNet2Calculator((sd, spec), device="cuda:0")
        out3 = run(calc_b)
        bad3 = [k for k in out1 if not np.array_equal(out1[k], out3[k])]
        if bad3:
            events.append(f"second fresh engine differs from the first engine's first evaluation in {bad3}")
        # which of the other GEMM backends agree (tolerance, they are different arithmetic)
        for be in (0, 1):
            calc_b.engine.set_gemm_backend(be)
            ob = run(calc_b)
            rec[f"backend{be}_vs_first_dF"] = float(np.abs(ob["forces"] - out1["forces"]).max())
        calc_b.engine.set_gemm_backend(2)
    if not args.no_oracle and not args.expect:
        from oracle.calculator_oracle import oracle_calculate

        ref = oracle_calculate(sd, inp)
        rec["vs_oracle"] = {"dE": float(np.abs(out1["energy"] - ref["energy"]).max()), "dF": float(np.abs(out1["forces"] - ref["forces"]).max()),
                            "dq": float(np.abs(out1["charges"] - ref["charges"]).max())}
        if not (rec["vs_oracle"]["dE"] < 1e-4 and rec["vs_oracle"]["dF"] < 1e-4 and rec["vs_oracle"]["dq"] < 1e-4):
            events.append(f"first evaluation is outside the tolerance vs the oracle: {rec['vs_oracle']}")
    if args.save:
        np.savez(args.save, **out1)
    rec["events"] = events
    rec["gpu_after"] = gpu_state()
    rec["seconds"] = round(time.time() - t_start, 2)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/first_touch_log.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
    if events or not args.quiet:
        print(f"[first_touch {args.tag}] {'EVENT: ' + ' | '.join(events) if events else 'clean'}  "
              f"{json.dumps({k: rec[k] for k in rec if k not in ('events', 'tag')})[:1500]}")
    sys.exit(3 if events else 0)


if __name__ == "__main__":
    main()
