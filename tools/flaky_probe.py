"""Hunt for a rare wrong result seen in test_oracle_parity_random_batch: fresh calculator + one evaluation, repeated."""
import sys, gc
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
from aimnetcentral_b200.structures import random_molecules

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = sys.argv[2] if len(sys.argv) > 2 else "fresh"
spec = ModelSpec(); sd = random_state_dict(0, spec)
coord, numbers = random_molecules(64, 50, seed=99)
inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(64, np.float32)}
other = random_molecules(8, 50, seed=5)
ref = None
bad = 0
calc = None
for r in range(reps):
    if mode == "fresh" or calc is None:
        calc = AIMNet2Calculator((sd, spec), device="cuda:0")
    if mode == "alternate":
        calc({"coord": other[0], "numbers": other[1], "charge": np.zeros(8, np.float32)}, forces=True)
    out = {k: v.cpu().numpy().copy() for k, v in calc(dict(inp), forces=True).items()}
    if ref is None:
        ref = out
        continue
    if not all(np.array_equal(ref[k], out[k]) for k in ref):
        bad += 1
        de = np.abs(out["energy"] - ref["energy"]); dq = np.abs(out["charges"] - ref["charges"]); df = np.abs(out["forces"] - ref["forces"]).max(axis=-1)
        print(f"rep {r}: MISMATCH max dE {de.max():.3e} dq {dq.max():.3e} dF {df.max():.3e}")
        print("   molecules with dE>1e-6:", np.nonzero(de > 1e-6)[0][:20], " n=", int((de > 1e-6).sum()))
        qa = np.nonzero(dq.reshape(-1) > 1e-6)[0]
        print("   atoms with dq>1e-6: n=", len(qa), " first", qa[:10], " last", qa[-10:])
        fa = np.nonzero(df.reshape(-1) > 1e-5)[0]
        print("   atoms with dF>1e-5: n=", len(fa), " first", fa[:10], " last", fa[-10:])
print(f"{mode}: {reps} evaluations, {bad} mismatches")
