"""Derive aimnetcentral_b200/data/dftd3_tables.npz from the reference's aimnet/dftd3_data.pt (Grimme DFT-D3 reference
C6 / CN tables).  Run once in the build container:  python -m oracle.make_d3_tables

Unpacking follows aimnet/modules/lr.py:1405-1422: c6ab[...,0] -> C6 reference, c6ab[...,1] -> reference CN of the
first element's a-th reference system (constant over the partner element and its index, verified below), so it is
stored as a compact (95,5) table.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AIMNET_REFERENCE_ROOT", "/root/reference")


def main():
    p = torch.load(os.path.join(REF, "aimnet", "dftd3_data.pt"), map_location="cpu", weights_only=True)
    c6ab = p["c6ab"].float().numpy()
    c6ref = np.ascontiguousarray(c6ab[..., 0])
    cn_i = c6ab[..., 1]  # (95,95,5,5)
    valid = c6ref != 0
    cnref = np.full((95, 5), -1.0, np.float32)
    for z in range(95):
        for a in range(5):
            vals = cn_i[z, :, a, :][valid[z, :, a, :]]
            if vals.size:
                assert np.all(vals == vals[0]), (z, a)
                cnref[z, a] = vals[0]
    # every valid entry must be reproduced by the compact table (both roles i and j)
    assert np.array_equal(np.where(valid, cn_i, 0), np.where(valid, cnref[:, None, :, None], 0))
    cn_j = c6ab[..., 2]
    assert np.array_equal(np.where(valid, cn_j, 0), np.where(valid, cnref[None, :, None, :], 0))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aimnetcentral_b200", "data",
                       "dftd3_tables.npz")
    np.savez_compressed(out, c6ref=c6ref, cnref=cnref, rcov=p["rcov"].float().numpy(), r4r2=p["r4r2"].float().numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
