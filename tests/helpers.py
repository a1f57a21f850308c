"""Shared helpers for the parity tests (oracle = checker only; see oracle/nca_oracle.py header)."""
import glob
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

DYNCA_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                     if not os.path.basename(p).startswith(("weights_", "enc_", "encoder", "callers")))
ENC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "enc_*.npz")))


def load_case(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(d["meta"]))
    t = {k: torch.from_numpy(d[k]) for k in d.files if k != "meta"}
    return t, meta


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def cond_for(t, meta):
    """cond tensor the oracle expects ([B,cc,H,W] or None) for a golden DyNCA case."""
    from oracle import nca_oracle as O
    if meta["cond"] == "cpe":
        return O.cpe2d(meta["B"], meta["H"], meta["W"])
    if meta["cond"] == "edges":
        return O.edge_extract(t["cond_img"], meta["edge_transform"])
    return None
