"""Shared helpers for the parity tests (fixtures, tolerances, comparison rules)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_CASES = ["beauty_ca", "beauty_dot", "men_ca", "men_dot_d128", "noresid_dot", "noresid_ca",
               "learnable_ca", "sinus_dot", "single_user_ca",
               # non-default module variants
               "idemb_dot", "mlpid_ca", "attr_dot", "attrctx_ca", "attrctx_dense_ca", "wdot_all", "wdotnorm_all"]

# north_star: fp32 scores within 1e-4 relative
FP32_RTOL = 1e-4


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return cfg, sd, z


def batch_of(z, split):
    keys = ["p_x", "p_a", "p_c", "o_x", "o_a", "o_c", "y_true"]
    return tuple(torch.from_numpy(z[f"{split}/in/{k}"]) for k in keys)


def oracle_cfg(cfg, **over):
    from oracle.carca_oracle import OracleConfig

    kw = dict(d=cfg["d"], n_heads=cfg["H"], n_blocks=cfg["n_blocks"], decoder=cfg["decoder"],
              residual_sa=cfg["residual_sa"], residual_ca=cfg["residual_ca"], p_drop=cfg["p"],
              learnable_pos=cfg["encoding"] == "learnable", sinus_pos=cfg["encoding"] == "positional",
              embedding=cfg.get("embedding", "all"), gamma=cfg.get("gamma", 0.9))
    kw.update(over)
    return OracleConfig(**kw)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))) if a.size else 0.0


def grad_err(a, b, floor=1e-5):
    """Max abs error normalised by the tensor's own scale (grads have many ~0 entries).  The 1e-5
    floor covers gradients that are mathematically zero (softmax is shift-invariant, so d/d WK.bias
    is pure rounding noise, ~1e-10 in the reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), floor)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def topk_equal_up_to_ties(scores_a, scores_b, k, tol=0.0):
    """Top-k index sets agree, allowing swaps among candidates whose reference scores tie
    (within tol) at the k-th place."""
    a = np.asarray(scores_a)
    b = np.asarray(scores_b)
    for ra, rb in zip(a, b):
        ia = np.argsort(-ra, kind="stable")[:k]
        ib = np.argsort(-rb, kind="stable")[:k]
        if set(ia) == set(ib):
            continue
        kth = np.sort(rb)[::-1][k - 1]
        for j in set(ia) ^ set(ib):
            if abs(rb[j] - kth) > tol:
                return False
    return True


def grad_floor(param_name):
    """d/d WK.bias is mathematically zero (softmax shift invariance): what either side holds is
    summation noise of terms ~1e-3, so it is compared on an absolute 1e-3 scale."""
    return 1e-3 if param_name.endswith("WK.bias") else 1e-5
