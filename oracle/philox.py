"""TEST INFRASTRUCTURE — not a product path.

numpy restatement of the counter-based dropout stream the CUDA kernels use
(`carca_replication_b200/csrc/philox.cuh`).  The reference draws its dropout
masks from torch's CPU `bernoulli_` (src/carca.py:218,286,289,406), a stream no
GPU kernel can reproduce, so train-mode parity is defined on OUR stream: the
oracle applies exactly the keep/drop decisions the kernels make, and everything
downstream of the mask follows the reference arithmetic.

Stream definition (shared by kernels and oracle):
  Philox4x32-10, key = (seed_lo, seed_hi),
  counter = (elem // 4 low 32 bits, elem // 4 high 32 bits, site, 0);
  element `elem` takes output word `elem % 4`;
  u = (word >> 8) * 2**-24  (exact in fp32);  keep  <=>  u >= p.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs uint32 arrays (or scalars)."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.broadcast_to(np.asarray(c1, dtype=np.uint32), c0.shape).copy()
    c2 = np.broadcast_to(np.asarray(c2, dtype=np.uint32), c0.shape).copy()
    c3 = np.broadcast_to(np.asarray(c3, dtype=np.uint32), c0.shape).copy()
    c0 = c0.copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def keep_mask(n_elems: int, p: float, seed: int, site: int) -> np.ndarray:
    """Boolean keep mask (True = kept) for flat element indices 0..n_elems-1."""
    if n_elems == 0:
        return np.zeros(0, dtype=bool)
    n_ctr = (n_elems + 3) // 4
    ctr = np.arange(n_ctr, dtype=np.uint64)
    c0 = (ctr & _MASK32).astype(np.uint32)
    c1 = (ctr >> np.uint64(32)).astype(np.uint32)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    o = philox4x32_10(c0, c1, np.uint32(site), np.uint32(0), k0, k1)
    words = np.stack(o, axis=1).reshape(-1)[:n_elems]
    u = (words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return u >= np.float32(p)
