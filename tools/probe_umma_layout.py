"""Discover how the tensor core reads an MN-major no-swizzle tf32 B operand: for each candidate
(lbo, sbo) find, for every smem float index i of the B image, which (n, k) it feeds."""
import sys; sys.path.insert(0, '/root/repo')
import ctypes as C, numpy as np, torch
from carca_replication_b200 import _native as N
lib = N.lib()
lib.carca_umma_probe.restype = C.c_int
dev = torch.device('cuda')
Nn, K = 16, 8
# A: K-major image [K/4][128][4], A[m][k] = 2^k  (exact in tf32)
a_img = np.zeros((K // 4, 128, 4), np.float32)
for k in range(K): a_img[k // 4, :, k % 4] = 2.0 ** k
a_t = torch.from_numpy(a_img.reshape(-1)).to(dev)
c = torch.zeros((128, Nn), dtype=torch.float32, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
size = Nn * K
def idesc(n, b_mn): return (1 << 4) | (2 << 7) | (2 << 10) | ((n >> 3) << 17) | ((128 >> 4) << 24) | ((1 << 16) if b_mn else 0)
def run(b_img, lbo, sbo, b_mn):
    b_t = torch.from_numpy(b_img.astype(np.float32)).to(dev)
    rc = lib.carca_umma_probe(C.c_void_p(c.data_ptr()), C.c_void_p(a_t.data_ptr()), a_t.numel(), C.c_void_p(b_t.data_ptr()), b_t.numel(),
                              Nn, 1, 128 * 16, 128, 0, lbo, sbo, 0, idesc(Nn, b_mn), C.c_void_p(status.data_ptr()), None)
    assert rc == 0, lib.carca_last_error()
    torch.cuda.synchronize()
    assert status.item() == 0
    return c[0].cpu().numpy().copy()      # row 0: sum_k 2^k * B'(n,k)
for b_mn in (1, 0):
    for (lbo, sbo) in [(Nn * 32, 128), (128, Nn * 32), (128, 128), (Nn*16, 128), (64, 128), (128, 64), (256, 128), (128, 256), (32,128),(128,32), (16, 128), (128, 16)]:
        # bit-plane decode: which smem index feeds (n,k)
        idx_of = np.zeros((Nn, K), np.int64)
        valid = np.ones((Nn, K), bool)
        ones = run(np.ones(size * 4), lbo, sbo, b_mn)   # image 4x larger to catch big strides
        for bit in range(9):
            img = ((np.arange(size * 4) >> bit) & 1).astype(np.float32)
            out = run(img, lbo, sbo, b_mn).astype(np.int64)
            for n in range(Nn):
                for k in range(K):
                    if (out[n] >> k) & 1: idx_of[n, k] |= (1 << bit)
        print(f"b_mn={b_mn} lbo={lbo} sbo={sbo} all-ones row: {ones[:4]}")
        print("  idx_of[n, k] for n=0..15 (rows), k=0..7:")
        for n in range(Nn): print("   n=%2d" % n, idx_of[n].tolist())
