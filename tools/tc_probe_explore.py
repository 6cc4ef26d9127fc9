"""Decode which shared-memory float the tensor core reads for logical A[m][k] (SWIZZLE_NONE):
the A image holds its own float index (mod 2048), B is an identity selector."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gan_mpc_b200 import _lib

NB, K = 32, 8
B = np.zeros((NB, K), np.float32)
for k in range(K):
    B[k, k] = 1.0
nfl = 16384
img = (np.arange(nfl) % 2048).astype(np.float32)
bs = NB * 16 + 16
def run(major, lbo, sbo):
    return _lib.tc_probe(img.reshape(1, -1), B, major, lbo, sbo, 0, 0, 0, bs, 128, bs, 128,
                         nfl * 4, nfl * 4 + 8 * bs + 256)
for major, lbo, sbo in ((2, 2048, 128), (3, 4096, 128), (3, 128, 4096), (3, 2048, 1024)):
    D = run(major, lbo, sbo)
    print(f"major={major&1} lbo={lbo} sbo={sbo}: float index (mod 2048) read for A[m][k]; nonzero={np.count_nonzero(D)}")
    for m in (0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 32, 33, 64, 127):
        print("  m=%3d:" % m, " ".join("%5d" % int(D[m, k]) for k in range(K)))
