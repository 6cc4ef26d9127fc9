"""tcgen05.mma rate for the planner's operand shapes (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_mpc_b200 import _lib
R = 200
for grid in (1, 148):
    print(f"--- grid={grid}  (cycles per MMA, M=128 K=8 tf32, SWIZZLE_NONE planner layout)")
    for N in (32, 64, 128):
        bl = N * 16 + 16
        row = []
        for nacc in (0, 2, 4):
            if nacc * N > 512: row.append(float('nan')); continue
            row.append(_lib.tc_mma_bench(N, 8, R, 3200, 128, 12800, bl, 128, 2 * bl, 0, nacc, grid))
        print(f"N={N:3d}: same-acc {row[0]:6.1f}   2 accs {row[1]:6.1f}   4 accs {row[2]:6.1f}")
    c = _lib.tc_mma_bench(64, 8, R, 3200, 128, 12800, 1040, 128, 2080, 0, 1, grid)
    print(f"planner pair (N=64 then dependent N=32 on the same A): {c:6.1f}")
