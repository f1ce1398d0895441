import sys
sys.path.insert(0, '/root/repo')
from dlv3p_b200 import ffi
M = 32768
for K, N in ((728, 728), (1536, 2048), (256, 728), (1024, 1536)):
    for res in (0, 1):
        for flags in (0, 1):
            ms = ffi.op_bb_time(0, [M, K, N, res, 0], 20, flags)
            print('gemm M=%d K=%d N=%d res=%d flags=%d: %.4f ms  %.1f TFLOP/s' % (M, K, N, res, flags, ms, 2.0 * M * K * N / ms / 1e9), flush=True)
