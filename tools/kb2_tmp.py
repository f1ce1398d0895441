import sys; sys.path.insert(0,'/root/repo')
from dlv3p_b200 import ffi
for th in (8,4):
    for dims in ((32,32,32,728,1,1),(32,256,256,128,1,1),(32,64,64,728,1,1)):
        ms=ffi.op_bb_time(1, list(dims)+[th], 20, 0)
        print('dw',dims,'TH',th,'%.4f ms'%ms)
for K,N,res in ((728,728,0),(728,728,1),(1536,2048,0),(1024,1536,0)):
    ms=ffi.op_bb_time(0,[32768,K,N,res,0],20,0); print('gemm',K,N,res,'%.4f ms %.1f TF'%(ms,2*32768*K*N/ms/1e9))
