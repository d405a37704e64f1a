import sys; sys.path.insert(0,'.')
import numpy as np
from pyvibdmc_b200 import kernels as K
for nc in (9,18):
    for mode in (0,1,2):
        best=1e9
        for _ in range(3):
            K.normals(4_000_000, nc, seed=1, step=2, rng_mode=mode)
            best=min(best,K.last_kernel_ms())
        print(nc, mode, round(best*1e3,1), 'us per 4e6 walkers;', round(4e6*nc*16/best/1e6,0),'GB/s equiv (16 B/normal)')
