import sys; sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
from pde_b200 import _lib as L, ops
U = ops.synth_field(256, 2048, 2048, seed=0, noise=0.05)
for blk in [(3, 8, 8), (3, 16, 16), (3, 32, 32)]:
    for _ in range(3):
        ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=blk)
    torch.cuda.synchronize()
