import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from pde_b200 import ops, _xfer
from pde_b200 import ks2d as K
U = ops.synth_field(2000, 100, 100, seed=2, noise=0.05).cpu().numpy()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("plain .cuda() ms", round(t(lambda: torch.from_numpy(U).cuda()), 2), " staged ms", round(t(lambda: _xfer.to_device(U)), 2))
d = _xfer.to_device(U)
print("plain .cpu() ms", round(t(lambda: d.cpu().numpy()), 2), " staged ms", round(t(lambda: _xfer.to_host(d)), 2))
for kw in (dict(method="pointwise", dictionary="true"), dict(method="blockwise", dictionary="true"), dict(method="blockwise", dictionary="rich", grid_search=True)):
    print(kw, "fit_from_field ms", round(t(lambda: K.fit_from_field(U, 0.5, 0.5, 1e-3, **kw), 3), 2))
print("build_dictionary_true (literal, NumPy in/out) ms", round(t(lambda: K.build_dictionary_true(U, 0.5, 0.5), 3), 2))
