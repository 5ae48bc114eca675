import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, lmpc_b200
from oracle import mpqp
z = np.load("tests/golden/invpend_explicit.npz")
g = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
N = 10_000_000
Th = np.asfortranarray(mpqp.sample_theta(z["lo"], z["hi"], N, seed=1))
L = lmpc_b200.lib()
U = np.zeros((1, N), order="F"); reg = np.zeros(N, np.int32)
def call(th, u, r): assert L.lmpc_gpu_evaluate_batch(g._t, th.ctypes.data, N, u.ctypes.data, r.ctypes.data) == 0
best = 1e9
for i in range(4):
    t0 = time.perf_counter(); call(Th, U, reg); dt = time.perf_counter() - t0
    if i: best = min(best, dt)
Tp = lmpc_b200.pinned_empty(Th.shape, order="F"); Tp[...] = Th
Up = lmpc_b200.pinned_empty((1, N), order="F"); rp = lmpc_b200.pinned_empty(N, np.int32)
bp = 1e9
for i in range(4):
    t0 = time.perf_counter(); call(Tp, Up, rp); dt = time.perf_counter() - t0
    if i: bp = min(bp, dt)
print("explicit evaluate e2e: pageable reused %.3e theta/s, pinned %.3e theta/s, equal %s" % (N / best, N / bp, np.array_equal(U, Up) and np.array_equal(reg, rp)))
