"""Host-side profile of the small-state path: BASELINE configs[0] (LUCJ n_reps=2, norb=12, nelec=(6,6)), NumPy in/out."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ffsim_b200 as ffsim

norb, nelec = 12, (6, 6)
op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=2, with_final_orbital_rotation=True, seed=12)
vec = ffsim.hartree_fock_state(norb, nelec)
for _ in range(5):
    out = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    out = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
torch.cuda.synchronize()
print("numpy in/out: %.3f ms per application" % ((time.perf_counter() - t0) / 50 * 1e3))
dev = torch.from_numpy(vec).cuda()
for _ in range(5):
    ffsim.apply_unitary(dev, op, norb=norb, nelec=nelec, copy=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    ffsim.apply_unitary(dev, op, norb=norb, nelec=nelec, copy=False)
torch.cuda.synchronize()
print("device tensor, copy=False: %.3f ms per application" % ((time.perf_counter() - t0) / 50 * 1e3))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    ffsim.apply_unitary(dev, op, norb=norb, nelec=nelec, copy=False)
b.record(); torch.cuda.synchronize()
print("  of which device time (events): %.3f ms" % (a.elapsed_time(b) / 50))
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    ffsim.apply_unitary(dev, op, norb=norb, nelec=nelec, copy=False)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
