import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import ffsim_b200 as ffsim
from ffsim_b200.distributed import ShardedVector, ROWS, COLS
from oracle import cref, models, rand, gates
lr = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr); rank = dist.get_rank()
for norb, nelec in [(6, (1, 6)), (6, (6, 1)), (5, (2, 0)), (4, (2, 2))]:
    rng = np.random.default_rng(norb)
    full = rand.random_state_vector(models.dim(norb, nelec), seed=rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    def err(sv, want): 
        return float(np.linalg.norm(sv.gather().cpu().numpy() - want) / np.linalg.norm(want))
    sv = ShardedVector.from_global(full, norb, nelec, device=dev)
    e0 = err(sv, full)
    sv.set_layout(COLS); e1 = err(sv, full)   # round trip rows->cols->rows
    sv = ShardedVector.from_global(full, norb, nelec, device=dev)
    ffsim.apply_orbital_rotation(sv, (ua, None), norb, nelec, copy=False); ea = err(sv, cref.apply_orbital_rotation(full, (ua, None), norb, nelec))
    sv = ShardedVector.from_global(full, norb, nelec, device=dev)
    ffsim.apply_orbital_rotation(sv, (None, ub), norb, nelec, copy=False); eb = err(sv, cref.apply_orbital_rotation(full, (None, ub), norb, nelec))
    sv = ShardedVector.from_global(full, norb, nelec, device=dev)
    sv.set_layout(COLS)
    ffsim.apply_diag_coulomb_evolution(sv, mat, 0.3, norb, nelec, copy=False); ed = err(sv, gates.apply_diag_coulomb_evolution(full, mat, 0.3, norb, nelec))
    sv = ShardedVector.from_global(full, norb, nelec, device=dev)
    ffsim.apply_diag_coulomb_evolution(sv, mat, 0.3, norb, nelec, copy=False); er = err(sv, gates.apply_diag_coulomb_evolution(full, mat, 0.3, norb, nelec))
    if rank == 0:
        print(os.environ.get("FFSIM_B200_EXCHANGE"), norb, nelec, "gather", e0, "roundtrip", e1, "alpha", ea, "beta", eb, "diag(cols)", ed, "diag(rows)", er, flush=True)
dist.barrier(); dist.destroy_process_group()
