"""torchrun --nproc-per-node N scripts/dist_gram_check.py: the row-sharded Gram (folded blocks, sharded precompute, one
all-gather) must equal the single-GPU Gram bit for bit."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from matchcake_b200 import ops as mops
from matchcake_b200.ml.kernels import FermionicPQCKernel
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
for n, nf, ns in [(32, 64, 1000), (8, 8, 37), (6, 12, 5)]:
    x = np.random.default_rng(0).random((ns, nf))
    k = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x)
    sharded = k(torch.as_tensor(x).cuda())
    # single-GPU reference: the same covariances through one un-sharded call
    cov = mops.cov_basis(k._x_to_sptm(torch.as_tensor(x).cuda()), n) if k._block_plan() is None or not k.use_block_factorisation else None
    if cov is not None:
        single = mops.gram(cov, cov, 2.0 ** -n)
    else:
        b = k._x_to_blocks(torch.as_tensor(x).cuda()); single = mops.gram_blocks4(b, b)
    err = float((sharded - single).abs().max())
    rect = k(torch.as_tensor(x).cuda(), torch.as_tensor(x[: max(2, ns // 3)]).cuda())
    if rank == 0:
        print(f"N={n} ns={ns}: max |sharded - single| = {err:.1e}, rect {tuple(rect.shape)}")
    assert err == 0.0
dist.barrier(); dist.destroy_process_group()
if rank == 0:
    print("dist gram check ok")
