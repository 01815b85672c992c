"""torchrun --nproc-per-node N scripts/dist_gram_check.py: the row-sharded Gram builds (folded row blocks written
straight into the final matrix, sharded covariance precompute, in-place all-gathers, tiled mirror pass) and the
row-sharded Nystroem transform must equal the single-GPU results bit for bit.  Run by tests/test_gpu_multi_rank.py
when >= 2 GPUs are visible; the output of the last hand run is kept under profiles/."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matchcake_b200 import ops as mops
from matchcake_b200.ml.kernels import FermionicPQCKernel, Nystroem
rank = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
# (n_qubits, n_features, n_samples): dense tile path (d = 64, in-place folded gather when ns % 2ws == 0 and the padded
# fallback otherwise), small registers, the block-factorised path (depth 1) and the >64-qubit block-tile path
for n, nf, ns in [(32, 64, 64 * ws), (32, 64, 1000), (8, 8, 37), (6, 12, 5), (16, 16, 16 * ws + 3), (40, 80, 8 * ws)]:
    x = np.random.default_rng(0).random((ns, nf))
    k = FermionicPQCKernel(n_qubits=n, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
    k.float32_gram = False
    k.fit(x)
    xd = torch.as_tensor(x).cuda()
    sharded = k(xd)
    blocks = k._block_plan() is not None and k.use_block_factorisation
    if not blocks:
        cov = mops.cov_basis(k._x_to_sptm(xd), n)
        single = mops.gram(cov, cov, 2.0 ** -n, floor2=k.pfaffian_epsilon)
    else:
        b = k._x_to_blocks(xd); single = mops.gram_blocks4(b, b, floor2=k.pfaffian_epsilon)
    err = float((sharded - single).abs().max())
    # rectangular (Nystroem transform shape): gathered, and left row-sharded
    m = max(2, ns // 3)
    rect = k(xd, xd[:m])
    if not blocks:
        rect_single = mops.gram(cov, cov[:m].contiguous(), 2.0 ** -n, floor2=k.pfaffian_epsilon)
    else:
        rect_single = mops.gram_blocks4(b, b[:m].contiguous(), floor2=k.pfaffian_epsilon)
    err_rect = float((rect - rect_single).abs().max())
    loc, (s, e) = k(xd, xd[:m], gather=False)
    err_loc = float((loc - rect_single[s:e]).abs().max()) if e > s else 0.0
    if rank == 0:
        print(f"N={n} ns={ns} ws={ws}: max |sharded - single| square {err:.1e}, rect {tuple(rect.shape)} {err_rect:.1e}, "
              f"row-sharded rect {err_loc:.1e}")
    assert err == 0.0 and err_rect == 0.0 and err_loc == 0.0, (err, err_rect, err_loc)
# Nystroem: sharded transform == single-rank transform rows
x = np.random.default_rng(1).random((40 * ws, 16))
kern = FermionicPQCKernel(n_qubits=8, rotations="X,Z", r_dtype=torch.float64, c_dtype=torch.complex128)
kern.float32_gram = False
ny = Nystroem(kern, n_components=12, random_state=0).fit(x)
full = ny.transform(x)
emb, (s, e) = ny.transform_sharded(x)
err = float(np.abs(emb.cpu().numpy() - full[s:e]).max()) if e > s else 0.0
assert err < 1e-12, err
if rank == 0:
    print(f"Nystroem transform_sharded rows [{s},{e}) vs gathered transform: max diff {err:.1e}")
dist.barrier(); dist.destroy_process_group()
if rank == 0:
    print("dist gram check ok")
