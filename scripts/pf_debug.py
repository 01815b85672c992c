import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import pfaffian
from matchcake_b200 import ops as mops
from tests.helpers import random_skew
for m in (14, 22, 30, 32, 46, 62, 64):
    rng = np.random.default_rng(m)
    a = random_skew(rng, (200, m, m)) / np.sqrt(m)
    ref = pfaffian.pfaffian_signed(a)
    out = mops.pfaffian(torch.as_tensor(a).cuda(), signed=True).cpu().numpy()
    rel = np.abs(out - ref) / np.abs(ref)
    bad = np.argsort(-rel)[:4]
    print(m, 'max rel', rel.max(), 'bad', [(int(i), float(rel[i]), float(out[i]), float(ref[i])) for i in bad])
