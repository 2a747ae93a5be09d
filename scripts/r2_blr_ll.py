import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import strumpack_b200 as sb
from strumpack_b200.fronts import laplacian_root_front
for k in (127, 181):
    F, _ = laplacian_root_front(k, 256, device="cuda")
    Fh = np.asfortranarray(F.cpu().numpy())
    del F
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-4, abs_tol=1e-12, leaf_size=256)
    for alg, name in ((sb.BLR_RL, "RL"), (sb.BLR_LL, "LL")):
        sb.BLRMatrix.compress_and_factor(Fh, o, factor_algorithm=alg)
        t0 = time.perf_counter()
        B = sb.BLRMatrix.compress_and_factor(Fh, o, factor_algorithm=alg)
        t = time.perf_counter() - t0
        print(json.dumps({"k": k, "alg": name, "host_path_s": t, "rank": B.rank, "launches": B.launches}), flush=True)
