"""per-call wall time of xnb_step_host at C2 for a few chunk counts (non-rebuild and rebuild steps separately)"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import parity_util as U

kw, _ = bench.workload("C2")
eps, sig, rc, dt = kw["epsilon"], kw["sigma"], kw["rcut"], kw["dt"]
for chunks in sys.argv[1:]:
    os.environ["XNB_HOST_CHUNKS"] = chunks
    ctx = U.make_ctx(kw, device=0)
    sh = torch.cuda.current_stream().cuda_stream
    ctx.first_iteration(eps, sig, rc, sh)
    n = ctx.n_inner
    hb = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")}
    hid = torch.empty(n, dtype=torch.int64).pin_memory()
    p = {k: v.data_ptr() for k, v in hb.items()}
    ctx.download_rvf(p["rx"], p["ry"], p["rz"], p["vx"], p["vy"], p["vz"], p["fx"], p["fy"], p["fz"], hid.data_ptr(), sh)
    t_plain, t_reb = [], []
    for it in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rb = ctx.step_host(dt, eps, sig, rc, in_r=(p["rx"], p["ry"], p["rz"]), in_v=(p["vx"], p["vy"], p["vz"]), out_r=(p["rx"], p["ry"], p["rz"]),
                           out_v=(p["vx"], p["vy"], p["vz"]), out_f=(p["fx"], p["fy"], p["fz"]), out_id=hid.data_ptr(), stream=sh)
        t = (time.perf_counter() - t0) * 1e3
        if it >= 3:
            (t_reb if rb else t_plain).append(t)
    print("chunks %s: plain step %.3f ms (min %.3f, n=%d), rebuild step %.3f ms (n=%d)" % (chunks, np.median(t_plain), np.min(t_plain), len(t_plain), np.median(t_reb) if t_reb else 0, len(t_reb)), flush=True)
    ctx.close()
