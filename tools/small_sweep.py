import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pspace_b200 import workloads as WL
from pspace_b200.device import DeviceContainer, fill_synthetic
def t(fn, inner=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best
cases = [(3,3,0),(4,3,1),(3,4,0),(5,3,1),(4,3,0),(5,4,1),(6,3,1)]
for nv, dm, bt in cases:
    c = DeviceContainer(WL.mixed_params(nv, dm), bt, False).initialize()
    N, Q = c.N, c.Q
    for m2 in (1, 64, 576):
        J = fill_synthetic(Q * m2, 7).reshape(Q, m2)
        out = torch.empty((N, N, m2), dtype=torch.float64, device='cuda')
        res = {}
        for cache in (0, 1):
            c.setOption("panel_cache", cache)
            for force in (0, 1):
                c.setOption("force_generic", force)
                res[(cache, force)] = t(lambda: c.projectMatrix(J, out=out))
                if cache == 0 and force == 0: plan = c.lastPlan()[:60]
        flop = 2.0 * N * N * Q * m2
        print(f"nv{nv} d{dm} bt{bt} N={N} Q={Q} m2={m2} flop={flop:.3g}: tma {res[(0,0)]:.4f} tma_cached {res[(1,0)]:.4f} generic {res[(0,1)]:.4f} | {plan}", flush=True)
    c.setOption("force_generic", 0); c.setOption("panel_cache", 0)
    # vector too
    for m in (1, 24, 576):
        f = fill_synthetic(Q * m, 9).reshape(Q, m)
        tv = t(lambda: c.projectVector(f))
        print(f"   vector m={m}: {tv:.4f} ms | {c.lastPlan()[:70]}", flush=True)
    c.close()
