import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from ml_b200 import cabi
n, d, k, steps = 10_000_000, 8, 16, 20
ctx = cabi.Context(1)
gen = cabi.Data.generate_gmm(ctx, n, d, k, seed=1)
host = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
hn = host.numpy(); hn[:] = gen.download(0, n); gen.close()
init = np.ascontiguousarray(hn[:k].T)
pageable = hn.copy()
for name, buf in (("pinned", hn), ("pageable", pageable), ("pinned", hn)):
    t = [time.perf_counter()]
    d2 = cabi.Data.upload(ctx, buf); ctx.synchronize(); t.append(time.perf_counter())
    em = cabi.Em(d2, k); t.append(time.perf_counter())
    cov = em.sample_covariance(); t.append(time.perf_counter())
    em.set_params(init, np.repeat(cov[None], k, axis=0), np.full(k, 1.0 / k)); t.append(time.perf_counter())
    for _ in range(steps): ll = em.step()
    t.append(time.perf_counter())
    p = em.get_params(); t.append(time.perf_counter())
    _, labels = em.emit(want_responsibilities=False, want_labels=True); t.append(time.perf_counter())
    em.close(); d2.close(); t.append(time.perf_counter())
    names = ["upload+shift", "em_create", "sample_cov", "set_params", f"{steps} steps", "get_params", "emit labels", "close"]
    print(name, {a: round((t[i + 1] - t[i]) * 1e3, 2) for i, a in enumerate(names)}, "total", round((t[-1] - t[0]) * 1e3, 1), flush=True)
