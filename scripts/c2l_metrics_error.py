import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import xgtest
import test_apply_gpu as T
pkg = xgtest.package()
for ni in (16, 48):
    c = T.Case(pkg, ni, 90, 45, 2)
    p = c.plan
    p.grad_setup(c.xt, c.yt)
    worst = {}; same = {}
    for t, m in enumerate(c.oracle_metrics()):
        got = p.grad_get_metrics(t)
        for k in xgtest.METRICS:
            scale = np.max(np.abs(m[k]))
            worst[k] = max(worst.get(k, 0.0), float(np.max(np.abs(got[k] - m[k])) / scale))
            same[k] = min(same.get(k, 1.0), float(np.mean(got[k] == m[k])))
    print(ni, {k: "%.2e" % v for k, v in worst.items()})
    print(ni, "fraction bit-identical (worst tile):", {k: round(v, 4) for k, v in same.items()})
