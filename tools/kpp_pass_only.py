"""One K-means++ distance pass per shape on generated data (the target of the ncu capture in profiles/)."""
import sys

sys.path.insert(0, ".")
from ml_b200 import cabi

ctx = cabi.Context(1)
for n, d in [(12_500_000, 16), (2_500_000, 64)]:
    data = cabi.Data.generate_gmm(ctx, n, d, 8, seed=3)
    c = data.download(0, 3)
    for i in range(3):
        data.kpp_update(c[i], first=(i == 0), want_nearest=False)
    ctx.synchronize()
    data.close()
ctx.close()
