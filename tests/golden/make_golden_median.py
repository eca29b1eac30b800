"""Geometric-median fixture from the UNMODIFIED reference (sae/training/geometric_median.py:23-85); run in the build container only.

    python tests/golden/make_golden_median.py       -> tests/golden/geometric_median.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, HERE)
    import _ref_shims
    _ref_shims.install()
    from vit_prisma.sae.training.geometric_median import compute_geometric_median
    cases = []
    for seed, n, d, outliers, maxiter in ((0, 257, 16, 0, 100), (1, 512, 48, 40, 100), (2, 64, 8, 8, 5), (3, 300, 32, 30, 200)):
        g = torch.Generator().manual_seed(seed)
        pts = torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)
        if outliers:
            pts[:outliers] += 25.0 * torch.randn(outliers, d, generator=g)          # heavy outliers: mean and median differ a lot
        out = compute_geometric_median(pts, maxiter=maxiter)
        cases.append(dict(seed=seed, n=n, d=d, outliers=outliers, maxiter=maxiter, median=out.median.clone(), mean=pts.mean(0),
                          termination=out.termination))
        print(seed, n, d, out.termination, float((out.median - pts.mean(0)).norm()))
    torch.save(cases, os.path.join(HERE, "geometric_median.pt"))


if __name__ == "__main__":
    main()
