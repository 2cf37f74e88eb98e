"""Device timing of the DynamicDepth cost-volume variants at the Cityscapes bench shape (B x 2 lookups x 96 bins x
64 ch at 48x128): which part of the pool path costs what.  python tools/time_cv_dyn.py [B] [case substring]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw
from mal_b200.utils.synthetic import CITYSCAPES_K, make_cost_volume_inputs


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    only = sys.argv[2] if len(sys.argv) > 2 else ""   # substring of the case name
    h = _capi.lib()
    dev = torch.device("cuda:0")
    cvd = make_cost_volume_inputs(B, 192, 512, channels=64, num_lookup=2, num_bins=96, seed=9, min_bin=0.5, max_bin=20.0,
                                  translation_scale=0.5, normalised_K=CITYSCAPES_K)
    cvd = {k: v.to(dev) for k, v in cvd.items()}
    occ = torch.zeros(B, 48, 128)
    occ[:, 12:30, 40:80] = 1.0
    occ = occ.to(dev)
    aug = torch.zeros(B, 1, 1, 1, device=dev)
    base = dict(current=cvd["current_feats"], lookup=cvd["lookup_feats"], poses=cvd["relative_poses"], K=cvd["K"],
                inv_K=cvd["inv_K"], bins=cvd["bins"])
    cases = [
        ("plain (ManyDepth mean over frames), quad kernel", dict(), {}),
        ("plain, lane kernel", dict(), {"MAL_CV_KERNEL": "lane"}),
        ("cv_min, no occlusion fill", dict(cv_min=True), {}),
        ("cv_min + set_1", dict(cv_min=True, occ=occ, occ_mode=raw.OCC_SET_1, aug_mask=aug), {}),
        ("cv_min + pool, descriptor volume", dict(cv_min=True, occ=occ, occ_mode=raw.OCC_POOL, pool_radius=1, pool_th=0.7, aug_mask=aug), {}),
    ]
    with torch.no_grad():
        for name, kw, env in cases:
            if only not in name:
                continue
            os.environ.update(env)
            us = timeit(lambda: raw.cost_volume(h, **base, **kw))
            for k in env:
                del os.environ[k]
            print("%-50s %9.1f us" % (name, us), flush=True)


if __name__ == "__main__":
    main()
