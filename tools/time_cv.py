"""Device timing of the ManyDepth cost-volume head at the bench shape (B x 1 lookup x 96 bins x 64 ch at 48x160),
CUDA events, rotating over 3 input sets.  A/B switches: MAL_CV_NO_DESC=1 (projections inside the sweep),
MAL_CV_MINB, MAL_CV_KERNEL=lane.  python tools/time_cv.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw, step as S
from mal_b200.utils.synthetic import to_device


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    h = _capi.lib()
    dev = torch.device("cuda:0")
    opt = S.default_opt(B)
    bufs = [to_device(S.synthetic_batch(opt, seed=1234 + 17 * i), dev) for i in range(3)]

    def cv(i):
        b = bufs[i % 3]
        return raw.cost_volume(h, current=b["current_feats"], lookup=b["lookup_feats"], poses=b["relative_poses"], K=b["K2"],
                               inv_K=b["inv_K2"], bins=b["bins"], apply_confidence=True, want_missing=False)

    ref = None
    for name, env in (("projection pre-pass (cv_desc_kernel) + sweep", {}), ("projections inside the sweep", {"MAL_CV_NO_DESC": "1"})):
        os.environ.update(env)
        with torch.no_grad():
            for i in range(6):
                out = cv(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(30):
                out = cv(i)
            e1.record()
            torch.cuda.synchronize()
        for k in env:
            del os.environ[k]
        vol = cv(0)["cost_volume"] if not env else None
        print("%-50s %8.1f us" % (name, e0.elapsed_time(e1) / 30 * 1e3), flush=True)
        if ref is None:
            ref = cv(0)
        else:
            os.environ.update(env)
            other = cv(0)
            for k in env:
                del os.environ[k]
            same = all(torch.equal(ref[k], other[k]) for k in ("cost_volume", "confidence", "argmin", "lowest_cost"))
            print("both paths give the same bits:", same)


if __name__ == "__main__":
    main()
