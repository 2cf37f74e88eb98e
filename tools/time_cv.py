"""Device timing of the ManyDepth cost-volume head at the bench shape (B x 1 lookup x 96 bins x 64 ch at 48x160),
CUDA events, rotating over 3 input sets.  A/B switches: MAL_CV_MINB, MAL_CV_KERNEL=lane, MAL_B200_LIB=<other build>.
python tools/time_cv.py [B]"""
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

    with torch.no_grad():
        for i in range(6):
            cv(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            cv(i)
        e1.record()
        torch.cuda.synchronize()
    print("cv_pack (lookup) + cv_sweep_quad_kernel  %8.1f us" % (e0.elapsed_time(e1) / 30 * 1e3), flush=True)


if __name__ == "__main__":
    main()
