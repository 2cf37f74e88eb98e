"""Device timing of the four photometric passes of a MAL step at the bench shape (CUDA events, rotating over 3
input sets > L2).  MAL_B200_LIB=<other .so> times another build of the library for A/B comparisons."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw, step as S
from mal_b200.utils.synthetic import to_device


def timeit(fn, iters=40, warm=6):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    tag = sys.argv[2] if len(sys.argv) > 2 else "lib"
    h = _capi.lib()
    dev = torch.device("cuda:0")
    opt = S.default_opt(B)
    bufs = [to_device(S.synthetic_batch(opt, seed=1234 + 17 * i), dev) for i in range(3)]
    ident = [raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], mode=raw.PHOTO_PRED,
                       want_selection=False)["min_reproj"] for b in bufs]
    masks = [(b["noise_main"][:, 0] > 0).float() for b in bufs]
    smz = [b["augmentation_mask"].reshape(-1) * 0 for b in bufs]

    def k_ident(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], mode=raw.PHOTO_PRED, want_selection=False,
                  finalize=False)

    def k_teacher(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], syn=[b["syn_-1"], b["syn_1"]],
                  depth=b["mono_disp"], K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]],
                  identity_min=ident[i % 3], noise=b["noise_mono"], with_grad=True)

    def k_ens(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["mono_disp"], depth_b=b["multi_disp"],
                  K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]], want_selection=False, finalize=False)

    def k_student(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["multi_disp"], K=b["K"],
                  inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]], pixel_mask=masks[i % 3], sample_mask=smz[i % 3],
                  with_grad=True)

    fin = os.environ.get("MAL_TIME_NOFIN") is None

    def k_teacher_nf(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], syn=[b["syn_-1"], b["syn_1"]],
                  depth=b["mono_disp"], K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]],
                  identity_min=ident[i % 3], noise=b["noise_mono"], with_grad=True, finalize=False)

    def k_student_nf(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["multi_disp"], K=b["K"],
                  inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]], pixel_mask=masks[i % 3], sample_mask=smz[i % 3],
                  with_grad=True, finalize=False)

    def k_cv(i):
        b = bufs[i % 3]
        raw.cost_volume(h, current=b["current_feats"], lookup=b["lookup_feats"], poses=b["relative_poses"], K=b["K2"],
                        inv_K=b["inv_K2"], bins=b["bins"], apply_confidence=True, want_missing=False)

    def k_smooth(i):
        b = bufs[i % 3]
        raw.smooth(h, disp=b["mono_disp"], img=b["color_0"], normalise=True, with_grad=True)

    res = {}
    with torch.no_grad():
        for name, fn in (("identity", k_ident), ("teacher", k_teacher), ("ensemble", k_ens), ("student", k_student)):
            res[name] = round(timeit(fn), 1)
        res["sum"] = round(sum(res.values()), 1)
        if fin:
            res["teacher_nofin"] = round(timeit(k_teacher_nf), 1)
            res["student_nofin"] = round(timeit(k_student_nf), 1)
        res["cv"] = round(timeit(k_cv), 1)
        res["smooth1"] = round(timeit(k_smooth), 1)
    print(tag, json.dumps(res))


if __name__ == "__main__":
    main()
