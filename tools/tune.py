"""Device timing of individual C-ABI calls at the bench shape (CUDA events, GPU-paced loops)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw, step as S
from mal_b200.utils.synthetic import to_device


def timeit(fn, iters=30, warm=5):
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
    h = _capi.lib()
    dev = torch.device("cuda:0")
    opt = S.default_opt(B)
    bufs = [to_device(S.synthetic_batch(opt, seed=1234 + 17 * i), dev) for i in range(3)]
    res = {}
    ident = [raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], mode=raw.PHOTO_PRED,
                       want_selection=False)["min_reproj"] for b in bufs]

    def cv(i):
        b = bufs[i % 3]
        raw.cost_volume(h, current=b["current_feats"], lookup=b["lookup_feats"], poses=b["relative_poses"], K=b["K2"],
                        inv_K=b["inv_K2"], bins=b["bins"], apply_confidence=True, want_missing=False)

    only = sys.argv[2] if len(sys.argv) > 2 else ""
    if only.startswith("cv"):
        os.environ["MAL_CV_MINB"] = only[2:] or "4"
        print("cv", timeit(cv, iters=3, warm=1))
        return
    for kern in ("quad", "lane"):
        os.environ["MAL_CV_KERNEL"] = kern
        for minb in ("3", "4", "5"):
            os.environ["MAL_CV_MINB"] = minb
            res["cost_volume %s minb=%s" % (kern, minb)] = timeit(cv)
    os.environ.pop("MAL_CV_MINB")
    os.environ.pop("MAL_CV_KERNEL")

    def k_ident(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], mode=raw.PHOTO_PRED, want_selection=False)

    def k_teacher(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], syn=[b["syn_-1"], b["syn_1"]],
                  depth=b["mono_disp"], K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]],
                  identity_min=ident[i % 3], noise=b["noise_mono"], with_grad=True)

    def k_ens(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["mono_disp"], K=b["K"],
                  inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]], want_selection=False)

    def k_student(i):
        b = bufs[i % 3]
        raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["multi_disp"], K=b["K"],
                  inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]], pixel_mask=(b["noise_main"][:, 0] > 0).float(),
                  sample_mask=b["augmentation_mask"].reshape(-1) * 0, with_grad=True)

    def k_smooth(i):
        b = bufs[i % 3]
        raw.smooth(h, disp=b["mono_disp"], img=b["color_0"], normalise=True, with_grad=True)

    def k_main(i):
        b = bufs[i % 3]
        raw.main_terms(h, multi=b["multi_disp"], mono=b["mono_disp"], pixel_mask=b["noise_main"][:, 0],
                       mono_reproj=ident[0], ens_reproj=ident[1], multi_reproj=ident[2], inputs_are_disp=True,
                       with_grad=True)

    with torch.no_grad():
        for name, fn in (("photo identity (PRED,2)", k_ident), ("photo teacher (WARP,4,grad,automask)", k_teacher),
                         ("photo ensemble (WARP,2)", k_ens), ("photo student (WARP,2,grad,masks)", k_student),
                         ("smooth fwd+bwd", k_smooth), ("main_terms fwd+bwd", k_main)):
            res[name] = timeit(fn)
    for k, v in res.items():
        print(f"{k:44s} {v:9.1f} us/call  {v / B:7.2f} us/frame")
    print(json.dumps({"B": B, "us": res}))


if __name__ == "__main__":
    main()
