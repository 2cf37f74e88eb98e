"""Quick device timing of individual C-ABI calls (CUDA events).  Development aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mal_b200 import _capi, raw
from mal_b200.utils.synthetic import make_photometric_inputs, to_device

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    h = _capi.lib()
    _capi.check(h.mal_check_device(0))
    dev = torch.device("cuda:0")
    inputs, t = make_photometric_inputs(B, 192, 640, seed=1)
    inputs, t = to_device(inputs, dev), to_device(t, dev)
    tgt = inputs[("color", 0, 0)]; src = [inputs[("color", -1, 0)], inputs[("color", 1, 0)]]
    ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
    res = {}
    res["identity(pred,2cand,nograd)"] = timeit(lambda: raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False))
    common = dict(target=tgt, src=src, depth=t[("mono_disp", 0)], K=inputs[("K", 0)], inv_K=inputs[("inv_K", 0)],
                  T=[t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)]])
    syn = [t[("syn", -1, 0)], t[("syn", 1, 0)]]
    res["warp,2cand,nograd"] = timeit(lambda: raw.photo(h, **common))
    res["warp,2cand,grad,automask"] = timeit(lambda: raw.photo(h, **common, identity_min=ident, noise=t["noise"][0], with_grad=True))
    res["warp,4cand,grad,automask"] = timeit(lambda: raw.photo(h, **common, syn=syn, identity_min=ident, noise=t["noise"][0], with_grad=True))
    res["warp,4cand,nograd"] = timeit(lambda: raw.photo(h, **common, syn=syn))
    for k, v in res.items():
        print(f"{k:34s} {v:9.1f} us/call  {v / B:7.2f} us/frame")
    print(json.dumps({"B": B, "us": res}))

if __name__ == "__main__":
    main()
