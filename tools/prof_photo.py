"""Launch each photometric pass of a MAL step a few times at the bench shape (for ncu: -k regex:photo_kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw, step as S
from mal_b200.utils.synthetic import to_device

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = _capi.lib()
dev = torch.device("cuda:0")
opt = S.default_opt(B)
b = to_device(S.synthetic_batch(opt, seed=1234), dev)
src = [b["color_-1"], b["color_1"]]
geom = dict(K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"], b["T_1"]])
mask = (b["noise_main"][:, 0] > 0).float()
with torch.no_grad():
    for _ in range(reps):
        ident = raw.photo(h, target=b["color_0"], src=src, mode=raw.PHOTO_PRED, want_selection=False, finalize=False)["min_reproj"]
        raw.photo(h, target=b["color_0"], src=src, syn=[b["syn_-1"], b["syn_1"]], depth=b["mono_disp"], identity_min=ident,
                  noise=b["noise_mono"], with_grad=True, **geom)
        raw.photo(h, target=b["color_0"], src=src, depth=b["mono_disp"], depth_b=b["multi_disp"], want_selection=False,
                  finalize=False, **geom)
        raw.photo(h, target=b["color_0"], src=src, depth=b["multi_disp"], pixel_mask=mask,
                  sample_mask=b["augmentation_mask"].reshape(-1) * 0, with_grad=True, **geom)
        raw.smooth(h, disp=b["mono_disp"], img=b["color_0"], normalise=True, with_grad=True, disp_b=b["multi_disp"], defer_fix=True)
torch.cuda.synchronize()
print("ok")
