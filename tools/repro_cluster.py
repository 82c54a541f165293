import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_codes_b200 import convops as co
from tools.bringup_conv import _pack_img, _nhwc, _rel
import torch.nn.functional as F
torch.manual_seed(0)
N, C, Kc, H = 5, 256, 256, 4
x = torch.randn(N, C, H, H, device="cuda").bfloat16()
w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
ref = F.conv2d(x.float(), w.float(), None, padding=1)
wp, keep = _pack_img(w)
o1 = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
co.conv_run([_nhwc(x)], co.taps_3x3(Kc), wp, Kc, o1)
torch.cuda.synchronize()
print("rel", _rel(o1.permute(0, 3, 1, 2), ref))
