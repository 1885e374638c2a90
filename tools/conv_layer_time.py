"""Time ONE conv layer's tcgen05 forward kernel in isolation (CUDA events around the C-ABI unit entry are dominated by
the layout converters, so run this under `ncu --metrics gpu__time_duration.sum -k regex:conv_tc_kernel`).
usage: python tools/conv_layer_time.py N CIN COUT H W K [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200"))
import torch  # noqa: E402

from fsb200.runtime import conv_forward  # noqa: E402

n, cin, cout, h, w, k = (int(a) for a in sys.argv[1:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
g = torch.Generator().manual_seed(0)
x = torch.randn(n, cin, h, w, generator=g).cuda()
wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
b = torch.randn(cout, generator=g).cuda()
for _ in range(reps):
    y = conv_forward(x, wt, b, "bf16x3")
torch.cuda.synchronize()
print("ok", tuple(y.shape))
