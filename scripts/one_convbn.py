"""One launch of the tcgen05 conv+BN+ReLU unit at a named Inception geometry (for `ncu --set full`).
    python scripts/one_convbn.py <unit name> [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from littlegan_b200 import kernels as K  # noqa: E402
from littlegan_b200.inception import unit_specs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Mixed_6e.branch7x7dbl_2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
u = [s for s in unit_specs() if s["name"] == name][0]
size = {"Conv2d_1a": 299, "Conv2d_2a": 149, "Conv2d_2b": 147, "Conv2d_3b": 73, "Conv2d_4a": 73, "Mixed_5": 35,
        "Mixed_6a": 35, "Mixed_6": 17, "Mixed_7a": 17, "Mixed_7": 8}
H = [v for k, v in size.items() if name.startswith(k)][0]
if name.startswith("Mixed_6a") and u["cin"] != 288 and u["s"] == 1:
    H = 35
cin = max(u["cin"], 8)
x = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
W = torch.randn(u["k"][0], u["k"][1], cin, u["cout"], device="cuda") * 0.05
wp = K.pack_conv_bn_weights(W)
scale, shift = torch.ones(u["cout"], device="cuda"), torch.zeros(u["cout"], device="cuda")
Ho = (H + 2 * u["p"][0] - u["k"][0]) // u["s"] + 1
Wo = (H + 2 * u["p"][1] - u["k"][1]) // u["s"] + 1
y = torch.empty(B, Ho, Wo, u["cout"], device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    K.conv2d_bn_relu(x, W, scale, shift, y, 0, stride=u["s"], pad=u["p"], wpack=wp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    K.conv2d_bn_relu(x, W, scale, shift, y, 0, stride=u["s"], pad=u["p"], wpack=wp)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 5 * 1e3
M, Kd = B * Ho * Wo, u["k"][0] * u["k"][1] * u["cin"]
print("%s batch %d: M %d K %d N %d  %.1f us  %.1f TFLOP/s" % (name, B, M, Kd, u["cout"], us, 2.0 * M * Kd * u["cout"] / us / 1e6))
