"""One launch of each hot kernel at its BASELINE-config-3 shape (batch 16), for `ncu --set full` captures:
   k_conv3r (48->48 @512^2 plain and with GroupNorm+SiLU applied in place), k_conv3 (96->48 @512^2 4-row streamed, 96->96 @256^2,
   144->144 @128^2, 192->96 @256^2), k_conv3w (192->192, 384->192 @64^2), k_conv1 (96->48 @512^2, 192->576 @64^2),
   k_attn_tc (2 heads x d=96, 4096 tokens), k_gn_act2 (48 ch @512^2)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_checks import OpHandle, DEV  # noqa: E402
B = 16
oh = OpHandle("fp16")
for (cin, hw, cout, k, impl) in [(48, 512, 48, 3, 11), (48, 512, 48, 3, 12), (96, 512, 48, 3, 2), (96, 256, 96, 3, 2), (144, 128, 144, 3, 2),
                                 (192, 256, 96, 3, 2), (192, 64, 192, 3, 7), (384, 64, 192, 3, 7), (96, 512, 48, 1, 5), (192, 64, 576, 1, 5)]:
    x = torch.randn(B, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    oh.conv2d(x, w, b, k, 1, k // 2, impl)
    del x, w
qkv = torch.randn(B, 576, 64, 64, device=DEV)
oh.attention(qkv, 2, 96, 1)
x = torch.randn(B, 48, 512, 512, device=DEV)
oh.groupnorm_act(x, torch.ones(48, device=DEV), torch.zeros(48, device=DEV), 8, 1)
torch.cuda.synchronize()
print("done")
