"""Stand-alone launcher of single conv shapes for ncu (not a test):  python tests/prof_conv.py fwd|wgrad"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
n, d, h, w, cin, cout = 288, 1, 128, 128, 96, int(sys.argv[2]) if len(sys.argv) > 2 else 96
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
x = torch.randn(n, d, h, w, cin, device="cuda"); be.round_tf32(x)
if mode == "fwd":
    wp = torch.randn(9 * cout * cin, device="cuda") / 30; be.round_tf32(wp)
    bias = torch.randn(cout, device="cuda"); scale = torch.rand(18, cout, device="cuda") + 0.5
    res = torch.randn(n, d, h, w, cout, device="cuda"); y = torch.empty(n, d, h, w, cout, device="cuda")
    for _ in range(3):
        be.conv_fwd(x, wp, bias, scale, res, y, n, d, h, w, cin, cout, 1, 3, 3, 18, 0, algo=algo)
else:
    dz = torch.randn(n, d, h, w, cout, device="cuda"); be.round_tf32(dz)
    dwp = torch.empty(9 * cout * cin, device="cuda")
    for _ in range(3):
        be.conv_wgrad(x, dz, dwp, n, d, h, w, cin, cout, 1, 3, 3)
torch.cuda.synchronize()
print("done", mode)
