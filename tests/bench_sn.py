"""Micro-benchmark of the spectral-norm power-iteration kernel (not a test): time vs number of folded calls G."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
for (R, K) in [(768, 6912), (384, 10368), (384, 3456), (192, 1728), (96, 864), (48, 432), (768, 768)]:
    w = torch.randn(R, K, device="cuda") / K ** 0.5
    row = []
    for G in (1, 4, 18):
        u = torch.nn.functional.normalize(torch.randn(R, device="cuda"), dim=0)
        v = torch.nn.functional.normalize(torch.randn(K, device="cuda"), dim=0)
        isg, uh, vh = torch.empty(G, device="cuda"), torch.empty(G, R, device="cuda"), torch.empty(G, K, device="cuda")
        ws = torch.zeros((G + 2) * R + 2 * G + 8, device="cuda")
        for _ in range(3):
            be.sn_power_iter(w, u, v, R, K, G, 1e-4, True, isg, uh, vh, ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            be.sn_power_iter(w, u, v, R, K, G, 1e-4, True, isg, uh, vh, ws)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 10 * 1e3)
    print(f"sn {R}x{K}: G=1 {row[0]:.0f} us, G=4 {row[1]:.0f} us, G=18 {row[2]:.0f} us  -> per extra iteration {(row[2]-row[0])/17:.1f} us")
