"""Operand builders shared by the profiling helpers (prof_kernels.py, prof_kernels2.py); not a test."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
torch.manual_seed(0)


def conv(n, d, h, w, cin, cout, kd, k, g, res=True, act=0, algo=0):
    taps = kd * k * k
    x = torch.randn(n, d, h, w, cin, device="cuda"); wp = torch.randn(taps * cout * cin, device="cuda") / 30
    be.round_tf32(x); be.round_tf32(wp)
    bias = torch.randn(cout, device="cuda"); scale = torch.rand(g, cout, device="cuda") + 0.5
    r = torch.randn(n, d, h, w, cout, device="cuda") if res else None
    y = torch.empty(n, d, h, w, cout, device="cuda")
    return lambda: be.conv_fwd(x, wp, bias, scale, r, y, n, d, h, w, cin, cout, kd, k, k, g, act, algo=algo)


def upconv(n, h, w, cin, cout, g):
    x = torch.randn(n, 1, h, w, cin, device="cuda"); wsp = torch.randn(16 * cout * cin, device="cuda") / 30
    be.round_tf32(x); be.round_tf32(wsp)
    bias = torch.randn(cout, device="cuda"); scale = torch.rand(g, cout, device="cuda") + 0.5
    y = torch.empty(n, 1, 2 * h, 2 * w, cout, device="cuda")
    dz = torch.randn(n, 1, 2 * h, 2 * w, cout, device="cuda"); be.round_tf32(dz)
    dx = torch.empty(n, 1, h, w, cin, device="cuda")
    return (lambda: be.upconv_fwd(x, wsp, bias, scale, None, y, n, h, w, cin, cout, g, 0),
            lambda: be.upconv_dgrad(dz, wsp, dx, n, h, w, cin, cout))


def wgrad(n, d, h, w, cin, cout, kd, k):
    x = torch.randn(n, d, h, w, cin, device="cuda"); dz = torch.randn(n, d, h, w, cout, device="cuda")
    be.round_tf32(x); be.round_tf32(dz)
    dwp = torch.empty(kd * k * k * cout * cin, device="cuda")
    return lambda: be.conv_wgrad(x, dz, dwp, n, d, h, w, cin, cout, kd, k, k)


