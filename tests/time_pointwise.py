"""HBM-bound passes of the step at the benchmark's shapes (tuning helper, not a test): ms and effective TB/s (bytes the pass must move)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
RND = _lib.FLAG_ROUND_TF32


def timeit(f, n=5):
    for _ in range(2): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(name, ms, nbytes):
    print(f"{name:64s} {ms:7.3f} ms  {nbytes / ms / 1e9:6.2f} TB/s", flush=True)


def R(*s):
    return torch.randn(*s, device="cuda")


# ---- conv_bwd_prep
for (rows, G, C, geom) in [(2883584, 1, 48, (32, 22, 64, 64)), (65536, 18, 96, None), (131072, 8, 48, None), (262144, 18, 48, None), (16384, 18, 192, None),
                           (360448, 1, 96, (32, 11, 32, 32))]:
    n = rows * G * C
    dy, y, res = R(rows * G, C), R(rows * G, C), R(rows * G, C)
    dz, dpre = torch.empty_like(dy), torch.empty_like(dy)
    bias, scale = R(C), torch.rand(G, C, device="cuda") + 0.5
    dbias, dscale = torch.zeros(C, device="cuda"), torch.zeros(G, C, device="cuda")
    ms = timeit(lambda: be.conv_bwd_prep(dy, y, None, bias, scale, dz, None, dbias, dscale, rows, G, C, 1 | RND))
    report(f"prep relu+dscale rows{rows} G{G} C{C}", ms, 12.0 * n)
    ms = timeit(lambda: be.conv_bwd_prep(dy, y, res, bias, scale, dz, None, dbias, dscale, rows, G, C, 0 | RND))
    report(f"prep res+dscale rows{rows} G{G} C{C}", ms, 16.0 * n)
    ms = timeit(lambda: be.conv_bwd_prep(dy, None, None, bias, scale, dz, None, dbias, None, rows, G, C, 0 | RND))
    report(f"prep scale only rows{rows} G{G} C{C}", ms, 8.0 * n)
    if geom:
        N, D, H, W = geom
        dyp = R(N, D // 2, H // 2, W // 2, C)
        ms = timeit(lambda: be.conv_bwd_prep(dyp, y, None, bias, scale, dz, None, dbias, dscale, rows, G, C, 0 | RND, pool=(2, 2, 2, D, H, W)))
        report(f"prep pooled+dscale rows{rows} G{G} C{C}", ms, 8.5 * n)
        ms = timeit(lambda: be.conv_bwd_prep(dyp, None, None, bias, scale, dz, None, dbias, None, rows, G, C, 0 | RND, pool=(2, 2, 2, D, H, W)))
        report(f"prep pooled, no y rows{rows} G{G} C{C}", ms, 4.5 * n)
    del dy, y, res, dz, dpre

# ---- BatchNorm
for (rows, G, C, H, W) in [(262144, 18, 96, 128, 128), (65536, 18, 96, 64, 64), (65536, 18, 192, 64, 64), (16384, 18, 192, 32, 32), (4096, 18, 384, 16, 16), (1024, 18, 768, 8, 8)]:
    n = rows * G * C
    x, dy, add = R(rows * G, C), R(rows * G, C), R(rows * G, C)
    dx = torch.empty_like(x)
    sums = torch.zeros(G, C, 2, device="cuda", dtype=torch.float64)
    red = torch.zeros(G, C, 2, device="cuda", dtype=torch.float64)
    a, b, mean, invstd = (torch.rand(G, C, device="cuda") + 0.5 for _ in range(4))
    os_ = torch.rand(G, C, device="cuda") + 0.5
    report(f"bn_stats rows{rows} G{G} C{C}", timeit(lambda: be.bn_stats(x, sums, rows, G, C)), 4.0 * n)
    report(f"bn_apply relu+round rows{rows} G{G} C{C}", timeit(lambda: be.bn_apply(x, a, b, dx, rows, G, C, 1 | RND, False, H, W)), 8.0 * n)
    report(f"bn_bwd_reduce rows{rows} G{G} C{C}", timeit(lambda: be.bn_bwd_reduce(dy, x, a, b, mean, invstd, red, rows, G, C, 1, False, H, W)), 8.0 * n)
    report(f"bn_bwd_apply rows{rows} G{G} C{C}", timeit(lambda: be.bn_bwd_apply(dy, x, a, b, mean, invstd, os_, red, dx, None, None, False, rows, G, C, 1 | RND, False, H, W, True)), 12.0 * n)
    report(f"bn_bwd_apply +add rows{rows} G{G} C{C}", timeit(lambda: be.bn_bwd_apply(dy, x, a, b, mean, invstd, None, red, dx, None, None, False, rows, G, C, 1, False, H, W, True, dx_add=add)), 16.0 * n)
    if rows * 4 * G * C * 4 < 4e9:
        dyu = R(rows * G * 4, C)
        report(f"bn_bwd_reduce up2 rows{rows} G{G} C{C}", timeit(lambda: be.bn_bwd_reduce(dyu, x, a, b, mean, invstd, red, rows, G, C, 1, True, H, W)), 20.0 * n)
        report(f"bn_bwd_apply up2 rows{rows} G{G} C{C}", timeit(lambda: be.bn_bwd_apply(dyu, x, a, b, mean, invstd, os_, red, dx, None, None, False, rows, G, C, 1, True, H, W, True)), 24.0 * n)
        del dyu
    del x, dy, add, dx
