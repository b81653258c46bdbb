import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
blocks = 148
for N in (48, 96, 192, 256):
    for shift in (0, 1, 3, 4, 8, 9):
        row = []
        for mode in (0, 1):
            out = torch.zeros(blocks, device="cuda")
            rc = be.lib.dgmr_debug_umma_rate(ctypes.c_void_p(out.data_ptr()), blocks, N, 4096, mode, shift, st)
            assert rc == 0
            torch.cuda.synchronize()
            row.append(f"mode{mode}: {out.mean().item():.1f}")
        print(f"N {N} shift {shift}: " + "  ".join(row) + f"   (floor {128*N/256:.0f} cyc)")
