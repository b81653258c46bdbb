"""Probe: may a tcgen05 A descriptor start at an arbitrary 128-byte row of a SWIZZLE_128B tile? (not a test)"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
torch.manual_seed(0)
N = 64
A = torch.randn(256, 32, device="cuda"); B = torch.randn(N, 32, device="cuda")
for mode in (0, 1):
    row = []
    for r0 in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 33, 65, 127, 128):
        C = torch.full((128, N), float("nan"), device="cuda")
        rc = be.lib.dgmr_debug_umma_shift(A.data_ptr(), B.data_ptr(), C.data_ptr(), N, r0, mode, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = A[r0:r0 + 128].double() @ B.double().t()
        err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
        row.append(f"r0={r0}:{err:.1e}")
    print("mode", mode, " ".join(row))
