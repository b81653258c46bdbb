"""Probe (not a test; to be run next round): how many mantissa bits does one tcgen05 kind::tf32 MMA chain keep while accumulating?

dgmr_debug_umma_shift computes C[128][N] = A[0:128, 0:32] . B[N, 0:32]^T in ONE accumulation chain of 4 MMAs (K = 8 each).
All operands below are exactly representable in TF32 and every product is exact in fp32, so any deviation from the fp64 result is
the accumulator / adder-tree precision, not operand rounding:
  row r of A = [1, 2^-s, 2^-s, ...]  with B = ones  ->  exact sum 1 + 31 * 2^-s; sweeping s shows where small addends are dropped
  (fp32 round-to-nearest keeps them down to s = 23/24).  A second sweep uses alternating signs to tell truncation from rounding.
Motivation: the full-size adjoint identity <conv(x),dy> = <x,dgrad(dy)> held only to 1.0e-4 on the B200 while fp32 CPU arithmetic
holds it to 1e-10 (tests/fullsize_props_draft.py).
"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
N = 16
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(A, B):
    C = torch.full((128, N), float("nan"), device="cuda")
    rc = be.lib.dgmr_debug_umma_shift(A.data_ptr(), B.data_ptr(), C.data_ptr(), N, 0, 0, st)
    assert rc == 0
    torch.cuda.synchronize()
    return C


for sign in ("same", "alternating"):
    print("addend signs:", sign)
    for s in range(8, 26):
        A = torch.zeros(256, 32, device="cuda")
        A[:, 0] = 1.0
        A[:, 1:] = 2.0 ** -s
        if sign == "alternating":
            A[:, 2::2] *= -1.0
        B = torch.ones(N, 32, device="cuda")
        C = run(A, B)
        exact = (A[:128].double() @ B.double().t())
        got = C.double()
        print(f"  s={s:2d}  exact-1 = {exact[0, 0].item() - 1:.6e}   got-1 = {got[0, 0].item() - 1:.6e}   rel err {((got - exact).abs().max() / exact.abs().max()).item():.2e}")
