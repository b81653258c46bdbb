"""Debug probe (not a test): isolate which tensor-core-path option breaks a block's gradients."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from block_cases import block_cases, run_block_case
from skillful_nowcasting_b200 import ops
cases = {c[0]: c for c in block_cases(True)}
orig = ops._round_
def kernel_only(t):
    ops._be().round_tf32(t)
    return t
def attr_only(t):
    t._dgmr_tf32 = True
    return t
def clone_round(t):
    return orig(t)
for label, fn in (("orig", orig), ("kernel_only(no attr)", kernel_only), ("attr_only(no kernel)", attr_only)):
    ops._round_ = fn
    ops.config.round_tf32 = True
    ops.config._dbg_round_act, ops.config._dbg_round_w, ops.config._dbg_round_dz = True, False, False
    ops.clear_pack_cache()
    try:
        run_block_case(cases["g_proj"], True, "cuda", 1e-3, 5e-2, tol_buf=1e-3)
        res = "ok"
    except AssertionError as e:
        res = "FAIL " + str(e)[:120]
    print(f"g_proj {label}: {res}")
