"""One launch of each headline tensor-core kernel at benchmark size, for `ncu --set full` (profiling helper, not a test):
    ncu --set full --clock-control none --import-source on -k regex:'conv_umma|conv_subpix' -c 12 -o gpurun_out/prof_r02 python tests/prof_kernels.py
Order of the profiled launches (after one untimed warm-up pass that ncu skips with -s): see LAUNCHES below."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from prof_kernels_util import conv, upconv, wgrad   # noqa: E402

up_f, up_d = upconv(288, 64, 64, 96, 96, 18)
LAUNCHES = [
    ("patch pair 96->96 @128^2 (g4 / up_g4.last... GBlock convs)", conv(288, 1, 128, 128, 96, 96, 1, 3, 18)),
    ("plain 96->48 @128^2 (up_g4.last_conv_3x3: the kernel VERDICT r1 names)", conv(288, 1, 128, 128, 96, 48, 1, 3, 18)),
    ("sub-pixel forward 96->96 64^2 -> 128^2 (up_g4.first_conv_3x3)", up_f),
    ("sub-pixel dgrad 96->96", up_d),
    ("temporal D 3-D 48->48 k333", conv(32, 22, 64, 64, 48, 48, 3, 3, 1, res=False)),
    ("ConvGRU level-4 step conv 48->96 @64^2, 16 images", conv(16, 1, 64, 64, 48, 96, 1, 3, 1)),
    ("row wgrad 96->96 @128^2", wgrad(288, 1, 128, 128, 96, 96, 1, 3)),
]
for rep in range(2):          # pass 0 = warm-up (attributes, tensor maps), pass 1 = the profiled launches
    for name, f in LAUNCHES:
        f()
    torch.cuda.synchronize()
for name, _ in LAUNCHES:
    print(name)
