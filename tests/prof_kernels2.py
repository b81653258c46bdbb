"""Second profiling set (round 2, after the column-stacked / narrow-row kernels), same conventions as prof_kernels.py:
    ncu --set full --clock-control none --import-source on -k regex:'conv_umma|conv_subpix' -s 5 -c 5 -o gpurun_out/prof_r02b python tests/prof_kernels2.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prof_kernels_util import conv, upconv, wgrad   # noqa: E402

up_f, _ = upconv(288, 32, 32, 192, 192, 18)
LAUNCHES = [
    ("patch pair 96->96 @64^2 G18 (9 calls / step)", conv(288, 1, 64, 64, 96, 96, 1, 3, 18)),
    ("patch pair 192->192 @32^2 G18", conv(288, 1, 32, 32, 192, 192, 1, 3, 18)),
    ("sub-pixel forward 192->192 32^2 -> 64^2", up_f),
    ("column-stacked pair, temporal D 3-D 48->48 k333", conv(32, 22, 64, 64, 48, 48, 3, 3, 1, res=False)),
    ("row wgrad 768->768 @16^2 (two image rows per K block)", wgrad(288, 1, 16, 16, 768, 768, 1, 3)),
]
for rep in range(2):          # pass 0 = warm-up (attributes, tensor maps), pass 1 = the profiled launches
    for name, f in LAUNCHES:
        f()
    torch.cuda.synchronize()
for name, _ in LAUNCHES:
    print(name)
