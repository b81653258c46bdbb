import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
def timeit(f):
    for _ in range(2): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/5
for (n,d,h,w,cin,cout,kd) in [(32,22,64,64,48,48,3),(32,22,64,64,48,48,1),(288,1,128,128,96,48,1),(32,22,64,64,64,64,3)]:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(kd*9*cout*cin,device="cuda")/30
    y = torch.empty(n,d,h,w,cout,device="cuda")
    ms = timeit(lambda: be.conv_fwd(x,wp,None,None,None,y,n,d,h,w,cin,cout,kd,3,3,1,0,algo=3))
    print(os.environ.get("DGMR_PATCH_DBG"), os.environ.get("DGMR_PATCH_MT"), f"{n}x{d}x{h}x{w} {cin}->{cout} kd{kd}: {ms:.3f} ms {2*n*d*h*w*cin*cout*9*kd/ms/1e9:.0f} TF/s")
