"""Time the small per-step ConvGRU convs under each algorithm (tuning helper, not a test)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
shapes = [(16,1,8,8,384,384),(16,1,16,16,192,192),(16,1,32,32,96,96),(16,1,64,64,48,48),(16,1,8,8,384,768),(16,1,16,16,192,384),(16,1,32,32,96,192),(16,1,64,64,48,96)]
def timeit(f):
    for _ in range(3): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/20*1000
for (n,d,h,w,cin,cout) in shapes:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(9*cout*cin,device="cuda")/30
    scale = torch.rand(1,cout,device="cuda")+0.5
    res = torch.randn(n,d,h,w,cout,device="cuda"); y = torch.empty(n,d,h,w,cout,device="cuda")
    out=[]
    for name, algo, acc in (("v1",2,0),("patch",3,0),("splitk",2,512)):
        try:
            if acc:
                f = lambda: be.conv_fwd(x,wp,None,scale,None,y,n,d,h,w,cin,cout,1,3,3,1,acc,algo=algo)
            else:
                f = lambda: be.conv_fwd(x,wp,None,scale,res,y,n,d,h,w,cin,cout,1,3,3,1,0,algo=algo)
            us = timeit(f)
            out.append(f"{name}: {us:.1f} us {2*n*h*w*cin*cout*9/us/1e6:.0f} TF/s")
        except Exception as e:
            out.append(f"{name}: n/a")
    print(f"{n}x{h}x{w} {cin}->{cout}: "+"  ".join(out))
