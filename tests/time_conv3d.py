"""Time the temporal-discriminator and small-channel conv shapes per algorithm (tuning helper, not a test)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
shapes = [(32,22,64,64,48,48,3),(32,22,64,64,8,48,3),(32,22,64,64,48,8,3),(32,11,32,32,96,96,3),(32,11,32,32,48,96,3),(32,11,32,32,96,48,3),
          (288,1,128,128,96,48,1),(288,1,64,64,96,48,1),(288,1,128,128,48,48,1),(128,1,128,128,8,48,1),(128,1,128,128,48,48,1),(128,1,64,64,48,96,1)]
def timeit(f):
    for _ in range(2): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/5
for (n,d,h,w,cin,cout,kd) in shapes:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(kd*9*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda")
    y = torch.empty(n,d,h,w,cout,device="cuda")
    out=[]
    for name, algo in (("v1",2),("patch",3)):
        try:
            ms = timeit(lambda: be.conv_fwd(x,wp,bias,None,None,y,n,d,h,w,cin,cout,kd,3,3,1,1,algo=algo))
            out.append(f"{name}: {ms:.3f} ms {2*n*d*h*w*cin*cout*9*kd/ms/1e9:.0f} TF/s")
        except Exception as e:
            out.append(f"{name}: n/a")
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} kd{kd}: "+"  ".join(out))
