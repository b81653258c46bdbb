"""Pair-persistent whole-row kernel (algo 5) vs halo-patch (algo 3) vs plain (algo 2) on the larger sampler / discriminator layers (tuning helper)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
def timeit(f, n=5):
    for _ in range(2): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for (n,d,h,w,cin,cout,kd,g) in [(16,1,16,16,192,192,1,1),(16,1,16,16,384,192,1,1),(288,1,64,64,96,96,1,18),(288,1,64,64,192,96,1,18),(288,1,128,128,96,96,1,18),(32,11,32,32,96,96,3,1),(288,1,128,128,48,96,1,1),
                                (256,1,32,32,96,96,1,8),(256,1,64,64,48,48,1,8),(16,1,32,32,96,192,1,1),(16,1,16,16,192,384,1,1),(16,1,64,64,48,96,1,1),(160,1,16,16,192,192,1,5)]:
    taps=kd*9
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(taps*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    y = torch.empty(n,d,h,w,cout,device="cuda")
    out=[]
    for name, algo in (("auto",0),("patch",3),("rows-pair",5)):
        try:
            ms = timeit(lambda: be.conv_fwd(x,wp,bias,scale,None,y,n,d,h,w,cin,cout,kd,3,3,g,1,algo=algo))
            out.append(f"{name}: {ms:.3f} ms {2*n*d*h*w*cin*cout*taps/ms/1e9:.0f} TF/s")
        except Exception as e:
            out.append(f"{name}: n/a")
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} k{kd}33 G{g}: "+"  ".join(out), flush=True)
