"""Column-stacked kernel (conv_kwstack.cu) vs the plain tcgen05 kernel on the benchmark's narrow-output layers (tuning helper, not a test)."""
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
for (n,d,h,w,cin,cout,kd,g) in [(32,22,64,64,48,48,3,1),(288,1,128,128,96,48,1,18),(256,1,64,64,48,48,1,8),(288,1,64,64,96,48,1,18),(32,11,32,32,96,48,3,1),
                                (32,22,64,64,16,48,1,1),(32,22,64,64,48,16,1,1),(16,1,64,64,48,48,1,1),(16,1,64,64,96,48,1,1)]:
    taps=kd*9
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(taps*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    y = torch.empty(n,d,h,w,cout,device="cuda"); y2 = torch.empty_like(y)
    out=[]
    for name, algo, pair in (("plain",2,0),("kwstack",4,0),("kwstack-pair",4,-1)):
        be.set_option("kwstack_pair", pair)
        ms = timeit(lambda: be.conv_fwd(x,wp,bias,scale,None,y if algo==2 else y2,n,d,h,w,cin,cout,kd,3,3,g,1,algo=algo))
        out.append(f"{name}: {ms:.3f} ms {2*n*d*h*w*cin*cout*taps/ms/1e9:.0f} TF/s")
    err = (y-y2).abs().max().item()/y.abs().max().item()
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} k{kd}33 G{g}: "+"  ".join(out)+f"  rel diff {err:.1e}", flush=True)
