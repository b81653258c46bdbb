"""Cost of the conv epilogue flags (tuning helper, not a test): the same launch with / without DGMR_FLAG_RES_UP2 / DGMR_FLAG_ROUND_OUT."""
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
for (n,d,h,w,cin,cout,kd,k,g) in [(288,1,128,128,96,48,1,3,18),(288,1,64,64,192,96,1,3,18),(32,22,64,64,8,48,3,3,1),(32,22,64,64,48,48,3,3,1),(256,1,64,64,48,48,1,3,8),(288,1,128,128,96,96,1,3,18)]:
    taps=kd*k*k
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(taps*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    res = torch.randn(n,d,h,w,cout,device="cuda"); res2 = torch.randn(n,d,h//2,w//2,cout,device="cuda"); y = torch.empty(n,d,h,w,cout,device="cuda")
    out=[]
    for name, r, act in (("plain+res",res,0),("res_up2",res2,2048),("relu",None,1),("relu+round_out",None,1|1024)):
        ms = timeit(lambda: be.conv_fwd(x,wp,bias,scale,r,y,n,d,h,w,cin,cout,kd,k,k,g,act))
        out.append(f"{name}: {ms:.3f} ms {2*n*d*h*w*cin*cout*taps/ms/1e9:.0f} TF/s")
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} k{kd}{k}{k}: "+"  ".join(out), flush=True)
