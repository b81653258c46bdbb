"""Time 1x1 conv shapes of the step (tuning helper, not a test)."""
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
for (n,d,h,w,cin,cout,g) in [(288,1,128,128,48,4,18),(32,22,64,64,8,48,1),(288,1,64,64,48,96,18),(288,1,32,32,96,192,18),(288,1,64,64,96,48,18),(288,1,16,16,192,384,18),(256,1,64,64,8,48,8),(288,1,32,32,192,96,18)]:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(cout*cin,device="cuda")/8
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    y = torch.empty(n,d,h,w,cout,device="cuda")
    ms = timeit(lambda: be.conv_fwd(x,wp,bias,scale,None,y,n,d,h,w,cin,cout,1,1,1,g,0,algo=2))
    gb = (x.numel()+y.numel())*4/1e9
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} k111: {ms:.3f} ms  {gb/ms:.2f} TB/s")
