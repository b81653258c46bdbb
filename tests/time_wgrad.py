"""Time single wgrad shapes with CUDA events (tuning helper, not a test)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
shapes = [(288,1,128,128,96,96,3),(288,1,128,128,96,48,3),(288,1,64,64,192,192,3),(288,1,64,64,96,96,3),(288,1,32,32,384,384,3),(288,1,32,32,192,192,3),
          (32,22,64,64,48,48,3),(128,1,128,128,8,48,3),(288,1,64,64,192,96,1)]
for (n,d,h,w,cin,cout,k) in shapes:
    kd = 3 if d > 1 else 1
    x = torch.randn(n,d,h,w,cin,device="cuda"); dz = torch.randn(n,d,h,w,cout,device="cuda")
    dwp = torch.empty(kd*k*k*cout*cin,device="cuda")
    out=[]
    for algo in (0,):
        for _ in range(2): be.conv_wgrad(x,dz,dwp,n,d,h,w,cin,cout,kd,k,k,algo=algo)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): be.conv_wgrad(x,dz,dwp,n,d,h,w,cin,cout,kd,k,k,algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        out.append(f"algo{algo}: {ms:.3f} ms {2*n*d*h*w*cin*cout*kd*k*k/ms/1e9:.0f} TF/s")
    print(f"{n}x{d}x{h}x{w} {cin}->{cout} k{kd}{k}{k}: "+"  ".join(out))
