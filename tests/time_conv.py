"""Time single conv shapes with CUDA events (tuning helper, not a test)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
shapes = [(288,1,128,128,96,96,18),(288,1,128,128,96,48,18),(288,1,64,64,192,192,18),(288,1,64,64,96,96,18),(288,1,32,32,384,384,18),(288,1,32,32,192,192,18)]
for (n,d,h,w,cin,cout,g) in shapes:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(9*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    res = torch.randn(n,d,h,w,cout,device="cuda"); y = torch.empty(n,d,h,w,cout,device="cuda")
    out=[]
    for algo in (2,3):
        for _ in range(2): be.conv_fwd(x,wp,bias,scale,res,y,n,d,h,w,cin,cout,1,3,3,g,0,algo=algo)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): be.conv_fwd(x,wp,bias,scale,res,y,n,d,h,w,cin,cout,1,3,3,g,0,algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        out.append(f"algo{algo}: {ms:.3f} ms {2*n*h*w*cin*cout*9/ms/1e9:.0f} TF/s")
    print(f"{n}x{h}x{w} {cin}->{cout}: "+"  ".join(out), "MT env", os.environ.get("DGMR_PATCH_MT"))
