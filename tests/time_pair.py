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
tag = " ".join(f"{k[11:]}={os.environ[k]}" for k in ("DGMR_PATCH_PAIR","DGMR_PATCH_DBG","DGMR_PATCH_MT") if k in os.environ)
for (n,d,h,w,cin,cout) in [(288,1,128,128,96,96),(288,1,64,64,192,192),(288,1,32,32,384,384)]:
    x = torch.randn(n,d,h,w,cin,device="cuda"); wp = torch.randn(9*cout*cin,device="cuda")/30
    res = torch.randn(n,d,h,w,cout,device="cuda"); y = torch.empty(n,d,h,w,cout,device="cuda")
    ms = timeit(lambda: be.conv_fwd(x,wp,None,None,res,y,n,d,h,w,cin,cout,1,3,3,1,0,algo=3))
    print(f"[{tag}] {n}x{h}x{w} {cin}->{cout}: {ms:.3f} ms {2*n*h*w*cin*cout*9/ms/1e9:.0f} TF/s")
