"""Plain vs halo-patch kernel on the weight-heavy 16x16 / 8x8 sampler layers (tuning helper, not a test)."""
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
for (n,h,w,cin,cout,g) in [(288,16,16,768,768,18),(288,16,16,384,384,18),(288,16,16,768,384,18),(288,16,16,384,768,1),(288,32,32,192,192,18),(288,32,32,384,192,18),(288,16,16,384,192,18),(288,16,16,192,384,1),(256,16,16,192,192,8)]:
    x = torch.randn(n,1,h,w,cin,device="cuda"); wp = torch.randn(9*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    y = torch.empty(n,1,h,w,cout,device="cuda")
    out=[]
    for name, algo, mt in (("plain",2,-1),("patch mt1",3,1),("rows",5,0),("rows-pair",5,-1)):
        be.set_option("patch_mt", mt if algo == 3 else -1)
        if algo == 5: be.set_option("kwstack_pair", mt)
        try:
            ms = timeit(lambda: be.conv_fwd(x,wp,bias,scale,None,y,n,1,h,w,cin,cout,1,3,3,g,1,algo=algo))
            out.append(f"{name}: {ms:.3f} ms {2*n*h*w*cin*cout*9/ms/1e9:.0f} TF/s")
        except Exception as e:
            out.append(f"{name}: n/a")
    print(f"{n}x{h}x{w} {cin}->{cout} G{g}: "+"  ".join(out), flush=True)
