"""Sub-pixel forward: halo-patch form vs whole-row CTA-pair form (tuning helper, not a test)."""
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
for (n,h,w,cin,cout,g) in [(288,16,16,384,384,18),(288,32,32,192,192,18),(288,8,8,768,768,18)]:
    x = torch.randn(n,1,h,w,cin,device="cuda"); wsp = torch.randn(16*cout*cin,device="cuda")/30
    bias = torch.randn(cout,device="cuda"); scale = torch.rand(g,cout,device="cuda")+0.5
    y = torch.empty(n,1,2*h,2*w,cout,device="cuda")
    out=[]
    for name, rows in (("patch",0),("rows-pair",1)):
        be.set_option("subpix_rows", rows)
        try:
            ms = timeit(lambda: be.upconv_fwd(x,wsp,bias,scale,None,y,n,h,w,cin,cout,g,1))
            out.append(f"{name}: {ms:.3f} ms {2*n*h*w*cin*cout*16/ms/1e9:.0f} TF/s")
        except Exception as e:
            out.append(f"{name}: n/a ({str(e)[:60]})")
    print(f"{n}x{h}x{w} {cin}->{cout} G{g}: "+"  ".join(out), flush=True)
