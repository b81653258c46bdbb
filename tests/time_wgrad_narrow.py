"""Row wgrad kernel on narrow images (W = 16 / 8: a K block is 2 / 4 whole image rows) vs the tap-wise kernel (tuning helper, not a test)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
shapes = [(288,16,16,768,768),(288,16,16,384,384),(288,16,16,768,384),(288,16,16,384,192),(256,16,16,192,192),(288,8,8,768,768),(288,8,8,384,768),(256,8,8,384,384),(288,32,32,192,192)]
for (n,h,w,cin,cout) in shapes:
    x = torch.randn(n,1,h,w,cin,device="cuda"); dz = torch.randn(n,1,h,w,cout,device="cuda")
    dwp = torch.empty(9*cout*cin,device="cuda"); ref = torch.empty_like(dwp)
    out=[]
    for algo in (2,3):
        tgt = ref if algo == 2 else dwp
        for _ in range(2): be.conv_wgrad(x,dz,tgt,n,1,h,w,cin,cout,1,3,3,algo=algo)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): be.conv_wgrad(x,dz,tgt,n,1,h,w,cin,cout,1,3,3,algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        out.append(f"{'tap' if algo==2 else 'row'}: {ms:.3f} ms {2*n*h*w*cin*cout*9/ms/1e9:.0f} TF/s")
    print(f"{n}x{h}x{w} {cin}->{cout}: "+"  ".join(out)+f"  rel diff {(ref-dwp).abs().max().item()/ref.abs().max().item():.1e}", flush=True)
