"""Time the small per-step ConvGRU convs under each kernel variant (tuning helper, not a test).
Two numbers per variant: `tp` = back-to-back independent launches (throughput), `dep` = a chain in which every launch consumes the previous
launch's output (the ConvGRU situation: latency).  Variants: v1 plain kernel, v1p persistent plain kernel, patch / patch1 halo-patch kernel
with / without CTA pairs, splitk tap-split accumulate mode."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
B = int(os.environ.get("GRU_B", "16"))
shapes = [(B,1,8,8,384,384),(B,1,16,16,192,192),(B,1,32,32,96,96),(B,1,64,64,48,48),(B,1,8,8,384,768),(B,1,16,16,192,384),(B,1,32,32,96,192),(B,1,64,64,48,96)]
def timeit(f, n=20):
    for _ in range(3): f(0)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
for (n,d,h,w,cin,cout) in shapes:
    x = torch.randn(n,d,h,w,cin,device="cuda")*0.1; wp = torch.randn(9*cout*cin,device="cuda")/ (9*cin)**0.5
    scale = torch.rand(1,cout,device="cuda")*0.1+0.9
    res = torch.randn(n,d,h,w,cout,device="cuda"); y = torch.empty(n,d,h,w,cout,device="cuda")
    bufs = [torch.randn(n,d,h,w,cin,device="cuda")*0.1, torch.empty(n,d,h,w,cin,device="cuda")]
    out=[]
    for name, algo, acc, opts in (("v1",2,0,{"umma_persist":0}),("v1p",2,0,{"umma_persist":2}),("patch",3,0,{}),("patch1",3,0,{"patch_pair":0}),("splitk",2,512,{})):
        for k_,v_ in (("umma_persist",-1),("patch_pair",-1)): be.set_option(k_, v_)
        for k_,v_ in opts.items(): be.set_option(k_, v_)
        try:
            if acc:
                f = lambda i: be.conv_fwd(x,wp,None,scale,None,y,n,d,h,w,cin,cout,1,3,3,1,acc,algo=algo)
            else:
                f = lambda i: be.conv_fwd(x,wp,None,scale,res,y,n,d,h,w,cin,cout,1,3,3,1,0,algo=algo)
            us = timeit(f)
            s = f"{name}: tp {us:.1f}"
            if cin == cout and not acc:     # dependent chain: out of launch i is the input of launch i+1
                g = lambda i: be.conv_fwd(bufs[i&1],wp,None,scale,None,bufs[1-(i&1)],n,d,h,w,cin,cout,1,3,3,1,0,algo=algo)
                s += f" dep {timeit(g):.1f}"
            out.append(s + " us")
        except Exception as e:
            out.append(f"{name}: n/a")
    print(f"{n}x{h}x{w} {cin}->{cout} ({2*n*h*w*cin*cout*9/1e9:.1f} GF): "+"  ".join(out), flush=True)
