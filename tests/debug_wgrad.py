"""Debug probe for the tcgen05 wgrad kernel (not a test): structured inputs to decode operand-layout errors."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skillful_nowcasting_b200 import _lib
be = _lib.backend()
def run(n,d,h,w,cin,cout,kd,kh,kw,x,dz,algo):
    g = torch.full((kd*kh*kw*cout*cin,), float('nan'), device='cuda')
    be.conv_wgrad(x.cuda(), dz.cuda(), g, n,d,h,w,cin,cout,kd,kh,kw, algo=algo)
    torch.cuda.synchronize()
    return g.cpu().reshape(kd*kh*kw, cout, cin)
n,d,h,w,cin,cout = 4,1,32,32,32,64
for name,(kh,kw) in (('1x1',(1,1)),('3x3',(3,3))):
    x = torch.ones(n,d,h,w,cin); dz = torch.ones(n,d,h,w,cout)
    gs = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,1); gu = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,2)
    print(name,'ones: simt', gs[:,0,0].tolist(), 'umma', gu[:,0,0].tolist(), 'umma min/max', gu.min().item(), gu.max().item(), 'nonzero', (gu!=0).sum().item(), 'of', gu.numel())
    # one-hot probes
    for (p0,a,b) in [((0,0,5,7),3,2), ((1,0,20,9),40,31), ((3,0,31,31),63,0)]:
        x = torch.zeros(n,d,h,w,cin); dz = torch.zeros(n,d,h,w,cout)
        x[p0][b] = 1.0; dz[p0][a] = 2.0
        gu = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,2); gs = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,1)
        print(name,'onehot',p0,a,b,'simt nz', gs.nonzero().tolist()[:4], gs[gs!=0][:4].tolist(), '| umma nz', gu.nonzero().tolist()[:8], gu[gu!=0][:8].tolist())
    # channel ramp: dz[p][co] = co+1, x[p][ci] = 1 -> D[tap][co][ci] = (co+1)*count
    x = torch.ones(n,d,h,w,cin); dz = (torch.arange(cout).float()+1).expand(n,d,h,w,cout).contiguous()
    gu = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,2); gs = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,1)
    t = gu.shape[0]//2
    print(name,'co-ramp simt', (gs[t,:8,0]/gs[t,0,0]).tolist(), 'umma', gu[t,:8,0].tolist(), gu[t,:4,:4].tolist())
    x = (torch.arange(cin).float()+1).expand(n,d,h,w,cin).contiguous(); dz = torch.ones(n,d,h,w,cout)
    gu = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,2); gs = run(n,d,h,w,cin,cout,1,kh,kw,x,dz,1)
    print(name,'ci-ramp simt', (gs[t,0,:8]/gs[t,0,0]).tolist(), 'umma', gu[t,0,:8].tolist())
