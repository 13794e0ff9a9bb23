"""Shape sweep: the shared-memory-resident gather (edgeconv_smem.cu) against the global-memory gather (edgeconv.cu),
train and eval entry points, bit-exact sel / arg / out. Run on a GPU box: python tools/check_gather_variants.py"""
import os, sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, _lib
dev='cuda'
torch.manual_seed(0)
for (B,N,k,Cp) in [(2,256,8,64),(2,256,20,64),(32,256,8,64),(2,2048,20,64),(3,1000,20,64),(2,4096,40,64),(1,8192,40,64),(3,700,7,128),(2,300,33,256),(5,1500,20,64)]:
    P=B*N
    idx=torch.stack([torch.stack([torch.randperm(N,device=dev)[:k] for _ in range(N)]) for _ in range(B)]).int().contiguous()
    g=ops.KnnGraph(idx); rev_ptr,_=g.reverse()
    table=torch.randn(P,2*Cp,device=dev); gamma=torch.randn(Cp,device=dev); coef=torch.randn(4*Cp,device=dev)
    res={}
    for mode in ('global','smem'):
        os.environ['FS_GATHER']=mode
        sel=torch.empty(P,Cp,device=dev); arg=torch.empty(P,Cp,dtype=torch.uint8,device=dev); sy=torch.empty(P,Cp,device=dev); st=ops._stats_buffer(Cp,dev)
        _lib.call("fs_edgeconv_gather", table, table, 0, table.stride(0), idx, B, N, k, Cp, gamma, rev_ptr, sel, arg, sy, st)
        out=torch.empty(P,Cp,device=dev); outb=torch.empty(P,Cp,device=dev,dtype=torch.bfloat16); arg2=torch.empty_like(arg)
        _lib.call("fs_edgeconv_fused_eval", table, table, 0, table.stride(0), idx, B, N, k, Cp, coef, out, 0, out.stride(0), arg2)
        _lib.call("fs_edgeconv_fused_eval", table, table, 0, table.stride(0), idx, B, N, k, Cp, coef, outb, 1, outb.stride(0), None)
        torch.cuda.synchronize()
        res[mode]=(sel,arg,sy,st[:3*Cp].clone(),out,arg2,outb)
    a,b=res['global'],res['smem']
    print((B,N,k,Cp),'sel',torch.equal(a[0],b[0]),'arg',torch.equal(a[1],b[1]),'sy',float((a[2]-b[2]).abs().max()),
          'stats',float(((a[3]-b[3]).abs()/(a[3].abs()+1e-6)).max()),'eval',torch.equal(a[4],b[4]),torch.equal(a[5],b[5]),torch.equal(a[6],b[6]))
