import time, numpy as np, torch, sys
sys.path.insert(0,'.')
import aim_b200 as A
dev=torch.device('cuda',0)
n=1<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=3):
    torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        t0=time.perf_counter(); f(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
print("H2D GB/s", n/t(lambda: d.copy_(h,non_blocking=True))/1e9)
print("D2H GB/s", n/t(lambda: h2.copy_(d2,non_blocking=True))/1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
print("bidir GB/s each", n/t(both)/1e9)
ms,rs=A.derive_knobs("wfa",150,0.04)
P=4_000_000
hp=[A.PinnedArray((P,),np.int32),A.PinnedArray((P,),np.int32),A.PinnedArray((P,rs),np.uint8),A.PinnedArray((P,rs),np.uint8)]
A.generate_pairs(4,P,150,0.04,rs,out=tuple(x.array for x in hp))
res=A.PinnedArray((P,),A.RESULT_DTYPE); ops=A.PinnedArray((P,2*rs),np.uint8)
par=A.AlignParams(algo="wfa",max_score=ms,read_size=rs,backtrace=True,reduce=True)
for it in range(3):
    t0=time.perf_counter(); r,o,ph=A.align_batch(par,*(x.array for x in hp),results=res.array,ops=ops.array); dt=time.perf_counter()-t0
    print("e2e %.1f ms  %.1fM pairs/s phases"%(dt*1e3,P/dt/1e6),[round(x,1) for x in ph])
