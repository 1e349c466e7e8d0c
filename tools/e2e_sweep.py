import time, numpy as np, sys, os
sys.path.insert(0,'.')
import aim_b200 as A
ms,rs=A.derive_knobs("wfa",150,0.04)
P=4_000_000
hp=[A.PinnedArray((P,),np.int32),A.PinnedArray((P,),np.int32),A.PinnedArray((P,rs),np.uint8),A.PinnedArray((P,rs),np.uint8)]
A.generate_pairs(4,P,150,0.04,rs,out=tuple(x.array for x in hp))
res=A.PinnedArray((P,),A.RESULT_DTYPE); ops=A.PinnedArray((P,2*rs),np.uint8)
par=A.AlignParams(algo="wfa",max_score=ms,read_size=rs,backtrace=True,reduce=True)
for mb in (16,32,64,96,192,384,768):
    os.environ["AIM_CHUNK_MB"]=str(mb)
    best=1e9
    for it in range(4):
        t0=time.perf_counter(); r,o,ph=A.align_batch(par,*(x.array for x in hp),results=res.array,ops=ops.array); dt=time.perf_counter()-t0
        if it: best=min(best,dt)
    print("chunk %4d MB  e2e %.1f ms  %.1fM pairs/s phases"%(mb,best*1e3,P/best/1e6),[round(x,1) for x in ph], flush=True)
