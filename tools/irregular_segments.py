"""Model of how often the rolling kernel meets an irregular column segment on BASELINE configs[3] (5-D lag grid):
fraction of the segments in which a floor moves out of step, and how those segments are spread over the warps (32
neighbouring columns of one row group). CPU only; similarity-transform approximation of the per-lag homography.

    python tools/irregular_segments.py [--rows 12]
"""
import numpy as np
rng=np.random.default_rng(0)
import sys
P = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 12
n=2048
cd=0.492
res=[]
for trial in range(60):
    th=np.deg2rad(rng.choice((np.arange(10)-5)*0.1))
    d1=rng.choice((np.arange(16)-8)*0.001)/cd
    d2=rng.choice((np.arange(16)-8)*0.001)/cd
    sx=rng.integers(14,34)/cd; sy=rng.integers(-4,16)/cd
    i=np.arange(n)[None,:]; j=np.arange(n)[:,None]
    c=1023.5
    x=c+((i-c)*np.cos(th)+(j-c)*np.sin(th))/(1+d1)+sx+0.5
    y=c+(-(i-c)*np.sin(th)+(j-c)*np.cos(th))/(1+d2)+sy+0.5
    # segments: rows grouped by P
    ns=n//P
    xs=x[:ns*P].reshape(ns,P,n); ys=y[:ns*P].reshape(ns,P,n)
    fx=np.floor(xs[:,0]); fy=np.floor(ys[:,0])
    vx=xs-fx[:,None]; vy=ys-(fy[:,None]+np.arange(P)[None,:,None])
    irr=((vx<0)|(vx>=1)|(vy<0)|(vy>=1)).any(axis=1)   # ns x n
    w=irr.reshape(ns,n//32,32).sum(axis=2)
    res.append((np.rad2deg(th),d1,d2,irr.mean(),(w==0).mean(),((w>0)&(w<=10)).mean(),(w>10).mean(), w[(w>0)&(w<=10)].mean() if ((w>0)&(w<=10)).any() else 0))
r=np.array(res)
print("mean irregular frac %.3f; warps none %.3f coop %.3f serial %.3f; mean needy in coop warps %.2f"%tuple(r[:,3:].mean(axis=0)))
for row in r[:12]: print(" ".join("%.4f"%v for v in row))
