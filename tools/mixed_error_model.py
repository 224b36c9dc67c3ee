"""NumPy emulation of the mixed-arithmetic kernel's sample arithmetic (csrc/coreg_lag_roll.cu, roll_segment_mixed):
how far its samples are from the reference's float32-stored samples, for images whose mean is far above their
contrast. Calibrates the per-lag guard `mixed_guard_trips`. CPU only.

    python tools/mixed_error_model.py        ->  profiles/r2_mixed_error_model.md quotes the output

For each mean level (sigma = 1): the reference sample S (FP64 spline of the float32 image) and its float32 store; the
old scheme (FP32 spline on the raw values) and the new one (FP32 spline on the image centred on its float32 pivot,
fractions quantised to 2^-23 like the kernel's, b = round32(t + p), bc = b - p).
"""
import numpy as np
rng=np.random.default_rng(1)
from scipy.ndimage import gaussian_filter
n=600
img=gaussian_filter(rng.standard_normal((n,n)),2.0); img=(img-img.mean())/img.std()
for mean in (0.0, 500/300, 3e4, 1e6):
    v32=(img+mean).astype(np.float32)
    p32=np.float32(v32.astype(np.float64).mean())
    c32=(v32-p32).astype(np.float32)
    # random sample points
    m=400000
    iy=rng.integers(1,n-2,m); ix=rng.integers(1,n-2,m)
    vx=rng.random(m); vy=rng.random(m)   # v = d+0.5 in [0,1)
    def spline(a, vx, vy, dt):
        a=a.astype(dt); vx=vx.astype(dt); vy=vy.astype(dt)
        half=dt(0.5)
        def rowq(r):
            ta=a[iy-1+r, ix-1]; tb=a[iy-1+r, ix]; tc=a[iy-1+r, ix+1]
            ca=half*(ta+tb); cb=tb-ta; cc=half*(ta+tc)-tb
            return (cc*vx+cb)*vx+ca
        q0,q1,q2=rowq(0),rowq(1),rowq(2)
        return ((half*(q0+q2)-q1)*vy+(q1-q0))*vy+half*(q0+q1)
    S=spline(v32.astype(np.float64),vx,vy,np.float64)          # reference sample in f64
    ref=S.astype(np.float32).astype(np.float64)                # reference's float32 store
    # mixed new: centred
    vxq=(np.round(vx*2**23)/2**23); vyq=(np.round(vy*2**23)/2**23)
    t=spline(c32,vxq,vyq,np.float32)
    b=((t+p32).astype(np.float32))
    bc=(b-p32).astype(np.float32).astype(np.float64)
    refc=ref-np.float64(p32)
    e=bc-refc
    # old mixed: spline on raw values in f32
    t_old=spline(v32,vxq,vyq,np.float32).astype(np.float64)
    e_old=t_old-ref
    sig=refc.std()
    tc=(t.astype(np.float64)-(S-np.float64(p32)))
    print(f"mean {mean:9.1f}: sigma_b {sig:.3f} | centred-spline err rms/sigma {tc.std()/sig:.2e} ({tc.std()/(2**-24*np.sqrt((c32.astype(np.float64)**2).mean())):.2f} x 2^-24 rms) | new sample err rms/sigma {e.std()/sig:.2e} frac!=0 {np.mean(e!=0):.2e} mean {e.mean()/sig:.1e} | old err rms/sigma {e_old.std()/sig:.2e}")
    # effect on r against a partner image a = S + noise
    a=refc+0.3*rng.standard_normal(m)
    def r(x,y): return np.corrcoef(x,y)[0,1]
    print("      dr new %.2e  dr old %.2e" % (abs(r(a,bc)-r(a,refc)), abs(r(a,t_old-np.float64(p32))-r(a,refc))))
