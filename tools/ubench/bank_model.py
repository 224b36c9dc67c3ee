import numpy as np
rng=np.random.default_rng(0)
roll=np.deg2rad(3.0); cd=0.492
def lane_xy(sub=(8,4)):
    i=np.arange(32)%sub[0]; j=np.arange(32)//sub[0]
    d1=i*1.0; d2=j*1.0
    dx=np.cos(roll)*d1+np.sin(roll)*d2; dy=-np.sin(roll)*d1+np.cos(roll)*d2
    return -dx/cd, -dy/cd
def wavefronts32(addr_words):
    # 32-bit loads: max multiplicity of distinct addresses per bank
    banks=addr_words%32
    w=0
    for b in range(32):
        w=max(w,len(set(addr_words[banks==b])))
    return w
def wavefronts64(addr_words):  # 64-bit loads: per half warp, 8B units; bank pairs
    tot=0
    for h in range(2):
        a=addr_words[h*16:(h+1)*16]
        banks=(a//2)%16   # 16 "double-banks"
        w=0
        for b in range(16):
            w=max(w,len(set(a[banks==b])))
        tot+=w
    return tot
for sub in ((8,4),(16,2),(4,8),(32,1)):
    lx,ly=lane_xy(sub)
    print("sub-patch",sub)
    for pitch in (192,196,200,204,208,212,216,220,224,228,232):
        s32=[];s64=[]
        for t in range(300):
            fx,fy=rng.random(2)*4
            x=np.floor(lx+fx+20).astype(int); y=np.floor(ly+fy+20).astype(int)
            w=0
            for r in range(3):
                for c in range(3):
                    w+=wavefronts32((y-1+r)*pitch+(x-1+c))
            s32.append(w)
            w=0
            for r in range(3):
                a0=((x-1)&~1)
                w+=wavefronts64((y-1+r)*pitch+a0)+wavefronts64((y-1+r)*pitch+a0+2)
            s64.append(w)
        print(f"  pitch {pitch} (mod32 {pitch%32:2d}): 9xLDS.32 = {np.mean(s32):5.1f} wavefronts   6xLDS.64 = {np.mean(s64):5.1f}")
