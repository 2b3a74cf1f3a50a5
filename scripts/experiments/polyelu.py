import numpy as np
f16=np.float16
def h(x): return np.asarray(x,dtype=np.float32).astype(f16)
def fma16(a,b,c): return h(a.astype(np.float32)*b.astype(np.float32)+c.astype(np.float32))  # single rounding approx (f32 exact product of 11-bit; sum rounding in f32 then f16)
def add16(a,b): return h(a.astype(np.float32)+b.astype(np.float32))
def poly_elu(z16, c, coef):
    # z16: float16 array of z' values (<0 relevant)
    M=h(1536.0)
    zc=np.maximum(z16, h(-13.0))
    u=add16(M,-zc)
    kf=add16(u,-M)
    r=add16(zc,kf)
    a3,a2,a1,a0=[h(v) for v in coef]
    p=fma16(np.broadcast_to(a3,r.shape),r,np.broadcast_to(a2,r.shape))
    p=fma16(p,r,np.broadcast_to(a1,r.shape))
    p=fma16(p,r,np.broadcast_to(a0,r.shape))
    km=(u.view(np.uint16)&0xF).astype(np.uint16)
    res=(p.view(np.uint16)-(km<<10)).astype(np.uint16).view(f16)
    f=fma16(res,np.broadcast_to(h(c/2),r.shape),np.broadcast_to(h(-c),r.shape))
    return np.where(z16<0,f,z16)
c=1.4426950408889634
# all negative fp16 values
allh=np.arange(0x8001,0xFC00,dtype=np.uint16).view(f16)   # -tiny .. -65504
z=allh
ref=c*(np.exp2(z.astype(np.float64))-1)
def err(coef):
    out=poly_elu(z,c,coef).astype(np.float64)
    e=np.abs(out-ref)
    return e.max(), z[e.argmax()]
base=[0.1103433,0.4852223,1.3865219,1.9998561]
print('base',err(base))
# mufu-like reference: exact 2^z rounded to f16 then fma
e16=h(np.exp2(z.astype(np.float64)))
fm=fma16(e16,np.broadcast_to(h(c),z.shape),np.broadcast_to(h(-c),z.shape)).astype(np.float64)
print('mufu-ideal',np.abs(fm-ref).max())
# local search over fp16-representable coefficient neighbours
import itertools
best=(err(base)[0],base)
cands=[]
for i,b in enumerate(base):
    b16=h(b); bits=int(b16.view(np.uint16))
    cands.append([float(np.uint16(bits+d).view(f16)) for d in (-2,-1,0,1,2)])
for co in itertools.product(*cands):
    e=err(list(co))[0]
    if e<best[0]: best=(e,list(co))
print('best',best)
