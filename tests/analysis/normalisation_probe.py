"""ANALYSIS SCRIPT (test infrastructure, not collected by pytest): which arithmetic does cv.findEssentialMat use to K-normalise
points?  Five-point calls with K are compared, bit for bit, with calls on pre-normalised coordinates and K = I for six candidate
forms.  Result with cv2 4.13.0 on an AVX2/FMA host: only fma(p, 1/f, -(c * (1/f))) reproduces cv2 (120 of 120); this is what
mathcore.cuh's cv_normalize_coord and oracle/pose_np.normalize_points implement."""
import numpy as np, cv2, math
K=np.array([[1173.854081,0,640.0],[0,1170.565083,512.0],[0,0,1.0]])
K2=np.array([[1173.854081,0,747.788206],[0,1170.565083,574.700374],[0,0,1.0]])
rng=np.random.default_rng(0)
def fma(a,b,c):
    # exact fma via math.fma if available (py3.13) else emulate with fractions
    from fractions import Fraction
    out=np.empty_like(a)
    for i,(x,y,z) in enumerate(zip(a.ravel(),np.broadcast_to(b,a.shape).ravel(),np.broadcast_to(c,a.shape).ravel())):
        out.ravel()[i]=float(Fraction(float(x))*Fraction(float(y))+Fraction(float(z)))
    return out
cands={
 "a (p-c)/f": lambda p,c,f: (p-c)/f,
 "b (p-c)*(1/f)": lambda p,c,f: (p-c)*(1.0/f),
 "c p*(1/f)+(-c*(1/f))": lambda p,c,f: p*(1.0/f)+(-c*(1.0/f)),
 "d p*(1/f)-c/f": lambda p,c,f: p*(1.0/f)-c/f,
 "e fma(p,1/f,-c*(1/f))": lambda p,c,f: fma(p,1.0/f,-c*(1.0/f)),
 "f fma(p,1/f,-c/f)": lambda p,c,f: fma(p,1.0/f,-(c/f)),
}
score={k:0 for k in cands}
N=60
for KK in (K,K2):
  for t in range(N):
    p1=(rng.uniform(0,1280,(5,2))*16).round()/16; p2=p1+rng.normal(size=(5,2))*8
    p1=p1.astype(np.float32).astype(np.float64); p2=p2.astype(np.float32).astype(np.float64)
    E0=cv2.findEssentialMat(p1,p2,KK,method=cv2.RANSAC,prob=0.999,threshold=1.0)[0]
    for name,fn in cands.items():
        a=np.stack([fn(p1[:,0],KK[0,2],KK[0,0]),fn(p1[:,1],KK[1,2],KK[1,1])],1)
        b=np.stack([fn(p2[:,0],KK[0,2],KK[0,0]),fn(p2[:,1],KK[1,2],KK[1,1])],1)
        E1=cv2.findEssentialMat(a,b,np.eye(3),method=cv2.RANSAC,prob=0.999,threshold=1.0)[0]
        if E0 is not None and E1 is not None and E0.shape==E1.shape and np.array_equal(E0,E1): score[name]+=1
print(score,"of",2*N)
