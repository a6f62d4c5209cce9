"""ANALYSIS SCRIPT (test infrastructure, not collected by pytest): numpy prototype of cv2's 5-point minimal solver used to find
out WHY the CUDA solver's inlier counts differed from cv2's on ill-conditioned samples:
  * basis_cv: the null-space basis cv::SVD(FULL_UV) returns for a 5x9 matrix -- Gram-Schmidt of fixed +-1/9 vectors whose signs
    are bit 8 of cv::RNG(0x12345678) draws;
  * solve_poly_dk: cv::solvePoly (Gauss-Seidel Durand-Kerner from (1+i)^k, 300 iterations).
Run as a script it searches the conventions (ordering of the 9 unknowns, which point set is "1") for the one that reproduces
cv2.findEssentialMat's five-point solutions IN ORDER: ('rowmajor', swap=False) does on 40 of 40 random samples."""
import sys, numpy as np, cv2
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pose_np as P

def cv_rng_signs():
    state=0x12345678
    out=[]
    for _ in range(4*9):
        state=((state&0xFFFFFFFF)*4164903690+(state>>32))&0xFFFFFFFFFFFFFFFF
        out.append(1.0 if (state&256)!=0 else -1.0)
    return np.array(out).reshape(4,9)/9.0
SIGNS=cv_rng_signs()

def basis_cv(Q):
    """null-space basis as cv::SVD(FULL_UV) builds it for a 5x9 matrix: rows 5..8 = Gram-Schmidt of fixed +-1/9 vectors"""
    _,_,vt=np.linalg.svd(Q,full_matrices=False)   # orthonormal row space (5x9)
    rows=[v for v in vt]
    for i in range(4):
        a=SIGNS[i].copy()
        for it in range(2):
            for r in rows:
                a=a-(a@r)*r
                s=np.abs(a).sum(); a=a/s
        a=a/np.linalg.norm(a)
        rows.append(a)
    return np.array(rows[5:9])

def solve_poly_dk(c, max_iters=300):
    """cv::solvePoly: coefficients c[0] + c[1] z + ... c[n] z^n (ascending)."""
    n=len(c)-1
    while n>1 and abs(c[n])<=np.finfo(float).eps: n-=1
    roots=[]; p=1+0j
    for i in range(n):
        roots.append(p); p=p*(1+1j)
    for it in range(max_iters):
        maxdiff=0.0
        for i in range(n):
            p=roots[i]; num=complex(c[n]); den=complex(c[n])
            for j in range(n):
                num=num*p+c[n-j-1]
                if j!=i:
                    d=p-roots[j]
                    if d!=0: den=den*d
            num=num/den
            roots[i]=p-num
            maxdiff=max(maxdiff,abs(num))
        if maxdiff<=0: break
    return roots

def five_point_cvlike(x1,x2,order="colmajor",swap=False,rootfinder="dk"):
    if swap: x1,x2=x2,x1
    Q=np.empty((5,9))
    for i in range(5):
        a,b=x1[i]; c,d=x2[i]
        if order=="colmajor": Q[i]=[c*a,d*a,a,c*b,d*b,b,c,d,1.0]
        else: Q[i]=[c*a,c*b,c,d*a,d*b,d,a,b,1.0]
    EE=basis_cv(Q)
    A=P.five_point_constraints(EE)
    try: Bm=np.linalg.solve(A[:,:10],A[:,10:])
    except np.linalg.LinAlgError: return []
    def rmz(e,f):
        re,rf=Bm[e],Bm[f]
        px=np.array([0.0,re[0],re[1],re[2]])-np.array([rf[0],rf[1],rf[2],0.0])
        py=np.array([0.0,re[3],re[4],re[5]])-np.array([rf[3],rf[4],rf[5],0.0])
        p1=np.array([0.0,re[6],re[7],re[8],re[9]])-np.array([rf[6],rf[7],rf[8],rf[9],0.0])
        return px,py,p1
    B=[rmz(4,5),rmz(6,7),rmz(8,9)]
    pm=np.polymul
    det=(pm(pm(B[0][0],B[1][1])-pm(B[0][1],B[1][0]),B[2][2])+pm(pm(B[0][1],B[1][2]),B[2][0])-pm(pm(B[0][2],B[1][1]),B[2][0])
         +pm(pm(B[0][2],B[1][0]),B[2][1])-pm(pm(B[0][0],B[1][2]),B[2][1]))
    det=np.atleast_1d(det)
    if len(det)<11: det=np.concatenate([np.zeros(11-len(det)),det])
    if rootfinder=="dk": roots=solve_poly_dk(det[::-1])
    else: roots=list(np.roots(det))
    sols=[]
    for r in roots:
        if abs(r.imag)>1e-10: continue
        z=r.real
        Bz=np.array([[np.polyval(B[i][0],z),np.polyval(B[i][1],z),np.polyval(B[i][2],z)] for i in range(3)])
        _,_,vt=np.linalg.svd(Bz); xy1=vt[2]
        if abs(xy1[2])<1e-10: continue
        x,y=xy1[0]/xy1[2],xy1[1]/xy1[2]
        Ev=x*EE[0]+y*EE[1]+z*EE[2]+EE[3]; Ev=Ev/np.linalg.norm(Ev)
        M=Ev.reshape(3,3)
        if order=="colmajor": pass
        sols.append(M)
    return sols

if __name__=="__main__":
    rng=np.random.default_rng(3)
    K=np.eye(3)
    agree={}
    for trial in range(40):
        X=rng.uniform(-1,1,(5,3))+np.array([0,0,4.0])
        R,_=cv2.Rodrigues(rng.normal(size=3)*0.1); t=rng.normal(size=3)*0.3
        x1=X[:,:2]/X[:,2:]; Xc=(R@X.T).T+t; x2=Xc[:,:2]/Xc[:,2:]
        Ecv=cv2.findEssentialMat(x1,x2,K,method=cv2.RANSAC,prob=0.999,threshold=1.0)[0]
        cvs=[Ecv[3*i:3*i+3] for i in range(len(Ecv)//3)]
        for order in ("colmajor","rowmajor"):
            for swap in (False,True):
                for tr in (False,True):
                    ms=five_point_cvlike(x1,x2,order,swap)
                    if tr: ms=[m.T for m in ms]
                    ok = len(ms)==len(cvs) and all(min(np.abs(a-b).max(),np.abs(a+b).max())<1e-6 for a,b in zip(ms,cvs))
                    same_set = len(ms)==len(cvs) and all(any(min(np.abs(a-b).max(),np.abs(a+b).max())<1e-6 for b in cvs) for a in ms)
                    k=(order,swap,tr); agree.setdefault(k,[0,0]); agree[k][0]+=ok; agree[k][1]+=same_set
    for k,v in agree.items(): print(k,"ordered-equal",v[0],"set-equal",v[1],"of 40")
