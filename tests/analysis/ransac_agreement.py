"""ANALYSIS SCRIPT (test infrastructure, not collected by pytest; run on a GPU box):
    python tests/analysis/ransac_agreement.py <width> <height> <nfeatures> <n_frames>
How often does the CUDA RANSAC end on the same essential matrix as cv2 (|E - E_cv2| < 1e-4), and how often does cv2 end on the
same one as ITSELF when fx moves by one ulp?  Numbers quoted in DESIGN.md section 0 (640x480/500: 291 vs 297 of 300;
1280x1024/500: 286 vs 293; 1280x1024/2000 is tests/test_full_sequence.py)."""
import sys, os, numpy as np, torch, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
W,H,NF,N=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3]),int(sys.argv[4])
def worker(args):
    path,kpath,lo,hi,NF=args
    import cv2; cv2.setNumThreads(1)
    from oracle import cv2_chain
    F=np.load(path,mmap_mode="r"); K=np.load(kpath); out=[]
    prev=cv2_chain.orb_features(np.ascontiguousarray(F[lo]),NF)
    for i in range(lo,hi):
        cur=cv2_chain.orb_features(np.ascontiguousarray(F[i+1]),NF)
        r=cv2_chain.frame_pair(None,None,K,NF,feats_prev=prev,feats_cur=cur)
        K1=K.copy(); K1[0,0]=np.nextafter(K1[0,0],1e9)
        s=cv2_chain.pose_from_points(r["p_prev"],r["p_cur"],K1)
        out.append((r["E"],r["ransac_mask"],s["E"],s["ransac_mask"],len(r["matches"])))
        prev=cur
    return lo,out
if __name__=="__main__":
    import multiprocessing as mp
    from droplet_visual_odometry_b200 import synth,_native
    frames,_,K=synth.render_sequence(N,W,H,device="cuda",start_index=200)
    fh=frames.cpu().numpy()
    with tempfile.TemporaryDirectory() as d:
        p,kp=os.path.join(d,"f.npy"),os.path.join(d,"k.npy"); np.save(p,fh); np.save(kp,K)
        wk=os.cpu_count(); ch=max(4,-(-(N-1)//(wk*3)))
        jobs=[(p,kp,lo,min(lo+ch,N-1),NF) for lo in range(0,N-1,ch)]
        with mp.get_context("spawn").Pool(wk) as pool: parts=pool.map(worker,jobs,chunksize=1)
    ref=[None]*(N-1)
    for lo,o in parts: ref[lo:lo+len(o)]=o
    ctx=_native.Context(W,H,nfeatures=NF,max_frames=min(N,149))
    B=ctx.max_frames-1; same=0; selfsame=0; mi=0; smi=0; nm=[]
    for lo in range(0,N-1,B):
        hi=min(lo+B,N-1)
        ctx.load_frames(frames[lo:hi+1],0); ctx.orb(0,hi-lo+1); ctx.pairs(0,0,hi-lo,K)
        ps=ctx.poses(0,hi-lo)
        for j in range(hi-lo):
            E=ps[j]["E"].reshape(3,3); rE,rm,sE,sm,n=ref[lo+j]
            arr=ctx.pair_arrays(j,ps[j]["n_matches"])
            de=lambda A,B_: min(np.abs(A-B_).max(),np.abs(A+B_).max())
            same+=de(E,rE)<1e-4; selfsame+=de(sE,rE)<1e-4
            mi+=np.array_equal(arr["ransac_mask"]>0,rm>0); smi+=np.array_equal(sm>0,rm>0); nm.append(n)
    print("%dx%d nf=%d: %d pairs, median matches %d; same model as cv2: %d (cv2 vs itself, fx + 1 ulp: %d); identical masks: %d (%d)"%(W,H,NF,N-1,np.median(nm),same,selfsame,mi,smi))
