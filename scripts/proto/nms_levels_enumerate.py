import sys, math, numpy as np, torch
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle import oracle as O
from vision_conglomerate_b200 import synth

def grid_dim(K):
    G = 1
    while G < 2048 and 2 * (G + 1) * (G + 1) <= K: G += 1
    return G

def boxes_for(dist, S=640, seed=7):
    raws = synth.raw_head_outputs(1, S, S, 80, dist, seed)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    pr = O.decode_inference(raws, anc, S, S, None)  # [1,N,85]
    p = pr[0]
    sc = 1/(1+np.exp(-p[:,1:81].max(1))) * 1/(1+np.exp(-p[:,0]))
    xywh = p[:,81:85].copy(); xywh[:,2:] += 4
    x1 = xywh[:,0]-xywh[:,2]/2; y1 = xywh[:,1]-xywh[:,3]/2
    b = np.stack([x1,y1,x1+xywh[:,2],y1+xywh[:,3]],1).astype(np.float32)
    keep = sc > 0.001
    return b[keep]

def iou_pairs_brute(b, t):
    n=len(b); out=set()
    area=(b[:,2]-b[:,0])*(b[:,3]-b[:,1])
    for i in range(n):
        xx1=np.maximum(b[i,0],b[i+1:,0]); yy1=np.maximum(b[i,1],b[i+1:,1])
        xx2=np.minimum(b[i,2],b[i+1:,2]); yy2=np.minimum(b[i,3],b[i+1:,3])
        w=np.clip(xx2-xx1,0,None); h=np.clip(yy2-yy1,0,None)
        inter=w*h; iou=inter/(area[i]+area[i+1:]-inter)
        for j in np.nonzero(iou>t)[0]: out.add((i,i+1+int(j)))
    return out

def simulate(b, t, levels=True, LMAX=16):
    n=len(b)
    w=b[:,2]-b[:,0]; h=b[:,3]-b[:,1]; cx=0.5*b[:,0]+0.5*b[:,2]; cy=0.5*b[:,1]+0.5*b[:,3]
    reach=(1-t)/t*1.01+0.01
    s=np.maximum(w,h)
    e=np.frexp(s)[1]-1   # exponent: s in [2^e, 2^(e+1))
    emin=e.min()
    lv=np.minimum(e-emin,LMAX-1) if levels else np.zeros(n,int)
    L=lv.max()+1
    dL=math.ceil(-math.log2(0.98*t))
    mnx,mxx,mny,mxy=cx.min(),cx.max(),cy.min(),cy.max()
    Kl=np.bincount(lv,minlength=L)
    G=[grid_dim(k) if k>0 else 0 for k in Kl]
    cap=grid_dim(n)**2
    while sum(g*g for g in G)>cap:
        i=int(np.argmax(G)); G[i]-=1
    def cxl(x,m): return int(min(max(math.floor((x-mnx)*(G[m]/(mxx-mnx) if mxx>mnx else 0)),0),G[m]-1))
    def cyl(y,m): return int(min(max(math.floor((y-mny)*(G[m]/(mxy-mny) if mxy>mny else 0)),0),G[m]-1))
    cells={}
    cellof=np.zeros(n,int)
    for i in range(n):
        m=lv[i]; c=(m,cyl(cy[i],m),cxl(cx[i],m)); cells.setdefault(c,[]).append(i); cellof[i]=0
    found=set(); entries=0
    # order within cell = index order (list order)
    for i in range(n):
        l=lv[i]; rx=reach*w[i]; ry=reach*h[i]
        for m in range(l,min(l+dL,L-1)+1):
            if G[m]==0: continue
            if m==l:
                ay=cyl(cy[i],m); ax=cxl(cx[i],m)
                y1=cyl(cy[i]+ry,m); xl=cxl(cx[i]-rx,m); x1=cxl(cx[i]+rx,m)
                for gy in range(ay,y1+1):
                    x0=ax if gy==ay else xl
                    for gx in range(x0,x1+1):
                        for j in cells.get((m,gy,gx),[]):
                            if gy==ay and gx==ax and j<=i: continue
                            entries+=1
                            found.add((min(i,j),max(i,j)))
            else:
                y0=cyl(cy[i]-ry,m); y1=cyl(cy[i]+ry,m); xl=cxl(cx[i]-rx,m); x1=cxl(cx[i]+rx,m)
                for gy in range(y0,y1+1):
                    for gx in range(xl,x1+1):
                        for j in cells.get((m,gy,gx),[]):
                            entries+=1
                            found.add((min(i,j),max(i,j)))
    return found, entries, L, G

if __name__=="__main__":
    dist=sys.argv[1]; t=float(sys.argv[2]); nmax=int(sys.argv[3]) if len(sys.argv)>3 else 100000
    b=boxes_for(dist)[:nmax]
    print(dist, "boxes", len(b))
    f1,e1,L1,G1=simulate(b,t,False)
    f2,e2,L2,G2=simulate(b,t,True)
    print("single-level entries", e1, "G", G1, "| multi-level entries", e2, "L", L2, "G", G2)
    if len(b)<=4000:
        truth=iou_pairs_brute(b,t)
        print("true pairs", len(truth), "missed single", len(truth-f1), "missed multi", len(truth-f2))
