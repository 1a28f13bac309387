import sys, math, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from nms_levels_enumerate import boxes_for, grid_dim
def count(b,t,levels,LMAX=16):
    n=len(b); w=b[:,2]-b[:,0]; h=b[:,3]-b[:,1]; cx=0.5*b[:,0]+0.5*b[:,2]; cy=0.5*b[:,1]+0.5*b[:,3]
    reach=(1-t)/t*1.01+0.01
    s=np.maximum(w,h); e=np.frexp(s)[1]-1; emin=e.min()
    lv=np.minimum(e-emin,LMAX-1) if levels else np.zeros(n,int); L=lv.max()+1
    dL=math.ceil(-math.log2(0.98*t))
    mnx,mxx,mny,mxy=cx.min(),cx.max(),cy.min(),cy.max()
    Kl=np.bincount(lv,minlength=L); G=[grid_dim(k) if k>0 else 0 for k in Kl]
    cap=grid_dim(n)**2
    while sum(g*g for g in G)>cap: G[int(np.argmax(G))]-=1
    tot=0
    for m in range(L):
        if G[m]==0: continue
        fx=lambda x: np.clip(np.floor((x-mnx)*(G[m]/(mxx-mnx))),0,G[m]-1).astype(int)
        fy=lambda y: np.clip(np.floor((y-mny)*(G[m]/(mxy-mny))),0,G[m]-1).astype(int)
        sel=lv==m
        H=np.zeros((G[m]+1,G[m]+1)); np.add.at(H,(fy(cy[sel])+1,fx(cx[sel])+1),1); I=H.cumsum(0).cumsum(1)
        def rect(y0,y1,x0,x1): return I[y1+1,x1+1]-I[y0,x1+1]-I[y1+1,x0]+I[y0,x0]
        for l in range(max(0,m-dL),m+1):
            a=lv==l
            if not a.any(): continue
            rx=reach*w[a]; ry=reach*h[a]
            if l==m:  # forward half approx: own row from ax, rows below full
                ay=fy(cy[a]); ax=fx(cx[a]); y1=fy(cy[a]+ry); xl=fx(cx[a]-rx); x1=fx(cx[a]+rx)
                tot+=rect(ay,ay,ax,x1).sum()/1.0
                below=np.where(y1>ay, rect(np.minimum(ay+1,G[m]-1),y1,xl,x1),0); tot+=below.sum()
            else:
                tot+=rect(fy(cy[a]-ry),fy(cy[a]+ry),fx(cx[a]-rx),fx(cx[a]+rx)).sum()
    return int(tot),L,G
for dist,t in (("R",0.65),("T",0.65)):
    b=boxes_for(dist)
    print(dist,len(b),"single",count(b,t,False),"multi",count(b,t,True))
