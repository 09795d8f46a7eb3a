"""Prototype of the GPU march algorithm in numpy/python, checked against the oracle."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
import oracle

def march(m, W,H,nn,r):
    span=2*r+1
    sx,sy=W*(nn[0]//2),H*(nn[1]//2)
    reg=m[sy-r:sy+H+r, sx-r:sx+W+r].astype(np.int64)
    PH,PW=reg.shape
    # vstart map
    vstart=np.zeros_like(reg)
    for c in range(PW):
        last={};start={}
        for p in range(PH):
            s=int(reg[p,c])
            if s not in last or p-last[s]>span: start[s]=p
            last[s]=p
            vstart[p,c]=start[s]
    items=[];weights=[];offs=[0]
    inv=np.float32(1.0)/np.float32(span*span)
    for y in range(H):
        lst=[]  # [item,cnt]
        for c in range(PW):
            col=reg[y:y+span,c]
            vals,cn=np.unique(col,return_counts=True)
            cin=dict(zip(vals.tolist(),cn.tolist()))
            cout={}
            if c>=span:
                v2,c2=np.unique(reg[y:y+span,c-span],return_counts=True)
                cout=dict(zip(v2.tolist(),c2.tolist()))
            inlist={e[0] for e in lst}
            for e in lst:
                e[1]+=cin.get(e[0],0)-cout.get(e[0],0)
            # deaths first
            lst=[e for e in lst if e[1]!=0]
            born=[s for s in cin if s not in inlist]
            keyed=[]
            for s in born:
                # last occurrence in window
                p=y+2*r
                while reg[p,c]!=s: p-=1
                keyed.append((int(vstart[p,c]),s))
            keyed.sort()
            for k,s in keyed: lst.append([s,cin[s]])
            if c>=2*r:
                for e in lst:
                    items.append(e[0]); weights.append(np.float32(e[1])*inv)
                offs.append(len(items))
    return np.array(items,np.uint16),np.array(weights,np.float32),np.array(offs,np.uint32)

rng=np.random.default_rng(1)
for t in range(40):
    W=int(rng.integers(4,20));H=int(rng.integers(4,20))
    r=int(rng.integers(1,min(W,H)//2+1))*2
    r=min(r,(min(W,H)//2)*2)
    if r==0: r=2
    B=int(rng.integers(1,30))
    kind=t%3
    if kind==0: m=rng.integers(0,B,(3*H,3*W))
    elif kind==1:
        bs=int(rng.integers(2,8)); m=rng.integers(0,B,(3*H//bs+1,3*W//bs+1)).repeat(bs,0).repeat(bs,1)[:3*H,:3*W]
    else:
        m=np.where(rng.random((3*H,3*W))<0.9,0,rng.integers(0,B,(3*H,3*W)))
    m=m.astype(np.uint16)
    a=oracle.run_port(m,(W,H),(3,3),r)
    b=march(m,W,H,(3,3),r)
    ok=all((x.shape==y.shape and (x.view(np.uint8)==y.view(np.uint8)).all()) for x,y in zip(a,b))
    print(t,W,H,r,B,kind,ok)
    assert ok
