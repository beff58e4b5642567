"""numpy model of the bit-expansion loops of k_nbh_bits on a C2-like configuration: iterations a warp of 32 particles of one cell needs when the
accept bits are walked per neighbour cell / per row of cells / per plane of cells / in one flat loop (sum over the loops of the longest lane).
Measured on the GPU: 294 iterations per group (per cell), 141 (per plane), as modelled here."""
import numpy as np
from scipy.spatial import cKDTree
rng=np.random.default_rng(1)
a=(4/0.8442)**(1/3); n=12; L=n*a
b=np.array([[0,0,0],[0,.5,.5],[.5,0,.5],[.5,.5,0]])
g=np.stack(np.meshgrid(*[np.arange(n)]*3,indexing='ij'),-1).reshape(-1,3)
x=((g[:,None,:]+b[None])*a).reshape(-1,3)+rng.normal(0,0.12,(n**3*4,3))   # thermal-ish disorder
x%=L
cs=2*a; nc=n//2
cell=np.floor(x/cs).astype(int)%nc
t=cKDTree(x,boxsize=L)
pairs=t.query_pairs(2.8,output_type='ndarray')
pairs=np.concatenate([pairs,pairs[:,::-1]])
i,j=pairs[:,0],pairs[:,1]
d=(cell[j]-cell[i]+1+nc)%nc-1   # -1,0,1
N=len(x)
tot=np.bincount(i,minlength=N)
print('avg entries',tot.mean(),'max',tot.max())
cid=(cell[:,2]*nc+cell[:,1])*nc+cell[:,0]
def summax(keys,nk):
    # per atom counts per key
    cnt=np.zeros((N,nk),int); np.add.at(cnt,(i,keys),1)
    # per cell max over atoms
    s=0; ncell=0
    for c in np.unique(cid):
        m=cnt[cid==c].max(0); s+=m.sum(); ncell+=1
    return s/ncell
print('per-cell loops Σmax', summax((d[:,2]+1)*9+(d[:,1]+1)*3+d[:,0]+1,27))
print('per-row Σmax', summax((d[:,2]+1)*3+(d[:,1]+1),9))
print('per-plane Σmax', summax(d[:,2]+1,3))
print('flat', summax(np.zeros(len(i),int),1))
