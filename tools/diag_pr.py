import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from cozo_b200 import capi
from oracle import oracle as O
from tests.util import rmat_edges
capi.init(0)
for scale,iters in [(14,50),(14,10),(16,10)]:
    n, src, dst = rmat_edges(scale, 16, 0x5EED0004)
    g = capi.Graph(n, src, dst); o = O.OracleGraph(n, src, dst)
    gs, git, gerr, ms = g.pagerank(0.85, 0.0, iters)
    os_, oit, oerr = o.pagerank(0.85, 0.0, iters, n_threads=16)
    rel = np.abs(gs-os_)/os_
    indeg = np.bincount(dst, minlength=n)
    w = np.argsort(-rel)[:5]
    print(scale, iters, 'maxrel', rel.max(), 'worst nodes indeg', indeg[w], 'rel', rel[w], 'max indeg', indeg.max(), 'ms', ms)
    # f64 truth
    import scipy.sparse as sp
    outdeg = np.bincount(src, minlength=n).astype(np.float64)
    A = sp.csr_matrix((np.ones(len(src)), (dst, src)), shape=(n,n))
    s = np.full(n, 1.0/n)
    for _ in range(iters):
        c = np.where(outdeg>0, s/np.maximum(outdeg,1), 0)
        s = (1-0.85)/n + 0.85*(A@c)
    print('  gpu vs f64 truth', np.max(np.abs(gs-s)/s), ' oracle vs truth', np.max(np.abs(os_-s)/s))
n = 5000
src = np.concatenate([np.arange(1, n), np.zeros(n - 1)]).astype(np.uint32)
dst = np.concatenate([np.zeros(n - 1), np.arange(1, n)]).astype(np.uint32)
g = capi.Graph(n, src, dst); o = O.OracleGraph(n, src, dst)
gs, _, _, _ = g.pagerank(0.85, 0.0, 5); os_, _, _ = o.pagerank(0.85, 0.0, 5)
rel = np.abs(gs-os_)/os_
print('hub test', rel.max(), np.argmax(rel), gs[:3], os_[:3])
