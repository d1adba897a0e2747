import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import synthetic_stack
from evcont_b200.mol import synthetic_mol, ao_bundle
from oracle import gradients as og, subspace as osub
import scipy.linalg

def tri(i, j): return i*(i+1)//2 + j
def run(n, natm, N, layout, seed):
    ovlp, one, two = synthetic_stack(n, N, seed, layout)
    # break the a<->b symmetry of the full layouts to test RG vs RH
    mol = synthetic_mol(n, natm, seed=seed+1)
    oe, ogr = og.get_energy_with_grad(mol, one, two, ovlp)
    ao = ao_bundle(mol)
    n2 = n*n; npair = n*(n+1)//2
    pairs = [(i, j) for i in range(n) for j in range(i+1)]
    pidx = {}
    for I,(i,j) in enumerate(pairs): pidx[(i,j)] = I; pidx[(j,i)] = I
    L2 = npair*(npair+1)//2
    # restore to full D2 (N,N,n,n,n,n) semantics
    def block(a, b):
        if layout == 6: return two[a, b]
        if layout == 5: return two[tri(a, b)] if a >= b else None
        if layout == 3: return osub.restore_exchange(two[a, b], n)
        if layout == 2: return osub.restore_exchange(two[tri(a, b)], n) if a >= b else None
    def pack(D):
        out = np.zeros(L2)
        for I,(i,j) in enumerate(pairs):
            for K,(k,l) in enumerate(pairs[:I+1]):
                tup = {(i,j,k,l),(j,i,k,l),(i,j,l,k),(j,i,l,k),(k,l,i,j),(l,k,i,j),(k,l,j,i),(l,k,j,i)}
                out[tri(I,K)] = 0.5*sum(D[t] for t in tup)
        return out
    Pt = N*(N+1)//2
    RH = np.zeros((Pt, n2+L2)); RG = np.zeros((Pt, n2+L2))
    for a in range(N):
        for b in range(a+1):
            p = tri(a,b)
            Dab = block(a,b).reshape(n,n,n,n)
            RH[p,:n2] = one[a,b].ravel(); RH[p,n2:] = pack(Dab)
            if a == b or layout in (5,2):
                RG[p,n2:] = RH[p,n2:]
            else:
                RG[p,n2:] = 0.5*(RH[p,n2:] + pack(block(b,a).reshape(n,n,n,n)))
            RG[p,:n2] = one[a,b].ravel() if a==b else 0.5*(one[a,b]+one[b,a]).ravel()
    # step
    S = ao['ovlp']; w, V = np.linalg.eigh(S); X = (V*(w**-0.5))@V.T
    eri = ao['eri']
    E = np.zeros((npair,npair)); 
    for I,(i,j) in enumerate(pairs):
        for K,(k,l) in enumerate(pairs): E[I,K] = eri[i,j,k,l]
    P0 = np.zeros((npair,npair)); sI = np.array([2.0 if i==j else 1.0 for i,j in pairs])
    for A,(a,b) in enumerate(pairs):
        for I,(i,j) in enumerate(pairs): P0[A,I] = X[a,i]*X[b,j] + X[a,j]*X[b,i]
    Q = P0 / sI[:,None]
    T = E @ Q
    h2p = Q.T @ T
    h1 = X.T @ ao['hcore'] @ X
    hvec = np.concatenate([h1.ravel(), np.array([h2p[I,K] for I in range(npair) for K in range(I+1)])])
    Hp = RH @ hvec
    H = np.zeros((N,N))
    for a in range(N):
        for b in range(a+1): H[a,b] = H[b,a] = Hp[tri(a,b)]
    ev, C = scipy.linalg.eigh(H, ovlp)
    c = C[:,0]; E0 = ev[0]
    wts = np.array([ (c[a]*c[a] if a==b else 2*c[a]*c[b]) for a in range(N) for b in range(a+1)])
    out7 = wts @ RG
    gamma = out7[:n2].reshape(n,n)
    Gm = np.zeros((npair,npair))
    for I in range(npair):
        for K in range(I+1):
            v = out7[n2+tri(I,K)]
            Gm[I,K] = Gm[K,I] = v*(2.0 if I==K else 1.0)
    U0 = T @ Gm
    Y = np.zeros((n,n))
    for a in range(n):
        for i in range(n):
            s = 0.0
            for b in range(n):
                for j in range(n):
                    s += X[b,j]*(2.0 if i==j else 1.0)*U0[pidx[(a,b)], pidx[(i,j)]]
            Y[a,i] = 2*s
    W = P0 @ Gm @ P0.T
    hc = ao['hcore']
    Z = hc @ X @ (gamma+gamma.T) + 0.5*Y
    rs = np.sqrt(w)
    G = -1.0/(rs[:,None]*rs[None,:]*(rs[:,None]+rs[None,:]))
    Om = V @ (G*(V.T@Z@V)) @ V.T
    OmS = Om + Om.T
    Pao = X @ gamma @ X.T
    ip1 = ao['eri_ip1']
    T2 = np.zeros((3,n))
    for m in range(n):
        for b in range(n):
            for cc in range(n):
                for d in range(n):
                    T2[:,m] += ip1[:,m,b,cc,d]*W[pidx[(m,b)],pidx[(cc,d)]]
    grad = np.zeros((natm,3))
    for A,(p0,p1) in enumerate(ao['aoslices']):
        for x in range(3):
            grad[A,x] = -(ao['ipovlp'][x,p0:p1]*OmS[p0:p1]).sum() + (ao['hcore_deriv'][A,x]*Pao).sum() - 0.5*T2[x,p0:p1].sum()
    grad += ao['grad_nuc']
    print(n, natm, N, layout, 'dE', abs(E0+ao['e_nuc']-oe), 'dgrad', np.abs(grad-ogr).max())
for layout in (6,5,3,2):
    run(4, 2, 3, layout, 5)
run(5,3,4,6,7)
