"""Walk-length statistics of a vocabulary's prefix trie on sample text, and a simulation of the
warp-level walk pool: how many loop iterations 32 lanes need for N walks handed out
greedily, against the ideal (total steps / 32), for the plain trie and for a trie whose
single-child chains are compressed into one step.  CPU only; uses the oracle for token boundaries.

    python tools/sim_walk_pool.py            # config c2
"""
import sys, random, collections
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from workloads import configs, corpus
from oracle import oracle
vb = configs.vocab_bytes('c2')
n_info = vb[17]|vb[18]<<8|vb[19]<<16
p=24; keys=[]
for _ in range(n_info):
    l=vb[p]; keys.append(vb[p+1:p+1+l]); p+=1+l+2+9+4
children=collections.defaultdict(dict); nid={b'':0}
for k in keys:
    for d in range(1,len(k)+1):
        pre=k[:d]
        if pre not in nid:
            nid[pre]=len(nid); children[nid[k[:d-1]]][pre[-1]]=nid[pre]
o = oracle.OracleVocab(vb,'port')
base = corpus.base_text(5, 1<<20, 0.3)
d, off = corpus.make_batch(base, 1500, 64, 960, 2, 'cpu'); d=d.numpy(); off=off.numpy()
def walk(txt,pos):
    node=0; steps=0; cs=0; inchain=False; chainlen=0
    q=pos
    while q<len(txt) and q-pos<40:
        c=children.get(node,{})
        b=txt[q]
        if b not in c:
            steps+=1   # the failing probe (often filtered, still one loop iteration)
            if not(len(c)==1 and node!=0 and inchain): cs+=1
            break
        if len(c)==1 and node!=0:
            if not inchain or chainlen>=16: cs+=1; chainlen=0; inchain=True
            chainlen+=1
        else:
            cs+=1; inchain=False
        node=c[b]; q+=1; steps+=1
    return steps, cs
allw=[]
for s in range(1500):
    raw=d[off[s]:off[s+1]].tobytes()
    enc=o.normalize(raw)
    ids,_=o.tokenize(raw)
    pos=0
    for t in ids:
        allw.append(walk(enc,pos))
        l=len(o.decode([int(t)],raw=True))
        if l>2: allw.append(walk(enc,pos+l-1))   # an alternative-like position inside the token
        pos+=l
a=np.array(allw)
print('walks',len(a),'mean steps',a[:,0].mean(),'mean compressed',a[:,1].mean())
print('hist steps',np.bincount(a[:,0])[:25])
print('hist comp ',np.bincount(a[:,1])[:25])
# pool simulation: 64 tasks, 32 lanes, greedy dynamic
rng=np.random.default_rng(0)
def sim(col, ntask=64, trials=4000, extra=0):
    tot_it=0; tot_steps=0
    for _ in range(trials):
        t=a[rng.integers(0,len(a),ntask),col]+extra
        lanes=np.zeros(32,int)
        for x in t:
            k=lanes.argmin(); lanes[k]+=x
        tot_it+=lanes.max(); tot_steps+=t.sum()
    return tot_it/trials, tot_steps/trials/32
for nt in (48,64,80):
    print('ntask',nt,'plain iters, ideal',sim(0,nt),'compressed',sim(1,nt))
