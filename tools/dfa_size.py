"""Size of the reversed-key Aho-Corasick automaton of a .vocab file: trie nodes, goto edges, DFA
transitions by target depth (what decided the table layout).
    python tools/dfa_size.py file.vocab [...]"""
import sys, struct, collections
def parse(path):
    b=open(path,'rb').read()
    u24=lambda o:b[o]|b[o+1]<<8|b[o+2]<<16
    ninfo=u24(17); maxlen=b[23]; o=24; keys=[]
    for k in range(ninfo):
        L=b[o]; keys.append(b[o+1:o+1+L]); o+=1+L+2+9+4
    return keys,b[0]
for path in sys.argv[1:]:
    keys,cap=parse(path)
    # reversed trie
    child=[{}]; depth=[0]; term=[False]
    for k in keys:
        n=0
        for c in reversed(k):
            d=child[n]
            if c not in d:
                d[c]=len(child); child.append({}); depth.append(depth[n]+1); term.append(False)
            n=d[c]
        term[n]=True
    N=len(child)
    # BFS fail links + full DFA via dict of non-root transitions
    fail=[0]*N
    order=[]
    dq=collections.deque()
    for c,n in child[0].items(): dq.append(n)
    # delta as dict per state only for transitions to depth>=1 (non-root)
    delta=[None]*N
    delta[0]=dict(child[0])
    while dq:
        s=dq.popleft(); order.append(s)
        f=fail[s]
        d=dict(delta[f]); 
        for c,n in child[s].items():
            fail[n]=delta[f].get(c,0)
            d[c]=n; dq.append(n)
        delta[s]=d
    tot=sum(len(d) for d in delta)
    ge2=sum(1 for d in delta for n in d.values() if depth[n]>=2)
    ge3=sum(1 for d in delta for n in d.values() if depth[n]>=3)
    ge4=sum(1 for d in delta for n in d.values() if depth[n]>=4)
    print(path.split('/')[-1],'keys',len(keys),'nodes',N,'goto',N-1,'dfa nonroot',tot,'>=2',ge2,'>=3',ge3,'>=4',ge4, 'avgkey',sum(map(len,keys))/len(keys))
