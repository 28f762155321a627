"""Hit rate of a direct-mapped shared-memory cache of automaton transitions, simulated on the real
state sequence of the bench text (1024 lanes of one SM sharing the cache, one byte per lane per
step).  Uses the host build of the product's tables (tests/host_sim.cpp).  Result on c2: 15 / 30 /
39 / 50 % hits at 2 K / 8 K / 16 K / 32 K entries -- quoted in DESIGN.md section 4.6.
    python tools/sim_transition_cache.py [config]"""
import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
from workloads import configs, corpus
name=sys.argv[1] if len(sys.argv)>1 else 'c2'
lib=C.CDLL('/root/repo/tests/_build/libhostsim.so')
lib.hs_load.restype=C.c_void_p; lib.hs_load.argtypes=[C.c_char_p,C.c_size_t,C.c_char_p,C.c_int]
lib.hs_table.restype=C.c_long; lib.hs_table.argtypes=[C.c_void_p,C.c_char_p,C.c_size_t,C.c_int,C.c_void_p,C.c_void_p,C.c_size_t]
vb=configs.vocab_bytes(name); err=C.create_string_buffer(256)
h=lib.hs_load(vb,len(vb),err,256)
cfg=configs.CONFIGS[name]
base=corpus.base_text(cfg['seed'],16<<20,cfg['code'])
data,off=corpus.make_batch(base,4096,cfg['lo'],cfg['hi'],cfg['seed']+1,'cpu')
data=data.numpy(); off=off.numpy()
LANES=1024
seqs=[]
for s in range(4096):
    raw=data[off[s]:off[s+1]].tobytes()
    cap=4*len(raw)+64
    w=np.zeros(cap,np.uint32); t=np.zeros(cap,np.uint8)
    n=lib.hs_table(h,raw,len(raw),0,w.ctypes.data,t.ctypes.data,cap)
    w=w[:n]; t=t[:n]
    st=(w&0xFFFFF).astype(np.uint64)
    # transition consumed at p: from state of p+1 with byte t[p]; order right-to-left
    keys=((st[1:]<<8)|t[:-1].astype(np.uint64))[::-1]
    depth3=None
    seqs.append(keys)
for N in (2048,8192,16384,32768):
    # 4 CTAs worth: take LANES lanes; each lane processes strings sequentially (string k, k+LANES...)
    cache=np.full(N,-1,np.int64); hits=0; tot=0
    lanes=[np.concatenate(seqs[l::LANES]) for l in range(LANES)]
    T=min(len(x) for x in lanes)
    M=np.stack([x[:T] for x in lanes]).astype(np.int64)
    for t_ in range(T):
        k=M[:,t_]
        idx=((k*0x9E3779B1)&0xFFFFFFFF)>>(32-int(np.log2(N)))
        hit=cache[idx]==k
        hits+=int(hit.sum()); tot+=LANES
        cache[idx]=k
    print(name,'cache entries',N,'hit rate %.3f'%(hits/tot))
