"""Executed warp instructions and active lanes per source region of one kernel.

    cuobjdump -xelf all libtmgpu.so; nvdisasm -g -c *.cubin > /tmp/all.sass      # SASS with line info
    ncu -i report.ncu-rep --page source --csv > src.csv
    python tools/ncu_by_source.py <mangled kernel name> src.csv
Joins the n-th SASS instruction of the kernel (nvdisasm order) with the n-th row of the source
page; prints 10-line source buckets that hold >= 0.4 % of the warp instructions."""
import re,csv,sys
kname=sys.argv[1]; srccsv=sys.argv[2]
lines=open('/tmp/all.sass').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+kname)][0]
cur=None; ins=[]
for l in lines[start+1:]:
    if l.startswith('//-----') : break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((cur,m.group(2)))
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
ie=ix['Instructions Executed']; te=ix['Thread Instructions Executed']
data=[]
for r in rows[2:]:
    try: data.append((float(r[ie]),float(r[te]),r[ix['Source']].strip()))
    except: pass
print('sass lines', len(ins), 'profile lines', len(data))
from collections import defaultdict
agg=defaultdict(lambda:[0,0])
for k in range(min(len(ins),len(data))):
    agg[ins[k][0]][0]+=data[k][0]; agg[ins[k][0]][1]+=data[k][1]
tot=sum(v[0] for v in agg.values())
import glob,os
files={}
for f in glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tokenmonster_b200', 'csrc', '*')):
    files[os.path.basename(f)]=open(f).read().split('\n')
# group by ranges of source lines (per file) in buckets
rng=defaultdict(lambda:[0,0])
for key,v in agg.items():
    f,l=key if key else ('?',0)
    rng[(f,l//10*10)][0]+=v[0]; rng[(f,l//10*10)][1]+=v[1]
for key,v in sorted(rng.items()):
    f,l=key
    if v[0]/tot<0.004: continue
    text=files.get(f,[''])[l].strip()[:70] if f in files and l < len(files[f]) else ''
    print('%5.1f%% act %4.1f  %s:%d-%d  %s'%(100*v[0]/tot, v[1]/v[0] if v[0] else 0, f,l,l+9,text))
print('total warp inst %.4g'%tot)
