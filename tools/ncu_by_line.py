"""Executed warp instructions, stall samples and active lanes per source LINE of one kernel.

    cuobjdump -xelf all libtmgpu.so; for f in *.cubin; do nvdisasm -g -c $f; done > /tmp/all.sass
    ncu -i report.ncu-rep --page source --csv --kernel-name regex:<kernel> > src.csv
    python tools/ncu_by_line.py <mangled kernel name> src.csv [min percent]
Joins the n-th SASS instruction of the kernel (nvdisasm order) with the n-th row of the source page."""
import re,csv,sys
from collections import defaultdict
kname=sys.argv[1]; srccsv=sys.argv[2]; thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.6
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
ie=ix['Instructions Executed']; te=ix['Thread Instructions Executed']; ss=ix['# Samples']
data=[]
for r in rows[2:]:
    try: data.append((float(r[ie]),float(r[te]),r[ix['Source']].strip(), float(r[ss])))
    except: pass
print('sass lines', len(ins), 'profile lines', len(data))
agg=defaultdict(lambda:[0,0,0])
for k in range(min(len(ins),len(data))):
    a=agg[ins[k][0]]; a[0]+=data[k][0]; a[1]+=data[k][1]; a[2]+=data[k][3]
tot=sum(v[0] for v in agg.values()); tots=sum(v[2] for v in agg.values())
import glob,os
files={}
for f in glob.glob('/root/repo/tokenmonster_b200/csrc/*'):
    try: files[os.path.basename(f)]=open(f).read().split('\n')
    except: pass
for key,v in sorted(agg.items(), key=lambda kv:(kv[0] or ('?',0))):
    f,l=key if key else ('?',0)
    if v[0]/tot<thr/100 and v[2]/tots<thr/100: continue
    text=files.get(f,[''])[l-1].strip()[:86] if f in files and l-1 < len(files[f]) else ''
    print('%5.1f%% inst  %5.1f%% samples  act %4.1f  %s:%d  %s'%(100*v[0]/tot, 100*v[2]/tots, v[1]/v[0] if v[0] else 0, f,l,text))
print('total warp inst %.4g'%tot)
