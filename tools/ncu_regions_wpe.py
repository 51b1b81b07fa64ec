import csv,sys,collections,bisect,re
rows=list(csv.reader(open(sys.argv[1]))); unit=float(sys.argv[2])
src=open('hsr_env_b200/csrc/hsrb_wpe.cuh').read().splitlines()
marks=[]
for i,l in enumerate(src):
    m=re.search(r'// -{4,} ?(.*)$',l) or re.search(r'// ---- (.*)$',l)
    if m: marks.append((i+1,m.group(1)[:50]))
    if '__device__ __noinline__' in l or '__device__ __forceinline__' in l or '__global__' in l: marks.append((i+1,'fn:'+l.strip()[:50]))
    if 'auto rows_jar' in l or 'auto cost_of' in l or 'auto ls_eval' in l: marks.append((i+1,'lambda:'+l.strip()[:40]))
marks.sort(); st=[a for a,_ in marks]
cur=None;key=None
agg=collections.defaultdict(lambda:[0,0,0])
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)<9 or r[0] in ('Line No','Function Name'): continue
    if r[0]!='':
        l=int(r[0])
        if cur=='hsrb_wpe.cuh': key='wpe:'+(marks[bisect.bisect_right(st,l)-1][1] if l>=st[0] else 'hdr')
        else: key=cur
        agg[key][1]+=int(r[7]); agg[key][2]+=int(r[6])
    else: agg[key][0]+=1
te=sum(v[1] for v in agg.values()); ts=sum(v[2] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:60s} static {v[0]:5d} exec/unit {v[1]/unit:7.0f} ({v[1]/te:5.3f}) samp {v[2]/ts:5.3f}")
