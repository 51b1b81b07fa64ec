"""Per-source-region dynamic instruction counts of the fast-path kernel from an ncu capture.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_regions.py src.csv <warps x substeps in the captured launch>

Regions are found by marker comments in hsr_env_b200/csrc/hsrb_push.cuh and hsr_core.h (needs -lineinfo)."""
import csv,collections,bisect,sys,re
rows=list(csv.reader(open(sys.argv[1])))
nwarpsub=float(sys.argv[2])
src=open('hsr_env_b200/csrc/hsrb_push.cuh').read().splitlines()
markers=[('impedance5','__device__ __forceinline__ double impedance5'),('collision:cull','PUSH_COLLISION_ATTR int push_collision'),('collision:jobs','while (__any_sync(0xffffffffu, bits != 0))'),('collision:other','Geom<float> A, B;'),('kernel:setup','hsrb_push_kernel(const __grid_constant__'),('poses','// ---------------------------------------------------------------- poses'),('limits','active joint limits'),('collision:call','// ---------------------------------------------------------------- collision (B.3)'),('smooth','smooth forces (closed form'),('rows','constraint rows: one contact per lane'),('solver:lambdas','Newton solver (B.7)'),('solver:init','bool solving = nefc_true != 0;'),('newton:grad','// ---- gradient component'),('newton:H','// ---- Hessian J^T (cone Hessians) J'),('newton:chol','// ---- Cholesky H'),('newton:fwd/bwd','// ---- forward solve'),('newton:jv','// ---- jv = J search'),('linesearch','// ---- exact line search'),('newton:update','if (alpha == 0.f) stop = true;'),('goal+euler','goal test on the poses'),('store','results: HBM once per action')]
push=[]
for name,pat in markers:
    for i,l in enumerate(src):
        if pat in l: push.append((i+1,name)); break
    else: print('marker not found',name)
push.sort()
pst=[a for a,_ in push]
core_src=open('hsr_env_b200/csrc/hsr_core.h').read().splitlines()
cm=[('groups','struct HostGrp'),('vec3/quat','template <typename T> struct V3'),('workspace','struct WS {'),('kinematics','HSR_HDC void kinematics_lane0'),('geom util','template <typename T> struct Geom'),('make_frame','HSR_HD void make_frame'),('add_contact','HSR_HDC void add_contact'),('mpr:tri','HSR_HD T origin_tri_dist2'),('hull_scan4','__device__ HSR_HULLSCAN_ATTR int hull_scan4'),('support_d','HSR_HD V3<double> support_d'),('mpr_support','HSR_HD void mpr_support'),('mpr','HSR_HD bool mpr_penetration_inl(const Geom<T>& g1, const Geom<T>& g2, GT tol, int max_iter, const Grp& g, GT& depth_,'),('box_box','HSR_HDN void box_box'),('narrow_pair','HSR_HD void narrow_pair'),('impedance','HSR_HD GT impedance'),('solver','HSR_HD int cone_zone'),('flops (+ wait at the block barrier after the solver)','HSR_HD bool model_articulated'),('forward','HSR_HDC void forward')]
core=[]
for name,pat in cm:
    for i,l in enumerate(core_src):
        if pat in l: core.append((i+1,name)); break
    else: print('core marker not found',name)
core.sort(); cst=[a for a,_ in core]
cur=None; key=None
agg=collections.defaultdict(lambda:[0,0,0,0])
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)<9 or r[0] in ('Line No','Function Name'): continue
    if r[0]!='':
        key=(cur,int(r[0])); agg[key][1]+=int(r[7]); agg[key][2]+=int(r[6]); agg[key][3]+=int(r[8])
    else: agg[key][0]+=1
tot_e=sum(v[1] for v in agg.values()); tot_s=sum(v[2] for v in agg.values())
b=collections.defaultdict(lambda:[0,0,0,0])
for (f,l),v in agg.items():
    if f=='hsr_core.h': name='core:'+(core[bisect.bisect_right(cst,l)-1][1] if l>=cst[0] else 'hdr')
    elif f=='hsrb_push.cuh': name='push:'+(push[bisect.bisect_right(pst,l)-1][1] if l>=pst[0] else 'hdr')
    else: name=f
    for k in range(4): b[name][k]+=v[k]
print('total executed',tot_e,'per warp-substep %.0f'%(tot_e/nwarpsub))
for name,v in sorted(b.items(), key=lambda kv:-kv[1][1]):
    if v[1]==0: continue
    print(f"{name:30s} static {v[0]:6d} exec/ws {v[1]/nwarpsub:8.0f} ({v[1]/tot_e:5.3f}) samples {v[2]/tot_s:6.3f} thr/inst {v[3]/max(1,v[1]):5.1f}")
