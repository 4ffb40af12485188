import sys; sys.path[:0]=['.','t-route_b200','tests']
import numpy as np
import test_gpu_api as T
from oracle import oracle as o
import helpers as H
from troute_b200 import synth
from troute_b200.routing.fast_reach import mc_reach
c=T._reference_style_case()
ref=T._call(o.compute_network_structured,c)[1]
got=T._call(mc_reach.compute_network_structured,c)[1]
lv=synth.levels_from_down(c['down'])
def report(name,a):
    bad=(a.view(np.int32)!=ref.view(np.int32))
    rows=np.nonzero(bad.any(1))[0]
    print(name,'bad values',bad.sum(),'bad rows',rows.size,'min level of bad rows',lv[rows].min() if rows.size else None)
    if rows.size:
        r=rows[np.argmin(lv[rows])]; col=np.nonzero(bad[r])[0][0]
        print('  first bad row',r,'level',lv[r],'is_lp',r in set(c['lp_rows'].tolist()),'col',col,a[r,col],ref[r,col])
report('api',got)
for zero in (False,True):
    case=dict(n=c['n'],down=c['down'],params=np.nan_to_num(c['params']) if zero else synth.channel_params(c['down'],seed=7),cols=list(c['cols']),qlat=c['qlat'],q0=c['q0'],nsteps=c['nsteps'],qts=c['qts'])
    case['up_ptr'],case['up_rows']=synth.upstream_csr(c['down'])
    kind=np.zeros(c['n'],np.uint8); kind[c['lp_rows']]=1
    case['kind']=kind; case['lp_rows']=c['lp_rows'].astype(np.int64); case['wbody']=c['wbody']
    for mode in (0,1,2):
        out,_,_=H.engine_route(case,False,mode=mode)
        report(f'engine zero={zero} mode={mode}',out)
