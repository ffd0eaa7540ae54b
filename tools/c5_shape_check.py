import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np
from stringtie_b200 import synth
from stringtie_b200.engine import Engine
t=time.time()
# few, huge bundles: long introns (Mb-scale spans), very deep
a=synth.make_batch(n_bundles=24, reads_per_bundle=400000.0, seed=5, max_intron=400000, mean_exons=16.0, expr_sigma=0.7)
print('gen',round(time.time()-t,1),'s reads',a['r_start'].size,'juncs',a['j_start'].size,'max span',int((a['b_end'].astype(np.int64)-a['b_start']).max()),'max reads',int(np.diff(a['b_read_off'].astype(np.int64)).max()), flush=True)
eng=Engine(0)
b=eng.upload(a)
for i in range(3): eng.run(b)
eng.set_profiling(True)
t=time.time(); eng.run(b); dt=time.time()-t
print('run ms',round(dt*1e3,2)); print({k:round(v,3) for k,v in eng.kernel_times()})
from tests import orc
cov,off,good,left,right=orc.run(a,threads=0)
tot=eng.fetch_cov_totals(b)
lens,_=orc.cov_layout(a)
ok=all(tot[i,1]==orc.get_cov(cov[off[i]:off[i+1]],int(lens[i]),1,0,int(lens[i])-2) for i in range(lens.size))
g2,l2,r2=eng.fetch_junction_support(b)
print('totals equal',ok,'junction support equal',bool(np.array_equal(g2,good) and np.array_equal(l2,left) and np.array_equal(r2,right)))
jp=eng.fetch_junction_pass(b); ref=orc.junction_pass(a,cov,off,good,left,right)
print('a7 equal',all(np.array_equal(jp[k],ref[k]) for k in ref))
