import sys, os
os.environ['ADMM_B200_GS_PROF']='1'
sys.path.insert(0,'/root/repo')
import numpy as np, bench, argparse
import __graft_entry__ as g
pkg=g.load_package()
args=argparse.Namespace(workload='beam_1m',model=1,admm_iters=20,linsolver=1,precision=0)
WL=sys.argv[1] if len(sys.argv)>1 else 'beam_1m'
scene=bench.make_scene(pkg,WL); mu,lam=pkg.meshes.lame(*bench.LAME)
sol=pkg.Solver(); sol.set_options(precision=0,timers=True)
sol.add_nodes(scene['verts'],scene['masses']); sol.add_tets(scene['verts'],scene['elems'],1,mu,lam); sol.set_pins(scene['pins'])
assert sol.initialize(dt=1/24,admm_iters=20,gravity=-9.8,linsolver=1)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
for _ in range(3): sol.step_device()
print(sol.runtime_data())
ncol=len(sol.colors()); passes=30*ncol
raw=sol.device().debug_get('gs_prof',16*148+1024+128*148)
p=raw[:16*148].reshape(148,16); tr=raw[16*148:16*148+1024].reshape(32,4,8).astype(np.int64)
hn=np.maximum(p[:,12],1.0); ok=p[:,12]>0
if ok.any():
    print('global timer, ns, mean over parts (and the slowest part): halo complete after the LATEST neighbour published %.0f (%.0f), after the EARLIEST %.0f, after this part\'s own previous publish %.0f'
          % ((p[ok,10]/hn[ok]).mean(), (p[ok,10]/hn[ok]).max(), (p[ok,4]/hn[ok]).mean(), (p[ok,11]/hn[ok]).mean()))
names=['wait (poll + halo barrier)','slice compute','end barrier','total','-','between passes']
print('colours',ncol,'passes',passes,' cycles per pass (mean over parts | max):', {n:(int(p[:,i].mean()/passes), int(p[:,i].max()/passes)) for i,n in enumerate(names)})
print('kernel phases (cycles, mean over parts): staging %.0f  r0+|b|^2 %.0f  sweeps %.0f  total %.0f'%(p[:,13].mean(),p[:,14].mean(),p[:,15].mean(),p[:,3].mean()))
print('dbg',os.environ.get('ADMM_B200_GS_DBG','0'),sol.device().info())
print('bwarps info: n_own/halo etc. in info above; solve us:', sol.device().time_kernels(10))

# timeline of passes 40..43 of part (dbg >> 8): per warp, cycles relative to warp 0's pass-40 start
# events: 0 pass start, 1 poll done, 2 after halo barrier, 3 slice done + published, 4 after end-of-pass barrier, 5 gather done, 6 update done (before publishing)
t00=tr[0,0,0]
if t00>0:
    for ps in range(4):
        print('pass',40+ps)
        for w in range(16):
            e=tr[w,ps]
            print('  warp %2d: '%w+' '.join('%6s'%(int(v-t00) if v>0 else '-') for v in e[:7]))
