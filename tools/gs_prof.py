import sys, os
os.environ['ADMM_B200_GS_PROF']='1'
sys.path.insert(0,'/root/repo')
import numpy as np, bench, argparse
import __graft_entry__ as g
pkg=g.load_package()
args=argparse.Namespace(workload='beam_1m',model=1,admm_iters=20,linsolver=1,precision=0)
scene=bench.make_scene(pkg,'beam_1m'); mu,lam=pkg.meshes.lame(*bench.LAME)
sol=pkg.Solver(); sol.set_options(precision=0,timers=True)
sol.add_nodes(scene['verts'],scene['masses']); sol.add_tets(scene['verts'],scene['tets'],1,mu,lam); sol.set_pins(scene['pins'])
assert sol.initialize(dt=1/24,admm_iters=20,gravity=-9.8,linsolver=1)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
for _ in range(3): sol.step_device()
print(sol.runtime_data())
ncol=len(sol.colors()); passes=30*ncol
p=sol.device().debug_get('gs_prof',16*148).reshape(148,16)
names=['wait','boundary','publish','total','interior']
print('colours',ncol,'passes',passes,' cycles per pass (mean over parts | max):', {n:(int(p[:,i].mean()/passes), int(p[:,i].max()/passes)) for i,n in enumerate(names)})
for nm,o in (('boundary warp0',5),('interior warp0',9)):
    n=p[:,o+3].sum()
    if n > 0: print(nm,'slices/pass %.2f'%(p[:,o+3].mean()/passes),'cycles per slice: meta %.0f gather %.0f tail %.0f'%(p[:,o].sum()/n,p[:,o+1].sum()/n,p[:,o+2].sum()/n))
print('kernel phases (cycles, mean over parts): staging %.0f  r0+|b|^2 %.0f  sweeps %.0f  total %.0f'%(p[:,13].mean(),p[:,14].mean(),p[:,15].mean(),p[:,3].mean()))
print(sol.device().info())
