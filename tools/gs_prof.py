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
p=sol.device().debug_get('gs_prof',4*148).reshape(148,4)
print('cycles per solve: wait mean/max %.0f %.0f | compute mean/max %.0f %.0f | publish mean/max %.0f %.0f | total %.0f'%(p[:,0].mean(),p[:,0].max(),p[:,1].mean(),p[:,1].max(),p[:,2].mean(),p[:,2].max(),p[:,3].mean()))
print('per pass (270): wait %.0f compute %.0f publish %.0f total %.0f cycles'%tuple(p.mean(0)/270))
print(sol.device().info())
