"""Cycle budget of one colour pass of the tiled Gauss-Seidel kernel (profiling instantiation, ADMM_B200_GS_PROF=1).
usage: python tools/gs_prof2.py [workload] [dbg]     dbg: 2 = no halo polling, 16 = no publish (timing experiments; results wrong)"""
import sys, os
os.environ['ADMM_B200_GS_PROF'] = '1'
wl = sys.argv[1] if len(sys.argv) > 1 else 'beam_1m'
if len(sys.argv) > 2:
    os.environ['ADMM_B200_GS_DBG'] = sys.argv[2]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import __graft_entry__ as g
pkg = g.load_package()
scene = bench.make_scene(pkg, wl); mu, lam = pkg.meshes.lame(*bench.LAME)
sol = pkg.Solver(); sol.set_options(precision=0, timers=True)
sol.add_nodes(scene['verts'], scene['masses']); sol.add_tets(scene['verts'], scene['elems'], 1, mu, lam); sol.set_pins(scene['pins'])
assert sol.initialize(dt=1 / 24, admm_iters=20, gravity=-9.8, linsolver=1)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
for _ in range(3):
    sol.step_device()
ncol = len(sol.colors()); passes = 30 * ncol
parts = 148
raw = sol.device().debug_get('gs_prof', 16 * parts + 1024 + 128 * parts)
p = raw[:16 * parts].reshape(parts, 16)
names = {5: 'interior tasks', 0: 'poll + barrier', 1: 'boundary tasks + publish', 2: 'end barrier'}
print(wl, 'dbg', os.environ.get('ADMM_B200_GS_DBG', '0'), sol.device().info())
print('cycles per pass, thread 0 of each part (mean | max over parts):', {n: (int(p[:, i].mean() / passes), int(p[:, i].max() / passes)) for i, n in names.items()},
      'sum', int(sum(p[:, i].mean() for i in names) / passes), ' spins per polling thread per pass %.2f' % (p[:, 8].mean() / passes))
print('kernel phases (cycles, mean over parts): staging %.0f  r0+|b|^2 %.0f  sweeps %.0f (%.0f per pass)  total %.0f' % (p[:, 13].mean(), p[:, 14].mean(), p[:, 15].mean(), p[:, 15].mean() / passes, p[:, 3].mean()))
print('solve ms:', sol.device().time_kernels(10)['global_ms'])
