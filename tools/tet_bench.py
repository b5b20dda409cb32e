"""Times the local step (tet prox kernel + its fix-up launch) alone on the 1M-tet beam: `time_kernels`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import __graft_entry__ as g
pkg = g.load_package()
model = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = bench.make_scene(pkg, 'beam_1m'); mu, lam = pkg.meshes.lame(*bench.LAME)
sol = pkg.Solver(); sol.set_options(precision=0, timers=True)
sol.add_nodes(scene['verts'], scene['masses']); sol.add_tets(scene['verts'], scene['elems'], model, mu, lam); sol.set_pins(scene['pins'])
assert sol.initialize(dt=1/24, admm_iters=20, gravity=-9.8, linsolver=1)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
for _ in range(3): sol.step_device()
t = sol.device().time_kernels(20)
n = len(scene['elems'])
print('model', model, 'minblocks', os.environ.get('ADMM_B200_TET_MINBLOCKS', 'default'), {k: round(v * 1e3, 2) for k, v in t.items()}, 'us;  local: %.2f G tet-prox/s, %.1f%% of 6540 GB/s at 208 B' % (n / t['local_ms'] / 1e6, 100 * 208 * n / (t['local_ms'] * 1e-3) / 6539.9e9))
