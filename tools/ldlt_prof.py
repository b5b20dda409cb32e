"""Where the block L D L^T solve spends its time: nanosecond stamps of CTA 0 at every phase boundary (ADMM_B200_LDLT_PROF=1).
usage: python tools/ldlt_prof.py [workload]      (beam_100k with LDLT, or cloth_512)"""
import os, sys
os.environ['ADMM_B200_LDLT_PROF'] = '1'
wl = sys.argv[1] if len(sys.argv) > 1 else 'cloth_512'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, numpy as np, bench
import __graft_entry__ as g
pkg = g.load_package()
args = argparse.Namespace(workload=wl, model=1, admm_iters=20, linsolver=0 if wl.startswith('beam') else 2, precision=0, floor=False, limits=False)
scene = bench.make_scene(pkg, wl)
sol = pkg.Solver(); sol.set_options(precision=0, timers=True)
bench.add_scene(sol, args, scene, pkg)
assert sol.initialize(dt=1 / 24, admm_iters=20, gravity=-9.8, linsolver=args.linsolver)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
for _ in range(2):
    sol.step_device()
info = sol.device().info(); print(info)
nl = int(info.split(' levels of which')[0].split()[-1]); cut = int(info.split('levels of which ')[1].split()[0])
p = sol.device().debug_get('ldlt_prof', 4 * nl + 8)
us = lambda a, b: (p[b] - p[a]) / 1e3
print('forward : own forest %.1f us, wait for the slowest CTA %.1f us' % (us(0, 1), us(1, 2)))
prev = 2
for lv in range(cut, nl):
    a, b = 8 + 4 * lv, 8 + 4 * lv + 1
    print('  level %2d: gather %.1f us  dense %.1f us' % (lv, us(prev, a), us(a, b))); prev = b
fwd_end = prev
for lv in range(nl - 1, cut - 1, -1):
    a, b = 8 + 4 * lv + 2, 8 + 4 * lv + 3
    print('  level %2d (backward): gather %.1f us  dense %.1f us' % (lv, us(prev, a), us(a, b))); prev = b
print('backward: own forest %.1f us, wait %.1f us' % (us(3, 4), us(4, 5)))
print('total %.1f us: forest fwd %.1f, top fwd %.1f, top bwd %.1f, forest bwd %.1f' % (us(0, 5), us(0, 2), us(2, fwd_end), us(fwd_end, 3), us(3, 5)))
print('solve ms (10 back to back):', sol.device().time_kernels(10)['global_ms'])
