"""What do the per-phase / per-kernel CUDA events of bench.py's timed region cost?  The same K resident steps of the headline
scene with the deferred timers on and off, each timed by two events around the whole run.
usage: python tools/timer_overhead.py [workload]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import __graft_entry__ as g
pkg = g.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else 'beam_1m'
scene = bench.make_scene(pkg, wl); mu, lam = pkg.meshes.lame(*bench.LAME)
stream = torch.cuda.Stream()
sol = pkg.Solver(); sol.set_options(precision=0, timers=False, stream=stream.cuda_stream)
sol.add_nodes(scene['verts'], scene['masses']); sol.add_tets(scene['verts'], scene['elems'], 1, mu, lam); sol.set_pins(scene['pins'])
assert sol.initialize(dt=1 / 24, admm_iters=20, gravity=-9.8, linsolver=1)
sol.set_x(scene['x0'].ravel()); sol.upload_state()
dev = sol.device()
K = 10
for mode in ('off', 'deferred', 'off', 'deferred'):
    sol.set_timers(False)
    dev.set_deferred_timers(mode == 'deferred')
    for _ in range(3): sol.step_device()
    if mode == 'deferred': dev.collect_timers()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(K): sol.step_device()
        e1.record(stream)
    torch.cuda.synchronize()
    if mode == 'deferred': dev.collect_timers()
    dev.set_deferred_timers(False)
    ms = e0.elapsed_time(e1) / K
    print('timers %-8s: %.4f ms per step, %.1f ADMM iters/s' % (mode, ms, 20e3 / ms))
