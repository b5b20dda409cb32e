#!/usr/bin/env python
"""bench.py -- ADMM iterations/s (and element-prox/s) of the B200 ADMM-elastic step.

    python bench.py [--gpus N] [--steps K] [--warmup W]           this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]     the reference's own CPU path

Default workload = the one BASELINE.json's metric is quoted on: a 1M-tet Neo-Hookean cantilever beam, 20 ADMM
iterations per step, NodalMultiColorGS global solve.  The other BASELINE configs are selected with flags:

    C2  --workload beam_100k --linsolver 0          100k-tet Neo-Hookean beam, prefactored LDLT global solve
    C3  --workload beam_1m --model 2 --floor        1M-tet StVK beam dropped on a Floor handled inside the GS sweep
    C4  --workload cloth_512 [--limits]             512x512 cloth (TriEnergyTerm), 2 corner pins (SpringPin), UzawaCG
    C5  --workload beam_8m --gpus 8                 8M-tet Neo-Hookean beam sharded over 8 GPUs

A "step" is one Solver::step() (src/Solver.cpp:35-110): `--admm-iters` (20) ADMM iterations, each = local step over
every element (prox kernel) + right-hand side assembly + global solve.  One JSON line is printed by rank 0; see
DESIGN.md "Measurement" for every field.

  value        ADMM iters/s with the state resident in HBM (Solver::step_device)
  e2e          the same through Solver::step(): x, v go host->device and back every step (pinned host)
  roofline     the dominant kernel, algorithmic bytes / CUDA-event time of that kernel
  kernels      the same for every hot kernel
  cpu_baseline the compiled reference (oracle/_ref) -- or the C oracle port -- on the same mesh on the host cores
  parity_vs_n1 (N > 1) max |x_N - x_1| against a single-GPU run of the same steps on rank 0; non-zero exit above tol
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # tets: (nx, ny, nz) unit cubes, 5 tets each (SURVEY.md 8d "Synthetic inputs")
    "beam_1m": ("tet", (320, 25, 25)),
    "beam_100k": ("tet", (100, 20, 10)),
    "beam_8m": ("tet", (640, 50, 50)),
    "beam_20k": ("tet", (40, 10, 10)),
    # triangles: mcl::factory::make_plane(n, n) (ShapeFactory.hpp:424-484), samples/sca2016/trianglestrain.cpp
    "cloth_512": ("tri", (512, 512)),
    "cloth_64": ("tri", (64, 64)),
}
LAME = (1e7, 0.399)        # admm::Lame soft rubber, samples/sca2016/beams.cpp:87
LAME_CLOTH = (100.0, 0.1)  # samples/sca2016/trianglestrain.cpp:48
LIMITS = (0.95, 1.05)      # :50-51
DENSITY = 1522.0           # samples/utils/AddMeshes.hpp:104-106
MODEL_NAMES = {0: "linear", 1: "neohookean", 2: "stvk"}
SOLVER_NAMES = {0: "LDLT", 1: "NodalMultiColorGS 30 sweeps omega=1.9", 2: "UzawaCG (pins as SpringPin energy terms, LDLT inside)"}
# N ranks vs one GPU after the same steps (same colours; only the summation order of a vertex's element shares and the
# fp32 rounding inside the sweeps differ): max |dx| / bounding-box diagonal.  fp32 has 6e-8; W + 2K + 2 steps x 20 ADMM
# iterations amplify it (measured at N = 2 after 25 steps: 4.7e-7); the fp32 gate of SURVEY.md 8d is 1e-4.
WEAK_TIMEOUT_S = 420       # watchdog of the optional 8M-tet leg at N = 8 (normally ~100 s incl. building the mesh)
TIMER_STRIDE = 4           # per-phase / per-kernel CUDA events on every 4th step of a timed region (see measure(): timed)
PARITY_TOL_REL = 2e-6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="beam_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--model", type=int, default=1, help="tets: 1 = NeoHookean (headline), 2 = StVK, 0 = linear")
    ap.add_argument("--admm-iters", type=int, default=20)
    ap.add_argument("--linsolver", type=int, default=None, help="1 = NodalMultiColorGS (tets default), 0 = LDLT, 2 = UzawaCG (cloth default)")
    ap.add_argument("--precision", type=int, default=0, help="element data: 0 = fp32 (production), 1 = fp64")
    ap.add_argument("--floor", action="store_true", help="BASELINE config 3 style: no pins, the beam drops on a Floor handled inside the GS sweep")
    ap.add_argument("--limits", action="store_true", help="cloth: strain limits 0.95 / 1.05 (trianglestrain.cpp:50-51)")
    ap.add_argument("--coloring", default="greedy", choices=["greedy", "random"], help="greedy = largest-degree-first (4 colours on the beams); random = randomised palette, the reference's colour count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the comparison with a single-GPU run")
    ap.add_argument("--no-weak", action="store_true", help="--gpus 8 on beam_1m: skip the extra weak-scaling measurement (8M tets)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=270.0, help="budget of the whole --impl reference run")
    a = ap.parse_args()
    if a.linsolver is None:
        a.linsolver = 2 if WORKLOADS[a.workload][0] == "tri" else 1
    return a


def load_package():
    import __graft_entry__ as g
    return g.load_package()


def make_scene(pkg, workload):
    kind, dims = WORKLOADS[workload]
    if kind == "tri":
        verts, tris = pkg.meshes.make_plane_sym(*dims)
        masses = pkg.meshes.lumped_masses_tris(verts, tris, 1.0).astype(np.float64)   # AddMeshes.hpp:189: weighted_masses(masses, 1.f)
        v64 = verts.astype(np.float64)
        return dict(kind="tri", verts=v64, elems=tris, masses=masses, pins=pkg.meshes.cloth_corner_pins(verts), x0=v64.copy(), dims=dims, floor_y=None)
    nx, ny, nz = dims
    verts, tets = pkg.meshes.make_tet_blocks(nx, ny, nz)
    masses = pkg.meshes.lumped_masses_tets(verts, tets, DENSITY).astype(np.float64)
    v64 = verts.astype(np.float64)
    pins = np.nonzero(v64[:, 0] < v64[:, 0].min() + 1e-2)[0].astype(np.int32)
    # start from a smoothly bent beam: the first two steps of an undeformed mesh hit the reference's
    # rest-state line-search pathology which would make its CPU arm take hours
    x0 = v64.copy()
    L = x0[:, 0].max() - x0[:, 0].min()
    s = (x0[:, 0] - x0[:, 0].min()) / L
    x0[:, 1] -= 0.08 * L * s * s
    x0[:, 2] += 0.02 * L * np.sin(3.0 * s)
    return dict(kind="tet", verts=v64, elems=tets, masses=masses, pins=pins, x0=x0, dims=dims, floor_y=float(x0[:, 1].min() - 0.05))


def add_scene(solver, args, scene, pkg):
    """The same calls for admm_b200::Solver and for the CPU checkers."""
    solver.add_nodes(scene["verts"], scene["masses"])
    if scene["kind"] == "tri":
        mu, lam = pkg.meshes.lame(*LAME_CLOTH)
        lim = LIMITS if args.limits else (-100.0, 100.0)
        solver.add_tris(scene["verts"], scene["elems"], mu, lam, lim[0], lim[1])
        solver.set_pins(scene["pins"])
        return
    mu, lam = pkg.meshes.lame(*LAME)
    solver.add_tets(scene["verts"], scene["elems"], args.model, mu, lam)
    if args.floor:
        solver.add_floor(scene["floor_y"])
    else:
        solver.set_pins(scene["pins"])


def workload_name(args, scene):
    if scene["kind"] == "tri":
        return ("%d-triangle cloth sheet (make_plane %dx%d, %d vertices), TriEnergyTerm%s, 2 corner pins, %d ADMM iters/step, %s, dt=1/24 s, g=-9.8"
                % (len(scene["elems"]), scene["dims"][0], scene["dims"][1], len(scene["verts"]), " with strain limits 0.95/1.05" if args.limits else "",
                   args.admm_iters, SOLVER_NAMES[args.linsolver]))
    nx, ny, nz = scene["dims"]
    return ("%d-tet %s %s (%dx%dx%d cubes x 5 tets), %d ADMM iters/step, %s, dt=1/24 s, g=-9.8"
            % (len(scene["elems"]), MODEL_NAMES.get(args.model, str(args.model)),
               "beam dropped on a Floor (no pins)" if args.floor else "cantilever beam", nx, ny, nz, args.admm_iters, SOLVER_NAMES[args.linsolver]))


def common_config(args, scene):
    """Identical in both arms' JSON lines (the driver compares them)."""
    esz = 4 if args.precision == 0 else 8
    per_elem = (16 + 19 * esz + 16 * 4 + 4) if scene["kind"] == "tet" else (16 + 11 * esz + 12 * 4)
    return {"workload": workload_name(args, scene), "n_elements": int(len(scene["elems"])), "n_verts": int(len(scene["verts"])),
            "admm_iters_per_step": args.admm_iters, "linsolver": args.linsolver,
            "l2": "no explicit flush: one ADMM iteration streams ~%.0f MB of element data (> 126 MB L2) between reuses" % (len(scene["elems"]) * per_elem / 1e6)
                  if len(scene["elems"]) * per_elem > 126e6 else
                  "no explicit flush; the %.0f MB of element data one ADMM iteration streams are below the 126 MB L2, so part of it can stay L2-resident between iterations (not the headline workload)" % (len(scene["elems"]) * per_elem / 1e6)}


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9 or not c[0].isdigit() or int(c[0]) != self.idx:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# -------------------------------------------------------------------------------------------------
# CPU arms: the reference's own implementation (oracle/_ref) or the C oracle port
# -------------------------------------------------------------------------------------------------
def cpu_solver(args, scene, pkg, admm_iters):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import checkers
    kind = "ref" if checkers.have_ref() else "oracle"
    if kind == "oracle" and not os.path.exists(checkers.ORACLE_PATH):
        import __graft_entry__ as g
        g.build()
    s = checkers.CpuSolver(kind)
    if kind == "ref" and os.environ.get("TORCHELASTIC_RUN_ID") and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun imposes OMP_NUM_THREADS=1 on its workers; the reference arm is meant to use all host cores
        try:
            checkers.ref_lib().ref_set_omp_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass  # an older oracle/_ref build without the setter
    add_scene(s, args, scene, pkg)
    if kind == "oracle" and args.linsolver == 1:
        raise RuntimeError("the oracle port takes its colours from a caller; build oracle/_ref for the CPU arm")
    t0 = time.time()
    assert s.initialize(dt=1.0 / 24, admm_iters=admm_iters, gravity=-9.8, linsolver=args.linsolver)
    init_s = time.time() - t0
    s.set_x(scene["x0"].ravel())
    threads = (checkers.ref_lib().ref_omp_threads() if kind == "ref" else checkers.oracle_lib().oracle_omp_threads())
    n_colors = len(s.get_colors()) if (kind == "ref" and args.linsolver == 1) else 0
    return s, ("reference" if kind == "ref" else "port"), int(threads), init_s, n_colors


def coloring_note(kind, n_colors, linsolver):
    if linsolver != 1:
        return "global solve: the reference's Eigen SimplicialLDLT (AMD ordering)"
    if kind == "reference":
        return "NodalMultiColorGS with the reference's own randomised colouring (graphcolor::color_matrix): %d colours" % n_colors
    return "NodalMultiColorGS"


def cpu_baseline(args, scene, pkg):
    """Rank 0, N=1: a bounded sample (about args.cpu_seconds of CPU work) of the same workload."""
    s, kind, threads, init_s, n_colors = cpu_solver(args, scene, pkg, 1)
    t0 = time.time()
    s.step()                       # 1 ADMM iteration: warm-up + calibration
    t_iter = time.time() - t0
    n_it = int(max(2, min(args.admm_iters, args.cpu_seconds / max(t_iter, 1e-3))))
    s._f("set_admm_iters")(s.h, n_it)
    t0 = time.time()
    s.step()                       # one Solver::step() of n_it ADMM iterations
    dt = time.time() - t0
    rd = s.runtime_data()
    loc, glob = rd["local_ms"], rd["global_ms"]
    n_el = len(scene["elems"])
    out = {
        "value": n_it / dt, "unit": "ADMM iters/s", "cores": threads, "kind": kind,
        "sample": "one Solver::step() of %d ADMM iterations (of the workload's %d) on the same %d-element mesh after a 1-iteration warm-up step, %.1f s; initialize() %.1f s not timed; %s"
                  % (n_it, args.admm_iters, n_el, dt, init_s, coloring_note(kind, n_colors, args.linsolver)),
        "elem_prox_per_s": n_el * n_it / (loc * 1e-3) if loc > 0 else None,
        "local_ms_per_iter": loc / n_it, "global_ms_per_iter": glob / n_it, "n_colors": n_colors,
    }
    s.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg_meshes = load_meshes_only()
    scene = make_scene(pkg_meshes, args.workload)
    K, W = args.steps, max(args.warmup, 0)
    try:
        s, kind, threads, init_s, n_colors = cpu_solver(args, scene, pkg_meshes, 1)
    except Exception as e:  # the checker libraries are missing: nothing to time
        print(json.dumps({"impl": "reference", "unavailable": "no CPU reference library: %s" % e}))
        return
    t0 = time.time()
    s.step()
    t_iter = time.time() - t0
    # each "step" is one Solver::step() of n_it ADMM iterations: the workload's full count unless the whole run would
    # not fit the budget (--ref-seconds), then a bounded sample of them
    n_it = int(max(1, min(args.admm_iters, args.ref_seconds / ((K + W) * max(t_iter, 1e-3)))))
    s._f("set_admm_iters")(s.h, n_it)
    for _ in range(W):
        s.step()
    loc = glob = 0.0
    t0 = time.time()
    for _ in range(K):
        s.step()
        rd = s.runtime_data()
        loc += rd["local_ms"]
        glob += rd["global_ms"]
    dt = time.time() - t0
    n_el = len(scene["elems"])
    value = K * n_it / dt
    sample = ("each step = one Solver::step() of %d ADMM iterations (%s) on the same %d-element mesh; %d threads; %s"
              % (n_it, "the workload's full count" if n_it == args.admm_iters else "a bounded sample of the workload's %d per step" % args.admm_iters,
                 n_el, threads, coloring_note(kind, n_colors, args.linsolver)))
    line = {
        "impl": "reference", "metric": "admm_iters_per_s", "value": value, "unit": "ADMM iters/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "strong",  # ONE mesh of fixed size whatever N
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args, scene),
        "cpu_baseline": {"value": value, "unit": "ADMM iters/s", "cores": threads, "kind": kind, "sample": sample, "sampled_iters_per_step": n_it,
                         "elem_prox_per_s": n_el * K * n_it / (loc * 1e-3) if loc > 0 else None,
                         "local_ms_per_iter": loc / (K * n_it), "global_ms_per_iter": glob / (K * n_it), "init_s": init_s, "n_colors": n_colors},
        "e2e": {"value": value, "unit": "ADMM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def load_meshes_only():
    """The mesh generator without the native libraries (the reference arm must not load our .so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("admm_b200_meshes", os.path.join(ROOT, "admm-elastic_b200", "meshes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class P(object):
        meshes = mod
    return P


# -------------------------------------------------------------------------------------------------
# the B200 arm
# -------------------------------------------------------------------------------------------------
def build_solver(pkg, args, scene, local_rank, stream, rank=0, world=1):
    sol = pkg.Solver()
    sol.set_options(device=local_rank, precision=args.precision, coloring=pkg.COLOR_RANDOM if args.coloring == "random" else pkg.COLOR_GREEDY,
                    timers=True, stream=stream.cuda_stream)
    if world > 1:
        sol.set_rank(rank, world)
    add_scene(sol, args, scene, pkg)
    t0 = time.time()
    assert sol.initialize(dt=1.0 / 24, admm_iters=args.admm_iters, gravity=-9.8, linsolver=args.linsolver)
    return sol, time.time() - t0


def measure(pkg, args, scene, torch, dist, stream, rank, world, local_rank, K, W, clocks=None):
    """W warm-up + K timed device-resident steps, then 2 + K timed steps through Solver::step() with host buffers.
    Returns the numbers of both timed regions and the final positions."""
    sol, init_s = build_solver(pkg, args, scene, local_rank, stream, rank, world)
    n_el, n_verts = len(scene["elems"]), len(scene["verts"])
    n_el_rank, n_verts_rank, owner = n_el, n_verts, None
    n_owned, n_ghost = n_verts, 0
    if world > 1:
        def all_gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        sol.mgpu_connect(all_gather_bytes)
        owner = sol.node_owner()
        n_el_rank = int((owner[scene["elems"]] == rank).any(axis=1).sum())   # cut elements are computed on both sides
        n_verts_rank = int((owner == rank).sum())
        n_owned, n_ghost = sol.mgpu_nodes()
        assert n_owned == n_verts_rank
    sol.set_x(scene["x0"].ravel())
    dev = sol.device()
    rp, _, _ = sol.system_matrix()
    nnz_L = int(rp[-1]) - n_verts  # off-diagonal entries of the scalar matrix
    n_colors = len(sol.colors()) if args.linsolver == 1 else 0

    def barrier():
        # This rank's solver work must be FINISHED before the NCCL kernel of dist.barrier() is launched: the solve kernel is
        # a cooperative launch that needs every SM whole (512 threads x 128 registers per CTA) and waits for its
        # neighbours on other GPUs; an NCCL kernel that slips in between two queued solver kernels on one rank and
        # between two others on the next one holds an SM on both, neither rank's next solve kernel can be scheduled, and
        # the NCCL kernels wait for each other's ranks forever.  (Seen once, at N = 8 in round 2, when the sampled timers let
        # the host finish queueing the steps long before the GPU had run them.)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn bracketed by barrier+sync, CUDA events on the solver's stream, max over ranks.  The per-phase and
        per-kernel events are recorded inside this region -- on every TIMER_STRIDE-th step, an event costs ~3 us of stream
        time and a fully instrumented step carries 0.64 ms of them (tools/timer_overhead.py) -- and read only after it
        (deferred timers, admm_b200_collect_timers): no host synchronise per step, the next step's launches queue behind
        the running one.  acc holds SUMS over the sampled steps, acc["steps"] of them."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sol.set_timers(False)
        dev.set_deferred_timers(True, stride=TIMER_STRIDE)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(n):
                fn()
            e1.record(stream)
        barrier()
        acc = dev.collect_timers()
        acc["kernels"] = {k: list(v) for k, v in dev.kernel_times().items()}   # events tightly around each hot kernel launch
        dev.set_deferred_timers(False)
        sol.set_timers(True)
        assert acc["steps"] == (n + TIMER_STRIDE - 1) // TIMER_STRIDE, acc
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc

    # ---- resident path: state stays in HBM -------------------------------------------------------
    sol.upload_state()  # x0 -> device
    for _ in range(W):
        sol.step_device()
    if clocks is not None:
        clocks.start()
    l0 = dev.launch_count()
    ms_res, acc = timed(sol.step_device, K)
    launches = dev.launch_count() - l0
    sol.sync_state()
    # ---- end to end: Solver::step() with host buffers every step ---------------------------------
    for _ in range(2):
        sol.step()
    ms_e2e, _ = timed(sol.step, K)
    clk = clocks.stop() if clocks is not None else None
    x_final = sol.get_x()
    out = dict(ms_res=ms_res, ms_e2e=ms_e2e, acc=acc, launches=launches, clk=clk, x=x_final, owner=owner, init_s=init_s, info=dev.info(),
               nnz_L=nnz_L, n_colors=n_colors, n_el_rank=n_el_rank, n_verts_rank=n_verts_rank, h2d=2 * 3 * 8 * (n_owned + n_ghost), d2h=2 * 3 * 8 * (n_owned + n_ghost))
    sol.close()
    return out


def merged_positions(torch, dist, m, rank, world):
    """Every rank's OWNED nodes merged into one array (all-reduce of the masked positions)."""
    x = m["x"].reshape(-1, 3).copy()
    if world == 1:
        return x
    x[m["owner"] != rank] = 0.0
    t = torch.from_numpy(x).cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = load_package()
    scene = make_scene(pkg, args.workload)
    n_el, n_verts = len(scene["elems"]), len(scene["verts"])
    K, W = args.steps, max(args.warmup, 3)
    iters = args.admm_iters
    stream = torch.cuda.Stream()
    if world > 1 and args.linsolver != 1:
        raise SystemExit("bench.py: only the NodalMultiColorGS solve shards over several GPUs (DESIGN.md 7); "
                         "LDLT / UzawaCG configurations are single-GPU (SURVEY.md 8e: replicas only)")

    m = measure(pkg, args, scene, torch, dist, stream, rank, world, local_rank, K, W, ClockSampler(local_rank) if rank == 0 else None)
    acc, ms_res, ms_e2e = m["acc"], m["ms_res"], m["ms_e2e"]
    finite = bool(np.isfinite(m["x"]).all())

    # ---- N > 1: the same steps on ONE GPU (rank 0), positions of every rank's owned nodes against it ----
    parity = None
    if world > 1 and not args.no_parity:
        x_n = merged_positions(torch, dist, m, rank, world)
        if rank == 0:
            if n_el <= 2000000:
                one = measure(pkg, args, scene, torch, dist, stream, 0, 1, local_rank, K, W)
                diff = float(np.abs(one["x"].reshape(-1, 3) - x_n).max())
                bbox = float(np.linalg.norm(scene["verts"].max(0) - scene["verts"].min(0)))
                parity = {"max_abs": diff, "max_over_bbox": diff / bbox, "tol_over_bbox": PARITY_TOL_REL, "bbox_diagonal_m": bbox, "unit": "m", "steps_compared": W + 2 * K + 2,
                          "what": "owned nodes of all %d ranks merged vs a single-GPU run of the same %d steps on rank 0 (same colours)" % (world, W + 2 * K + 2),
                          "single_gpu_value": iters * K / (one["ms_res"] * 1e-3)}
            else:
                parity = {"max_abs": None, "max_over_bbox": None, "tol_over_bbox": PARITY_TOL_REL, "skipped": "the single-GPU run of this mesh is not part of the default bench (see tests/test_multi_gpu.py)"}
        dist.barrier()

    # N > 1: ONE mesh sharded over the ranks (strong scaling) -- the job's ADMM iterations, not a sum
    value = iters * K / (ms_res * 1e-3)
    e2e = iters * K / (ms_e2e * 1e-3)
    # bytes per step of Solver::step(): x and v of every rank's owned + ghost nodes, up and down, summed over the ranks
    h2d_bytes, d2h_bytes = m["h2d"], m["d2h"]
    if world > 1:
        t = torch.tensor([float(h2d_bytes), float(d2h_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d_bytes, d2h_bytes = int(t[0].item()), int(t[1].item())

    # ---- roofline: algorithmic bytes / CUDA-event durations from the timed region ----
    peak, peak_src = measured_peaks()
    n_sampled = max(int(acc["steps"]), 1)   # steps of the timed region that carried events
    n_launch = n_sampled * iters
    # average launch duration of each hot kernel: CUDA events recorded on the solver's stream right before and after
    # every launch inside the timed region (admm_b200_kernel_times).  The step_breakdown phases below also contain
    # the helpers (scratch memset, the queue consumer of degenerate elements) and the gaps between launches.
    kt = acc["kernels"]

    def avg(k, fallback):
        ms_k, n_k = kt.get(k, (0.0, 0))
        return ms_k / n_k * 1e-3 if n_k else fallback
    t_local = avg("tet_local_kernel", acc["local_ms"] / n_launch * 1e-3)
    t_asm = avg("assemble_kernel", acc["assemble_ms"] / n_launch * 1e-3)
    n_solves = max(kt.get("solve_kernel", (0.0, 0))[1], 1)
    t_glob = avg("solve_kernel", (acc["global_ms"] - acc["assemble_ms"]) / n_launch * 1e-3)
    esz = 4 if args.precision == 0 else 8
    n_el_rank, n_verts_rank, nnz_L = m["n_el_rank"], m["n_verts_rank"], m["nnz_L"]

    def roof(b, t, note):
        a = b / t / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                "ms_per_launch": t * 1e3, "algorithmic_bytes_per_launch": b, "peak_source": peak_src, "note": note}

    kernels = {}
    if scene["kind"] == "tet":
        # per launch = per rank: this rank's elements / nodes
        local_name = "tet_local_kernel"
        kernels[local_name] = roof(n_el_rank * (16 + 9 * esz + 9 * esz + 4 * 3 * esz + 9 * esz + 9 * esz), t_local,
                                   "208 B/tet-prox (fp32): idx 16 + Dm^-1 36 + u in/out 72 + x gather 48 + z 36")
        kernels["assemble_kernel"] = roof(n_el_rank * (9 * esz + 9 * esz + 16) + n_verts_rank * 24, t_asm, "88 B/tet + 24 B/vertex")
    else:
        local_name = "tri_local_kernel"
        kernels[local_name] = roof(n_el_rank * (16 + 4 * esz + 6 * esz + 6 * esz + 3 * 3 * esz + 6 * esz), t_local,
                                   "140 B/tri-prox (fp32): idx 16 + rest pose 16 + u in/out 48 + x gather 36 + z 24")
        kernels["assemble_kernel"] = roof(n_el_rank * (6 * esz + 4 * esz + 16) + n_verts_rank * 24, t_asm, "56 B/tri + 24 B/vertex")
    info = m["info"]
    if args.linsolver == 1:
        kernels["mcgs_kernel"] = roof(30 * (20 * nnz_L // world + 36 * n_verts_rank), t_glob,
                                      "30 sweeps x (20 B x nnz(L) + 36 B x n_verts) = what a streaming sweep would move (SURVEY 8d); one persistent launch per ADMM "
                                      "iteration keeps matrix and iterate in shared memory, so the figure can exceed the HBM peak: the kernel's real limits are shared-memory "
                                      "wavefronts and the inter-SM latency of the halo exchange (DESIGN.md 4.1), see 'traffic' for what it actually reads from DRAM")
        kernels["mcgs_kernel"]["actual_limit"] = "shared-memory gather wavefronts + inter-SM latency of the dependent halo exchanges (n_colours x 30 per solve)"
    else:
        # prefactored L D L^T solve: every factor entry once forward and once backward (value 8 + column 4), the right-hand side,
        # the work vector and the result as double4 (SURVEY.md 8d "LDLT path")
        nnz_f = int(info.split("nnz(L) ")[1].split(",")[0]) if "nnz(L) " in info else 0
        kernels["ldlt_solve_kernel"] = roof(2 * nnz_f * 12 + n_verts * (32 * 4 + 8 + 4), t_glob,
                                            "2 x nnz(L_factor) x 12 B + 140 B/vertex per solve (3 right-hand sides fused); level-scheduled: bound by the number of dependency levels "
                                            "x grid-barrier latency, not by HBM")
        kernels["ldlt_solve_kernel"]["solves_per_admm_iter"] = n_solves / float(n_launch)
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * (n_solves / float(n_launch) if k == "ldlt_solve_kernel" else 1.0))
    tr = load_traffic()
    for k in kernels:
        key = k if k != "ldlt_solve_kernel" else "%s@%s" % (k, args.workload)
        if key in tr and world == 1 and (args.workload in ("beam_1m", "cloth_512") or k == "ldlt_solve_kernel"):
            kernels[k]["traffic"] = tr[key]   # dram bytes per launch from the committed ncu --set full capture of this workload
    roofline = dict(kernels[dominant], kernel=dominant)

    cfg = common_config(args, scene)
    line = {
        "metric": "admm_iters_per_s", "value": value, "unit": "ADMM iters/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,  # one mesh of fixed size sharded over the ranks: total work does not grow with N
        "dtype": "f32 elements + f64 nodes/solve" if args.precision == 0 else "f64", "data": "synthetic",
        "config": cfg,
        "details": {"nnz_L_offdiag": nnz_L, "n_colors": m["n_colors"], "coloring": args.coloring,
                    "multi_gpu": ("one mesh sharded by node ownership over %d ranks; cut elements computed on both sides; neighbour values and solved cut positions pushed into peer memory by the solve kernel (CUDA IPC over NVLink), no NCCL call in the data path" % world) if world > 1 else "single GPU",
                    "n_elements_this_rank": n_el_rank, "n_verts_this_rank": n_verts_rank, "init_s": m["init_s"], "global_solve_kernel": info},
        "elem_prox_per_s": n_el / (acc["local_ms"] / n_launch * 1e-3),  # whole local phase (kernel + helpers), all ranks' elements
        "e2e": {"value": e2e, "unit": "ADMM iters/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / K, "api": "admm_b200::Solver::step() -> admm_b200_step_host"},
        "gpu_launches": int(m["launches"]),
        "roofline": roofline, "kernels": kernels,
        "step_breakdown_ms": {"local": acc["local_ms"] / n_sampled, "assemble": acc["assemble_ms"] / n_sampled, "solve": (acc["global_ms"] - acc["assemble_ms"]) / n_sampled,
                              "device_step": acc["step_ms"] / n_sampled},
        "timer_sampling": {"stride": TIMER_STRIDE, "sampled_steps": n_sampled, "of_steps": K,
                           "note": "step_breakdown_ms, roofline and kernels come from CUDA events recorded inside the timed region on every %d-th step; an event costs ~3 us of stream time, so an instrumented step is ~0.6 ms longer than an uninstrumented one (tools/timer_overhead.py)" % TIMER_STRIDE},
        "clocks": m["clk"], "finite": finite,
    }
    if scene["kind"] == "tet":
        line["tet_prox_per_s"] = line["elem_prox_per_s"]
    if parity is not None:
        line["parity_vs_n1"] = parity

    # ---- --gpus 8 on the headline mesh: the weak-scaling companion (8M tets, 1M per GPU) as an extra key ----
    if world == 8 and args.workload == "beam_1m" and not args.no_weak:
        wargs = argparse.Namespace(**vars(args))
        wargs.workload = "beam_8m"
        # Safety net: this extra leg must never cost the main line.  If it has not finished after WEAK_TIMEOUT_S (a cross-GPU
        # wait that never ends cannot be interrupted from inside the process), rank 0 prints the line without it and every
        # rank leaves; the driver's clock then still sees a complete run.
        import threading

        def give_up():
            if rank == 0:
                line["weak_8m"] = {"value": None, "error": "the 8M-tet leg did not finish within %d s and was abandoned" % WEAK_TIMEOUT_S}
                print(json.dumps(line), flush=True)
            os._exit(0)
        watchdog = threading.Timer(WEAK_TIMEOUT_S, give_up)
        watchdog.daemon = True
        watchdog.start()
        wscene = make_scene(pkg, "beam_8m")
        wm = measure(pkg, wargs, wscene, torch, dist, stream, rank, world, local_rank, min(K, 5), 3)
        watchdog.cancel()
        line["weak_8m"] = {"value": iters * min(K, 5) / (wm["ms_res"] * 1e-3), "unit": "ADMM iters/s", "e2e": iters * min(K, 5) / (wm["ms_e2e"] * 1e-3),
                           "workload": workload_name(wargs, wscene), "n_elements_this_rank": wm["n_el_rank"], "steps": min(K, 5),
                           "vs_single_gpu_1m": iters * min(K, 5) / (wm["ms_res"] * 1e-3) / (parity["single_gpu_value"] if parity and parity.get("single_gpu_value") else float("nan")),
                           "finite": bool(np.isfinite(wm["x"]).all())}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args, scene, pkg)
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "error": str(e)}
    if rank == 0:
        print(json.dumps(line))
    bad_parity = bool(parity and parity.get("max_over_bbox") is not None and not (parity["max_over_bbox"] <= PARITY_TOL_REL))
    if world > 1:
        flag = torch.tensor([1.0 if bad_parity else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        bad_parity = bool(flag.item() > 0)
        dist.barrier()
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench.py: non-finite positions")
    if bad_parity:
        raise SystemExit("bench.py: the %d-GPU positions differ from the single-GPU run by more than %g x the bounding-box diagonal" % (world, PARITY_TOL_REL))


def load_traffic():
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
