#!/usr/bin/env python
"""bench.py -- ADMM iterations/s (and tet-prox/s) of the B200 ADMM-elastic step on the workload
BASELINE.json's metric is quoted on: a 1M-tet Neo-Hookean cantilever beam.

    python bench.py [--gpus N] [--steps K] [--warmup W]           this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]     the reference's own CPU path

A "step" is one Solver::step() (src/Solver.cpp:35-110): `--admm-iters` (20) ADMM iterations, each =
local step over every tet (prox kernel) + right-hand side assembly + NodalMultiColorGS solve (30
sweeps).  One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every field.

  value   ADMM iters/s with the state resident in HBM (Solver::step_device)
  e2e     the same through Solver::step(): x, v go host->device and back every step (pinned host)
  roofline   the dominant kernel (mcgs_kernel), algorithmic bytes / CUDA-event time
  kernels    the same for the tet prox kernel and the assembly kernel
  cpu_baseline   the compiled reference (oracle/_ref) -- or the C oracle port -- on the same mesh on
                 the host cores, a bounded sample of ADMM iterations
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
    # name: (nx, ny, nz) unit cubes, 5 tets each (SURVEY.md 8d "Synthetic inputs")
    "beam_1m": (320, 25, 25),
    "beam_100k": (100, 20, 10),
    "beam_8m": (640, 50, 50),
    "beam_20k": (40, 10, 10),
}
LAME = (1e7, 0.399)     # admm::Lame soft rubber, samples/sca2016/beams.cpp:87
DENSITY = 1522.0        # samples/utils/AddMeshes.hpp:104-106
MODEL_NAMES = {0: "linear", 1: "neohookean", 2: "stvk"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="beam_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--model", type=int, default=1, help="1 = NeoHookean (headline), 2 = StVK, 0 = linear")
    ap.add_argument("--admm-iters", type=int, default=20)
    ap.add_argument("--linsolver", type=int, default=1, help="1 = NodalMultiColorGS (headline), 0 = LDLT")
    ap.add_argument("--precision", type=int, default=0, help="element data: 0 = fp32 (production), 1 = fp64")
    ap.add_argument("--floor", action="store_true", help="BASELINE config 3 style: no pins, the beam drops on a Floor handled inside the GS sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=120.0, help="budget of the whole --impl reference run")
    return ap.parse_args()


def load_package():
    import __graft_entry__ as g
    return g.load_package()


def make_scene(pkg, workload):
    nx, ny, nz = WORKLOADS[workload]
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
    return dict(verts=v64, tets=tets, masses=masses, pins=pins, x0=x0, dims=(nx, ny, nz), floor_y=float(v64[:, 1].min() - 0.05))


def workload_name(args, scene):
    nx, ny, nz = scene["dims"]
    return ("%d-tet %s %s (%dx%dx%d cubes x 5 tets), %d ADMM iters/step, %s, dt=1/24 s, g=-9.8"
            % (len(scene["tets"]), MODEL_NAMES.get(args.model, str(args.model)),
               "beam dropped on a Floor (no pins)" if getattr(args, "floor", False) else "cantilever beam", nx, ny, nz, args.admm_iters,
               "NodalMultiColorGS 30 sweeps omega=1.9" if args.linsolver == 1 else "LDLT"))


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
    mu, lam = pkg.meshes.lame(*LAME)
    s.add_nodes(scene["verts"], scene["masses"])
    s.add_tets(scene["verts"], scene["tets"], args.model, mu, lam)
    if getattr(args, "floor", False):
        s.add_floor(scene["floor_y"])
    else:
        s.set_pins(scene["pins"])
    if kind == "oracle" and args.linsolver == 1:
        raise RuntimeError("the oracle port takes its colours from a caller; build oracle/_ref for the CPU arm")
    t0 = time.time()
    assert s.initialize(dt=1.0 / 24, admm_iters=admm_iters, gravity=-9.8, linsolver=args.linsolver)
    init_s = time.time() - t0
    s.set_x(scene["x0"].ravel())
    threads = (checkers.ref_lib().ref_omp_threads() if kind == "ref" else checkers.oracle_lib().oracle_omp_threads())
    return s, ("reference" if kind == "ref" else "port"), int(threads), init_s


def cpu_baseline(args, scene, pkg):
    """Rank 0, N=1: a bounded sample (about args.cpu_seconds of CPU work) of the same workload."""
    s, kind, threads, init_s = cpu_solver(args, scene, pkg, 1)
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
    n_tets = len(scene["tets"])
    out = {
        "value": n_it / dt, "unit": "ADMM iters/s", "cores": threads, "kind": kind,
        "sample": "one Solver::step() of %d ADMM iterations (of the workload's 20) on the same %d-tet mesh after a 1-iteration warm-up step, %.1f s; initialize() %.1f s not timed"
                  % (n_it, n_tets, dt, init_s),
        "tet_prox_per_s": n_tets * n_it / (loc * 1e-3) if loc > 0 else None,
        "local_ms_per_iter": loc / n_it, "global_ms_per_iter": glob / n_it,
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
        s, kind, threads, init_s = cpu_solver(args, scene, pkg_meshes, 1)
    except Exception as e:  # the checker libraries are missing: nothing to time
        print(json.dumps({"impl": "reference", "unavailable": "no CPU reference library: %s" % e}))
        return
    t0 = time.time()
    s.step()
    t_iter = time.time() - t0
    # each "step" is a bounded sample: one Solver::step() of n_it (<= admm_iters) ADMM iterations
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
    n_tets = len(scene["tets"])
    value = K * n_it / dt
    sample = ("each step = one Solver::step() of %d ADMM iterations (of the workload's %d per step) on the same %d-tet mesh; %d threads"
              % (n_it, args.admm_iters, n_tets, threads))
    line = {
        "impl": "reference", "metric": "admm_iters_per_s", "value": value, "unit": "ADMM iters/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "n_tets": n_tets, "n_verts": len(scene["verts"]),
                   "admm_iters_per_step": args.admm_iters, "sampled_iters_per_step": n_it},
        "cpu_baseline": {"value": value, "unit": "ADMM iters/s", "cores": threads, "kind": kind, "sample": sample,
                         "tet_prox_per_s": n_tets * K * n_it / (loc * 1e-3) if loc > 0 else None,
                         "local_ms_per_iter": loc / (K * n_it), "global_ms_per_iter": glob / (K * n_it), "init_s": init_s},
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
    n_tets, n_verts = len(scene["tets"]), len(scene["verts"])
    K, W = args.steps, max(args.warmup, 3)
    iters = args.admm_iters
    mu, lam = pkg.meshes.lame(*LAME)

    stream = torch.cuda.Stream()
    sol = pkg.Solver()
    sol.set_options(device=local_rank, precision=args.precision, coloring=pkg.COLOR_GREEDY, timers=True, stream=stream.cuda_stream)
    if world > 1:
        sol.set_rank(rank, world)
    sol.add_nodes(scene["verts"], scene["masses"])
    sol.add_tets(scene["verts"], scene["tets"], args.model, mu, lam)
    if args.floor:
        sol.add_floor(scene["floor_y"])
    else:
        sol.set_pins(scene["pins"])
    t0 = time.time()
    assert sol.initialize(dt=1.0 / 24, admm_iters=iters, gravity=-9.8, linsolver=args.linsolver)
    init_s = time.time() - t0
    n_tets_rank, n_verts_rank = n_tets, n_verts
    if world > 1:
        def all_gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        sol.mgpu_connect(all_gather_bytes)
        owner = sol.node_owner()
        n_tets_rank = int((owner[scene["tets"]] == rank).any(axis=1).sum())   # cut elements are computed on both sides
        n_verts_rank = int((owner == rank).sum())
    sol.set_x(scene["x0"].ravel())
    dev = sol.device()
    rp, _, _ = sol.system_matrix()
    nnz_L = int(rp[-1]) - n_verts  # off-diagonal entries of the scalar matrix
    n_colors = len(sol.colors()) if args.linsolver == 1 else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn bracketed by barrier+sync, CUDA events on the solver's stream, max over ranks.  The per-phase and
        per-kernel events of every step are recorded inside this region but read only after it (deferred timers,
        admm_b200_collect_timers): no host synchronise per step, the next step's launches queue behind the running one."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sol.set_timers(False)
        dev.set_deferred_timers(True)
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
        assert acc["steps"] == n, acc
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
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = dev.launch_count()
    ms_res, acc = timed(sol.step_device, K)
    launches = dev.launch_count() - l0
    sol.sync_state()
    # ---- end to end: Solver::step() with host buffers every step ---------------------------------
    for _ in range(2):
        sol.step()
    ms_e2e, _ = timed(sol.step, K)
    clk = clocks.stop() if rank == 0 else None
    x_final = sol.get_x()
    finite = bool(np.isfinite(x_final).all())

    # N > 1: ONE mesh sharded over the ranks (strong scaling) -- the job's ADMM iterations, not a sum
    value = iters * K / (ms_res * 1e-3)
    e2e = iters * K / (ms_e2e * 1e-3)
    state_bytes = 2 * 3 * n_verts * 8

    # ---- roofline: algorithmic bytes / CUDA-event durations from the timed region ----
    peak, peak_src = measured_peaks()
    n_launch = K * iters
    # average launch duration of each hot kernel: CUDA events recorded on the solver's stream right before and after
    # every launch inside the timed region (admm_b200_kernel_times).  The step_breakdown phases below also contain
    # the helpers (scratch memset, the queue consumer of degenerate elements) and the gaps between launches.
    kt = acc["kernels"]
    def avg(k, fallback):
        ms_k, n_k = kt.get(k, (0.0, 0))
        return ms_k / n_k * 1e-3 if n_k else fallback
    t_local = avg("tet_local_kernel", acc["local_ms"] / n_launch * 1e-3)
    t_asm = avg("assemble_kernel", acc["assemble_ms"] / n_launch * 1e-3)
    t_glob = avg("solve_kernel", (acc["global_ms"] - acc["assemble_ms"]) / n_launch * 1e-3)
    esz = 4 if args.precision == 0 else 8
    # per launch = per rank: this rank's elements / nodes
    bytes_prox = n_tets_rank * (16 + 9 * esz + 9 * esz + 4 * 3 * esz + 9 * esz + 9 * esz)     # 208 B/tet in fp32
    bytes_asm = n_tets_rank * (9 * esz + 9 * esz + 16) + n_verts_rank * 24                     # 88 B/tet + 24 B/vertex
    sweeps = 30
    bytes_gs = sweeps * (20 * nnz_L // world + 36 * n_verts_rank) if args.linsolver == 1 else None

    def roof(b, t, note):
        a = b / t / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                "ms_per_launch": t * 1e3, "algorithmic_bytes_per_launch": b, "peak_source": peak_src, "note": note}

    kernels = {
        "tet_local_kernel": roof(bytes_prox, t_local, "208 B/tet-prox (fp32): idx 16 + Dm^-1 36 + u in/out 72 + x gather 48 + z 36"),
        "assemble_kernel": roof(bytes_asm, t_asm, "88 B/tet + 24 B/vertex"),
    }
    if bytes_gs:
        kernels["mcgs_kernel"] = roof(bytes_gs, t_glob, "30 sweeps x (20 B x nnz(L) + 36 B x n_verts) = what a streaming sweep would move (SURVEY 8d); one persistent launch per ADMM "
                                      "iteration keeps matrix and iterate in shared memory, so the figure can exceed the HBM peak: the kernel's real limits are shared-memory "
                                      "wavefronts and the inter-SM latency of the halo exchange (DESIGN.md 4.1), see 'traffic' for what it actually reads from DRAM")
        kernels["mcgs_kernel"]["actual_limit"] = "shared-memory gather wavefronts (~2100 cycles per colour pass) + inter-SM latency of 120 dependent halo exchanges per solve"
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"])
    tr = load_traffic()
    for k in kernels:
        if k in tr:
            kernels[k]["traffic"] = tr[k]
    roofline = dict(kernels[dominant], kernel=dominant)

    line = {
        "metric": "admm_iters_per_s", "value": value, "unit": "ADMM iters/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32 elements + f64 nodes/solve" if args.precision == 0 else "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "n_tets": n_tets, "n_verts": n_verts, "nnz_L_offdiag": nnz_L,
                   "n_colors": n_colors, "admm_iters_per_step": iters,
                   "multi_gpu": ("one mesh sharded by node ownership over %d ranks; cut elements computed on both sides; neighbour values and solved cut positions pushed into peer memory by the solve kernel (CUDA IPC over NVLink), no NCCL call in the data path" % world) if world > 1 else "single GPU",
                   "n_tets_this_rank": n_tets_rank, "n_verts_this_rank": n_verts_rank,
                   "l2": "no explicit flush: one ADMM iteration streams ~%.0f MB of element data (> 126 MB L2) between reuses" % (n_tets * (16 + 19 * esz + 16 * 4 + 4) / 1e6),
                   "init_s": init_s, "global_solve_kernel": dev.info()},
        "tet_prox_per_s": n_tets / (acc["local_ms"] / n_launch * 1e-3),  # whole local phase (kernel + helpers), all ranks' tets
        "e2e": {"value": e2e, "unit": "ADMM iters/s", "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                "ms_per_step": ms_e2e / K, "api": "admm_b200::Solver::step() -> admm_b200_step_host"},
        "gpu_launches": int(launches),
        "roofline": roofline, "kernels": kernels,
        "step_breakdown_ms": {"local": acc["local_ms"] / K, "assemble": acc["assemble_ms"] / K, "solve": (acc["global_ms"] - acc["assemble_ms"]) / K,
                              "device_step": acc["step_ms"] / K},
        "clocks": clk, "finite": finite,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args, scene, pkg)
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "error": str(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench.py: non-finite positions")


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
