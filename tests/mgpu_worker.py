"""Worker of the multi-rank tests: launched once per rank by tests/test_multi_gpu.py (and usable with
torchrun).  mode "plan": CPU only (gloo), checks that the ranks' exchange plans fit together.
mode "step": one GPU per rank, steps a sharded beam and compares with a single-GPU run of the same scene."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    import scenes
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = g.load_package()

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    if mode == "plan":
        import scipy.sparse as sp
        verts, tets = pkg.meshes.make_tet_blocks(24, 5, 4)
        n = len(verts)
        rows, cols = np.repeat(tets, 4, axis=1).ravel(), np.tile(tets, (1, 4)).ravel()
        A = sp.csr_matrix((np.random.RandomState(0).rand(rows.size) + 0.1, (rows, cols)), shape=(n, n))
        A = (A + A.T).tocsr()
        A.sort_indices()
        colors = pkg.color_matrix(A.indptr, A.indices, A.data, 0)
        off = np.zeros(len(colors) + 1, np.int32)
        off[1:] = np.cumsum([len(c) for c in colors])
        nodes = np.concatenate(colors).astype(np.int32)
        rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
        pos = np.ascontiguousarray(verts.astype(np.float64))
        ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        sms = 8  # a small "SM count" so that a 3k-node mesh has non-trivial parts
        mask, ghost, owner = np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        rc = pkg.cuda_lib.admm_b200_mgpu_plan_check(n, ip(rp), ip(ci), dp(va), len(colors), ip(off), ip(nodes), dp(pos), sms, world, rank,
                                                    mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), ip(ghost), ip(owner))
        assert rc == 0, pkg.cuda_lib.admm_b200_last_error(None)
        part = np.zeros(n, np.int32)
        assert pkg.cuda_lib.admm_b200_plan_parts(n, ip(rp), ip(ci), dp(va), dp(pos), sms * world, ip(part)) == 0
        assert (part // sms == owner).all()                       # plan_parts is the partition finalize uses
        allr = all_gather((mask, ghost, owner))
        for r in range(world):
            assert (allr[r][2] == owner).all()                     # every rank computes the same ownership
        coo = A.tocoo()
        cut = owner[coo.row] != owner[coo.col]
        for a in range(world):
            ghost_a = allr[a][1]
            # what rank a reads from others == neighbours (in A) of its nodes that it does not own
            expect = np.zeros(n, np.int32)
            sel = cut & (owner[coo.row] == a)
            expect[coo.col[sel]] = 1
            assert (ghost_a == expect).all()
            for b in range(world):
                if a == b:
                    continue
                mask_b = allr[b][0]
                reads_from_b = (ghost_a == 1) & (owner == b)
                pushes_to_a = ((mask_b >> a) & 1).astype(bool) & (owner == b)
                assert (reads_from_b == pushes_to_a).all()         # every ghost is pushed, nothing else is
        # element sharding: an element is kept by the ranks owning one of its nodes; every element is kept
        kept = np.zeros(len(tets), np.int32)
        mine = (owner[tets] == rank).any(axis=1)
        counts = all_gather(int(mine.sum()))
        t = torch.from_numpy(mine.astype(np.int32))
        dist.all_reduce(t)
        assert (t.numpy() >= 1).all()
        dup = float((t.numpy() > 1).mean())
        if rank == 0:
            print(json.dumps({"ok": True, "elements_per_rank": counts, "duplicated_fraction": dup, "ghosts": [int(a[1].sum()) for a in allr]}))
    elif mode == "step":
        # variants: beam (pinned cantilever), floor (StVK beam dropped on a Floor handled inside the sweep, no pins),
        # cloth (triangles + Gauss-Seidel pins): each sharded over the ranks and compared with a single-GPU run
        torch.cuda.set_device(rank)
        variant = sys.argv[2] if len(sys.argv) >= 3 else "beam"
        if variant == "cloth":
            scene = scenes.cloth(pkg.meshes, 40)
            x0 = scene[0].ravel().copy()
        else:
            scene = scenes.beam(pkg.meshes, 48, 6, 6)
            x0 = scenes.bend(scene[0]).ravel()
        floor_y = float(scene[0][:, 1].min() - 0.03)

        def build(r, w):
            s = pkg.Solver()
            s.set_options(device=rank, precision=pkg.FP32, timers=False)
            if w > 1:
                s.set_rank(r, w)
            if variant == "cloth":
                s.add_nodes(scene[0], scene[2])
                s.add_tris(scene[0], scene[1], 100.0 / 2.2, 100.0 * 0.1 / (1.1 * 0.8), 0.95, 1.05)
                s.set_pins(scene[3])
                assert s.initialize(dt=1.0 / 24, admm_iters=8, gravity=-9.8, linsolver=1)
            elif variant == "floor":
                scenes.build_tet_scene(s, scene, 2, linsolver=1, iters=8, floor=floor_y, pin=False)
            else:
                scenes.build_tet_scene(s, scene, 1, linsolver=1, iters=8)
            return s

        n_steps = 6 if variant == "floor" else 3
        s = build(rank, world)
        s.mgpu_connect(all_gather)
        s.set_x(x0)
        for _ in range(n_steps):
            s.step()
        owner = s.node_owner()
        n_owned, n_ghost = s.mgpu_nodes()
        assert n_owned == int((owner == rank).sum()) and 0 < n_ghost < len(owner)
        x = s.get_x().reshape(-1, 3)
        # step() moves only this rank's nodes (owned + ghost): most of what it does not own still holds the start values
        untouched = (x[owner != rank] == x0.reshape(-1, 3)[owner != rank]).all(axis=1)
        assert int((~untouched).sum()) <= n_ghost and untouched.sum() > 0
        xm = torch.from_numpy(np.where((owner == rank)[:, None], x, 0.0))
        dist.all_reduce(xm)
        if rank == 0:
            ref = build(0, 1)
            ref.set_x(x0)
            for _ in range(n_steps):
                ref.step()
            xr = ref.get_x().reshape(-1, 3)
            err = float(np.abs(xr - xm.numpy()).max())
            landed = bool(variant != "floor" or np.abs(xr[:, 1] - floor_y).min() < 1e-9)
            print(json.dumps({"ok": bool(err < 2e-6 and landed), "err": err, "variant": variant, "landed": landed, "n_nodes": int(len(owner)),
                              "owned": [int((owner == r).sum()) for r in range(world)], "info": s.device().info()}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
