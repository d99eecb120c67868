"""CPU coverage of the N > 1 path's host logic: slab planes, scene partitioning and the unique-id
broadcast, run as a real world_size-2 torch.distributed job on the gloo backend."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from salva_b200 import scenes, slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_planes_split_particles_evenly_and_cover_everything():
    sc = scenes.scene_c2(24)
    h = np.float32(0.025) * np.float32(2.0) * np.float32(2.0)
    pos = sc["fluids"][0]["positions"]
    for nranks in (2, 3, 4):
        planes = slab.slab_planes(pos, h, nranks)
        assert planes[0] == slab.INT32_MIN and planes[-1] == slab.INT32_MAX and len(planes) == nranks + 1
        assert all(b - a >= 2 for a, b in zip(planes[1:-2], planes[2:-1]))
        masks = [slab.owned_mask(pos, h, planes[r], planes[r + 1]) for r in range(nranks)]
        assert np.array_equal(np.sum(masks, axis=0), np.ones(len(pos), int))      # disjoint cover
        counts = np.array([m.sum() for m in masks])
        assert counts.max() <= 1.35 * counts.mean()                                 # balanced to a cell column
        parts = [slab.partition_scene(sc, r, nranks, planes) for r in range(nranks)]
        ids = np.concatenate([p["fluids"][0]["ids"] for p in parts])
        assert np.array_equal(np.sort(ids), np.arange(len(pos)))                     # global ids survive
        for p in parts:
            assert len(p["boundaries"][0]["positions"]) == len(sc["boundaries"][0]["positions"])  # all boundaries


def test_unique_id_broadcast_and_partition_under_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from salva_b200 import scenes, slab
        dist.init_process_group("gloo")
        rank, ws = dist.get_rank(), dist.get_world_size()
        uid = slab.broadcast_unique_id(lambda: bytes(range(128)), rank)
        sc = scenes.scene_c2(16)
        part = slab.partition_scene(sc, rank, ws)
        n = len(part["fluids"][0]["positions"])
        import torch
        t = torch.tensor([n], dtype=torch.int64)
        dist.all_reduce(t)
        out = dict(rank=rank, uid_ok=uid == bytes(range(128)), n=n, total=int(t[0]), slab=part["slab"])
        json.dump(out, open(os.path.join(%r, "out%%d.json" %% rank), "w"))
        dist.destroy_process_group()
    """ % (ROOT, str(tmp_path))))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    outs = [json.load(open(tmp_path / ("out%d.json" % k))) for k in range(2)]
    assert all(o["uid_ok"] for o in outs)
    assert outs[0]["total"] == outs[1]["total"] == 16 ** 3 == outs[0]["n"] + outs[1]["n"]
    assert outs[0]["slab"][1] == outs[1]["slab"][0]


def test_rebalance_redistribution_under_gloo_world_size_2(tmp_path):
    """Plane re-balancing: the host logic — global column histogram, new planes, all-to-all of the particle
    state — as a real 2-rank gloo job: particles are conserved (ids), every particle lands on the owner of its column and the
    counts end up balanced to a cell column."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from salva_b200 import scenes, slab
        dist.init_process_group("gloo")
        rank, ws = dist.get_rank(), dist.get_world_size()
        sc = scenes.scene_c2(16)
        pos = sc["fluids"][0]["positions"]
        h = np.float32(0.1)
        cols = slab.cell_columns(pos, h)
        # a deliberately bad split: rank 0 owns 1/4 of the columns
        cut = int(cols.min()) + (int(cols.max()) + 1 - int(cols.min())) // 4
        mine = (cols < cut) if rank == 0 else (cols >= cut)
        ids = np.nonzero(mine)[0].astype(np.uint32)
        p = pos[mine]
        v = (p * 2).astype(np.float32)
        c = (p * 3).astype(np.float32)
        imb, counts = slab.imbalance(len(p))
        p2, v2, c2, i2, planes = slab.redistribute(p, v, c, ids, h)
        own = slab.owned_mask(p2, h, planes[rank], planes[rank + 1])
        out = dict(rank=rank, imb=imb, n_before=int(len(p)), n_after=int(len(p2)), all_owned=bool(own.all()),
                   payload_ok=bool(np.array_equal(v2, (p2 * 2).astype(np.float32)) and np.array_equal(c2, (p2 * 3).astype(np.float32))
                                   and np.array_equal(p2, pos[i2])), ids=i2.tolist(), planes=[int(x) for x in planes])
        json.dump(out, open(os.path.join(%r, "rb%%d.json" %% rank), "w"))
        dist.destroy_process_group()
    """ % (ROOT, str(tmp_path))))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    outs = [json.load(open(tmp_path / ("rb%d.json" % k))) for k in range(2)]
    assert outs[0]["imb"] > 0.3 and outs[0]["imb"] == outs[1]["imb"]
    assert all(o["all_owned"] and o["payload_ok"] for o in outs)
    assert sorted(outs[0]["ids"] + outs[1]["ids"]) == list(range(16 ** 3))
    assert outs[0]["planes"] == outs[1]["planes"]
    n = np.array([o["n_after"] for o in outs], float)
    assert n.max() / n.mean() - 1.0 <= 0.15


def test_planes_from_histogram_matches_slab_planes():
    sc = scenes.scene_c2(20)
    h = np.float32(0.1)
    pos = sc["fluids"][0]["positions"]
    cols = slab.cell_columns(pos, h)
    lo = int(cols.min())
    hist = np.bincount(cols - lo)
    for nranks in (2, 3, 4):
        assert slab.planes_from_histogram(hist, lo, nranks) == slab.slab_planes(pos, h, nranks)


def test_slab_scene_generates_each_ranks_slab_without_the_others():
    """scenes.slab_scene (what bench.py uses for the C4 slices): per-rank generation must reproduce the partition of the whole
    scene bit for bit — also for compressed lattices, where cell boundaries no longer fall between fixed lattice planes."""
    for comp, amp in ((1.0, 0.05), (0.9, 0.05), (0.93, 0.3)):
        full = scenes.scene_c4(32, 6, 5, compress=comp, amplitude=amp)
        h = np.float32(0.1)
        for nranks in (2, 4):
            total = 0
            for r in range(nranks):
                sc = scenes.slab_scene(lambda **kw: scenes.scene_c4(32, 6, 5, **kw), r, nranks, 32, compress=comp, amplitude=amp)
                p = sc["fluids"][0]["positions"]
                assert slab.owned_mask(p, h, sc["slab"][0], sc["slab"][1]).all()
                ref = slab.partition_scene(full, r, nranks, planes=sc["planes"])
                assert np.array_equal(ref["fluids"][0]["positions"], p) and np.array_equal(ref["fluids"][0]["ids"], sc["fluids"][0]["ids"])
                total += len(p)
            assert total == 32 * 6 * 5
