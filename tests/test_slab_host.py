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
