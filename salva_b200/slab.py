"""Host-side plumbing of the 1-D slab decomposition: plane selection, scene partitioning and the
NCCL unique-id broadcast over torch.distributed.  Pure numpy + torch.distributed (works with the gloo backend on CPU,
which is how tests/test_slab_host.py covers it without a GPU); the per-step exchange itself lives in the CUDA library
(salva_b200/csrc/sph_slab.inl).

The reference has no domain decomposition (one address space, unbounded hash grid: hgrid.rs:22-25); cells are the
reference's own floor(x / h) columns (hgrid.rs:41-52), so a slab is a run of consecutive cell columns.
"""
import numpy as np

INT32_MIN, INT32_MAX = -2 ** 31, 2 ** 31 - 1


def cell_columns(positions, h):
    """floor(x / h) per particle with f32 division, exactly as the engine and the reference bin particles."""
    x = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)[:, 0]
    return np.floor(x / np.float32(h)).astype(np.int64)


def slab_planes(positions, h, nranks, min_width=2):
    """Cell-column planes [p_0 = -inf, p_1, ..., p_nranks = +inf] that split the particles into nranks slabs of
    (nearly) equal particle count along x.  Every finite slab is at least `min_width` columns wide."""
    cols = cell_columns(positions, h)
    lo, hi = int(cols.min()), int(cols.max()) + 1
    if hi - lo < min_width * nranks:
        raise ValueError("domain of %d cell columns is too narrow for %d slabs" % (hi - lo, nranks))
    hist = np.bincount(cols - lo, minlength=hi - lo)
    cum = np.concatenate([[0], np.cumsum(hist)])
    planes = [INT32_MIN]
    prev = lo
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        p = lo + int(np.searchsorted(cum, target, side="left"))
        p = max(p, prev + min_width)
        p = min(p, hi - min_width * (nranks - r))
        planes.append(p)
        prev = p
    planes.append(INT32_MAX)
    return planes


def owned_mask(positions, h, cell_lo, cell_hi):
    cols = cell_columns(positions, h)
    return (cols >= cell_lo) & (cols < cell_hi)


def partition_scene(scene, rank, nranks, planes=None):
    """Scene of rank `rank`: its slab of every fluid (with global particle ids) and ALL boundary particles."""
    h = np.float32(scene["particle_radius"]) * np.float32(scene["smoothing_factor"]) * np.float32(2.0)
    allpos = np.concatenate([f["positions"] for f in scene["fluids"]])
    if planes is None:
        planes = slab_planes(allpos, h, nranks)
    lo, hi = planes[rank], planes[rank + 1]
    out = dict(scene)
    out["fluids"] = []
    base = 0
    for f in scene["fluids"]:
        m = owned_mask(f["positions"], h, lo, hi)
        g = dict(f)
        g["positions"] = np.ascontiguousarray(f["positions"][m])
        if f.get("velocities") is not None:
            g["velocities"] = np.ascontiguousarray(f["velocities"][m])
        g["ids"] = (base + np.nonzero(m)[0]).astype(np.uint32)
        base += len(f["positions"])
        out["fluids"].append(g)
    out["slab"] = (int(lo), int(hi))
    out["planes"] = planes
    return out


def broadcast_unique_id(make_id, rank, device=None):
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes (torch.distributed broadcast)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def populate_slab(world, scene, rank, nranks, unique_id, planes=None):
    """Add rank's slab of `scene` to `world` and join the decomposition.  Returns (fluid handles, boundary handles)."""
    from . import scenes
    part = scene if scene.get("partitioned") else partition_scene(scene, rank, nranks, planes)
    fh, bh = scenes.populate(world, part)
    for h, f in zip(fh, part["fluids"]):
        world.set_ids(h, f["ids"])
    lo, hi = part["slab"]
    world.init_slab(unique_id, rank, nranks, lo, hi)
    return fh, bh
