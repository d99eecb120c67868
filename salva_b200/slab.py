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


# ---- plane re-balancing ("re-balanced when imbalance > ~5 %") -------------------------------------------
def planes_from_histogram(hist, lo, nranks, min_width=2):
    """Planes that split a per-cell-column particle histogram (column lo + k -> hist[k]) into nranks slabs of nearly equal
    count, every slab at least min_width columns wide (same rule as slab_planes)."""
    hist = np.asarray(hist, np.int64)
    hi = lo + len(hist)
    if hi - lo < min_width * nranks:
        raise ValueError("domain of %d cell columns is too narrow for %d slabs" % (hi - lo, nranks))
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


def _exchange_rows(chunks, width, dtype, group=None):
    """All-to-all of variable-length row blocks: chunks[d] (k_d x width) goes to rank d; returns the received blocks in
    rank order.  all_to_all_single on NCCL (device tensors), pairwise isend/irecv elsewhere (gloo has no all_to_all)."""
    import torch
    import torch.distributed as dist
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.tensor([len(c) for c in chunks], dtype=torch.int64)
    use_nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_nccl else torch.device("cpu")
    incoming = torch.zeros(ws, dtype=torch.int64, device=dev)
    cdev = counts.to(dev)
    if use_nccl:
        dist.all_to_all_single(incoming, cdev, group=group)
    else:
        gathered = [torch.zeros(ws, dtype=torch.int64) for _ in range(ws)]
        dist.all_gather(gathered, counts, group=group)
        incoming = torch.stack([g[rank] for g in gathered])
    incoming = [int(x) for x in incoming.cpu()]
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint32): torch.int32}[np.dtype(dtype)]
    send = [torch.from_numpy(np.ascontiguousarray(c, dtype).reshape(-1, width).view(np.int32 if dtype == np.uint32 else dtype)) for c in chunks]
    if use_nccl:
        flat = torch.cat(send).to(dev) if sum(len(c) for c in chunks) else torch.zeros((0, width), dtype=tdt, device=dev)
        out = torch.empty((sum(incoming), width), dtype=tdt, device=dev)
        dist.all_to_all_single(out, flat, output_split_sizes=incoming, input_split_sizes=[len(c) for c in chunks], group=group)
        out = out.cpu().numpy()
        recv, o = [], 0
        for k in incoming:
            recv.append(out[o:o + k])
            o += k
    else:
        recv = [None] * ws
        recv[rank] = send[rank].numpy()
        bufs, reqs = {}, []
        for d in range(ws):
            if d == rank:
                continue
            bufs[d] = torch.empty((incoming[d], width), dtype=tdt)
            if incoming[d]:
                reqs.append(dist.irecv(bufs[d], src=d, group=group))
            if len(send[d]):
                reqs.append(dist.isend(send[d].contiguous(), dst=d, group=group))
        for q in reqs:
            q.wait()
        for d in bufs:
            recv[d] = bufs[d].numpy()
    return [np.ascontiguousarray(r).view(dtype).reshape(-1, width) for r in recv]


def imbalance(n_local, group=None):
    """max / mean - 1 of the per-rank particle counts."""
    import torch
    import torch.distributed as dist
    use_nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_nccl else torch.device("cpu")
    ws = dist.get_world_size(group)
    t = torch.zeros(ws, dtype=torch.int64, device=dev)
    t[dist.get_rank(group)] = int(n_local)
    dist.all_reduce(t, group=group)
    c = t.cpu().numpy().astype(np.float64)
    return float(c.max() / max(c.mean(), 1.0) - 1.0), c.astype(np.int64)


def redistribute(pos, vel, vc, ids, h, group=None, min_width=2):
    """Pure host logic of a re-balance (also runs under gloo): global cell-column histogram -> new planes -> every particle goes
    to the rank that owns its column.  Returns (pos, vel, vc, ids, planes) of this rank after the exchange."""
    import torch
    import torch.distributed as dist
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    use_nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_nccl else torch.device("cpu")
    cols = cell_columns(pos, h) if len(pos) else np.zeros(0, np.int64)
    ext = torch.tensor([-(int(cols.min()) if len(cols) else 2 ** 40), int(cols.max()) if len(cols) else -2 ** 40], dtype=torch.int64, device=dev)
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    lo, hi = -int(ext[0]), int(ext[1]) + 1
    hist = torch.from_numpy(np.bincount(cols - lo, minlength=hi - lo).astype(np.int64)).to(dev)
    dist.all_reduce(hist, group=group)
    planes = planes_from_histogram(hist.cpu().numpy(), lo, ws, min_width)
    dest = np.searchsorted(np.asarray(planes[1:-1], np.int64), cols, side="right")
    order = np.argsort(dest, kind="stable")
    bounds = np.searchsorted(dest[order], np.arange(ws + 1))
    sel = [order[bounds[d]:bounds[d + 1]] for d in range(ws)]
    out = []
    for arr, width, dt in ((pos, 3, np.float32), (vel, 3, np.float32), (vc, 3, np.float32), (ids, 1, np.uint32)):
        a = np.ascontiguousarray(arr, dt).reshape(-1, width)
        out.append(np.concatenate(_exchange_rows([a[s] for s in sel], width, dt, group)))
    return out[0], out[1], out[2], out[3].reshape(-1), planes


def rebalance(world, fluid, threshold=0.05, group=None):
    """Re-balance the slab planes of a running slab world when max/mean - 1 of the per-rank particle counts exceeds
    `threshold`: positions, velocities, velocity_changes and ids of every particle move to the rank that owns its cell
    column under the new planes (collective: every rank calls it at the same step).  Returns the imbalance before, and
    whether a re-balance happened."""
    imb, _ = imbalance(world.num_particles(fluid), group)
    if imb <= threshold:
        return imb, False
    import torch.distributed as dist
    rank = dist.get_rank(group)
    pos, vel = world.read_fluid(fluid)
    vc = world.debug(fluid, "velocity_change")
    ids = world.read_ids(fluid)
    pos, vel, vc, ids, planes = redistribute(pos, vel, vc, ids, world.h, group)
    world.replace_particles(fluid, pos, vel, vc, ids)
    world.set_slab(planes[rank], planes[rank + 1])
    return imb, True
