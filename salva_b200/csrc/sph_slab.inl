// sph_slab.inl — multi-GPU 1-D slab decomposition along x.  Included by sph_engine.cu.
//
// One process per GPU.  Rank r owns the particles whose cell x-coordinate floor(x/h) lies in [lo_r, hi_r); every
// step it receives the neighbours' boundary cell columns as GHOST particles (one cell = one kernel radius wide).
// The x-major sort keeps a whole yz-plane of cells contiguous, so after the sort the arrays are laid out as
//     [ left ghosts | my left boundary column | interior | my right boundary column | right ghosts ]
// and every per-iteration exchange (rho, kappa, v*, normals) is an ncclSend/ncclRecv of a contiguous array range
// with NO pack kernel.  Sender and receiver agree on the order inside a column because both run the same stable
// counting sort over the same particles in the same arrival order (deterministic mode is forced on).
// Only the first exchange of a step (positions/velocities of the boundary columns, before the sort) and the
// migration of particles that crossed a plane need a compaction.
//
// NCCL is bound at run time (dlopen) so the library has no link-time dependency on a particular libnccl.
#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int NCCL_CHAR = 0, NCCL_FLOAT = 7, NCCL_UINT64 = 5, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

bool nccl_load(std::string* err) {
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        *err = std::string("cannot load libnccl: ") + dlerror();
        return false;
    }
    NcclApi a;
    a.lib = h;
#define NSYM(field, name)                                          \
    *(void**)(&a.field) = dlsym(h, name);                          \
    if (!a.field) {                                                \
        *err = std::string("libnccl lacks symbol ") + name;        \
        return false;                                              \
    }
    NSYM(GetUniqueId, "ncclGetUniqueId")
    NSYM(CommInitRank, "ncclCommInitRank")
    NSYM(CommDestroy, "ncclCommDestroy")
    NSYM(Send, "ncclSend")
    NSYM(Recv, "ncclRecv")
    NSYM(AllReduce, "ncclAllReduce")
    NSYM(GroupStart, "ncclGroupStart")
    NSYM(GroupEnd, "ncclGroupEnd")
    NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
    g_nccl = a;
    return true;
}

#define NC(call)                                                                                                    \
    do {                                                                                                            \
        int r_ = (call);                                                                                            \
        if (r_ != 0) return w->fail(SPH_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
    } while (0)

void slab_release(sph_world* w) {
    SlabState& S = w->slab;
    if (S.comm && S.own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(S.comm);
    S.comm = nullptr;
    S.d_cnt.release(); S.flag.release(); S.flag2.release(); S.flag_o.release(); S.gid_l.release(); S.gid_r.release(); S.d_cnt64.release();
    for (int a = 0; a < 3; ++a) {
        S.out_l[a].release();
        S.out_r[a].release();
    }
    S.active = false;
}

inline bool slab_on(const sph_world* w) { return w->slab.active; }
inline int slab_left(const sph_world* w) { return w->slab.rank > 0 ? w->slab.rank - 1 : -1; }
inline int slab_right(const sph_world* w) { return w->slab.rank + 1 < w->slab.nranks ? w->slab.rank + 1 : -1; }

// Exchange element counts with both neighbours: out[0] = what the left rank sends me, out[1] = from the right.
sph_status slab_exchange_counts(sph_world* w, uint32_t to_left, uint32_t to_right, uint32_t* from_left, uint32_t* from_right) {
    SlabState& S = w->slab;
    uint32_t h[4] = {to_left, to_right, 0, 0};
    CU(cudaMemcpyAsync(S.d_cnt.p, h, sizeof h, cudaMemcpyHostToDevice, w->st));
    NC(g_nccl.GroupStart());
    if (slab_left(w) >= 0) {
        NC(g_nccl.Send(S.d_cnt.p + 0, 4, NCCL_CHAR, slab_left(w), S.comm, w->st));
        NC(g_nccl.Recv(S.d_cnt.p + 2, 4, NCCL_CHAR, slab_left(w), S.comm, w->st));
    }
    if (slab_right(w) >= 0) {
        NC(g_nccl.Send(S.d_cnt.p + 1, 4, NCCL_CHAR, slab_right(w), S.comm, w->st));
        NC(g_nccl.Recv(S.d_cnt.p + 3, 4, NCCL_CHAR, slab_right(w), S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    CU(cudaMemcpyAsync(h, S.d_cnt.p, sizeof h, cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    *from_left = slab_left(w) >= 0 ? h[2] : 0;
    *from_right = slab_right(w) >= 0 ? h[3] : 0;
    return SPH_OK;
}

// Per-iteration ghost refresh of one per-particle array (elem = bytes per particle): my boundary columns go to the
// neighbours, their boundary columns land in my ghost ranges.  Contiguous ranges, no packing.
sph_status slab_refresh(sph_world* w, void* array, size_t elem) {
    if (!slab_on(w)) return SPH_OK;
    SlabState& S = w->slab;
    char* a = static_cast<char*>(array);
    NC(g_nccl.GroupStart());
    if (slab_left(w) >= 0) {
        if (S.sl_count) NC(g_nccl.Send(a + (size_t)S.sl_begin * elem, (size_t)S.sl_count * elem, NCCL_CHAR, slab_left(w), S.comm, w->st));
        if (S.gl_count) NC(g_nccl.Recv(a, (size_t)S.gl_count * elem, NCCL_CHAR, slab_left(w), S.comm, w->st));
    }
    if (slab_right(w) >= 0) {
        if (S.sr_count) NC(g_nccl.Send(a + (size_t)S.sr_begin * elem, (size_t)S.sr_count * elem, NCCL_CHAR, slab_right(w), S.comm, w->st));
        if (S.gr_count) NC(g_nccl.Recv(a + (size_t)S.gr_begin * elem, (size_t)S.gr_count * elem, NCCL_CHAR, slab_right(w), S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    w->stats_exchanges++;
    return SPH_OK;
}

// Sum a small device float buffer over all ranks (error means of the Jacobi loops: dfsph_solver.rs:153-158).
sph_status slab_allreduce(sph_world* w, float* buf, size_t n) {
    if (!slab_on(w)) return SPH_OK;
    NC(g_nccl.AllReduce(buf, buf, n, NCCL_FLOAT, NCCL_SUM, w->slab.comm, w->st));
    return SPH_OK;
}

// Step prologue in slab mode:
//   1. drop last step's ghosts and hand particles that left [lo, hi) to the neighbour that now owns them (migration),
//   2. send the new boundary columns (pos, vel, vc) to the neighbours and append theirs as ghosts.
// On exit the arrays hold [owned (N) | ghosts] in arbitrary order; the counting sort follows.
sph_status slab_begin_step(sph_world* w) {
    SlabState& S = w->slab;
    if (w->fluids.size() != 1) return w->fail(SPH_ERR_INVALID, "slab decomposition supports one fluid per world");
    if (w->desc.solver != SPH_SOLVER_DFSPH || w->tile) return w->fail(SPH_ERR_INVALID, "slab decomposition supports DFSPH with gather_backend 0");
    for (auto& fr : w->fluids[0].forces)
        if (fr.d.kind == SPH_FORCE_BECKER2009_ELASTICITY) return w->fail(SPH_ERR_INVALID, "Becker2009 elasticity is not slab-decomposed");
    const uint32_t n_slots = (uint32_t)w->Ntot;  // layout of the previous step: owned range + ghosts (or owned only)
    const uint32_t ob = w->own_begin, on = (uint32_t)w->N;
    int c = w->cur;
    // ---- 1. classify the owned particles by their CURRENT cell column -------------------------------------------
    CU(S.flag.ensure(3 * (size_t)n_slots + 8));
    CU(S.flag_o.ensure((size_t)on + 8));
    uint32_t* fk = S.flag.p;                // keep
    uint32_t* fl = S.flag.p + n_slots;      // leaves to the left
    uint32_t* fr = S.flag.p + 2 * (size_t)n_slots;  // leaves to the right
    CU(cudaMemsetAsync(S.flag.p, 0, 3 * (size_t)n_slots * sizeof(uint32_t), w->st));
    LAUNCH(k_slab_classify, on, 256, w->pos[c].p, w->orig[c].p, ob, on, n_slots, S.lo, S.hi, S.has_left, S.has_right, fk, fl, fr, S.flag_o.p);
    // counts = last flag + last scan value; scans in place (exclusive)
    TRY(scan_exclusive(w, S.flag_o.p, on));   // new original index of kept particles (stable in old original order)
    uint32_t tail[6];
    // read the last flags before scanning
    if (n_slots) {
        CU(cudaMemcpyAsync(&tail[0], fk + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&tail[1], fl + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&tail[2], fr + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
    }
    // keep an unscanned copy of the flags for the scatter
    CU(S.flag2.ensure(3 * (size_t)n_slots + 8));
    CU(cudaMemcpyAsync(S.flag2.p, S.flag.p, 3 * (size_t)n_slots * sizeof(uint32_t), cudaMemcpyDeviceToDevice, w->st));
    TRY(scan_exclusive(w, fk, n_slots));
    TRY(scan_exclusive(w, fl, n_slots));
    TRY(scan_exclusive(w, fr, n_slots));
    uint32_t nk = 0, nl = 0, nr = 0;
    if (n_slots) {
        CU(cudaMemcpyAsync(&tail[3], fk + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&tail[4], fl + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&tail[5], fr + n_slots - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaStreamSynchronize(w->st));
        nk = tail[0] + tail[3];
        nl = tail[1] + tail[4];
        nr = tail[2] + tail[5];
    }
    uint32_t im_l = 0, im_r = 0;
    TRY(slab_exchange_counts(w, nl, nr, &im_l, &im_r));
    const uint32_t n_new = nk + im_l + im_r;
    // ---- migrate: kept particles compact into the other buffer, emigrants into send buffers -----------------------
    for (int a = 0; a < 3; ++a) {
        CU(S.out_l[a].ensure(std::max<uint32_t>(nl, 1)));
        CU(S.out_r[a].ensure(std::max<uint32_t>(nr, 1)));
    }
    CU(S.gid_l.ensure(std::max<uint32_t>(nl, 1)));
    CU(S.gid_r.ensure(std::max<uint32_t>(nr, 1)));
    w->Ntot = n_new;  // sizes the destination buffers (ghost room is added below)
    w->N = n_new;
    w->fluids[0].n = n_new;
    w->fluids[0].pending_delete.assign(n_new, 0);
    recompute_offsets(w);
    {   // make sure the destination (c^1) buffers can hold kept + immigrants (+ ghosts appended later grow again)
        size_t need = n_new;
        CU(w->pos[c ^ 1].ensure(need));
        CU(w->vel[c ^ 1].ensure(need));
        CU(w->vc[c ^ 1].ensure(need));
        CU(w->orig[c ^ 1].ensure(need));
        CU(w->gid[c ^ 1].ensure(need));
    }
    LAUNCH(k_slab_scatter, n_slots, 256, n_slots, ob, S.flag2.p, S.flag2.p + n_slots, S.flag2.p + 2 * (size_t)n_slots, fk, fl, fr, S.flag_o.p, w->pos[c].p,
           w->vel[c].p, w->vc[c].p, w->orig[c].p, w->gid[c].p, w->pos[c ^ 1].p, w->vel[c ^ 1].p, w->vc[c ^ 1].p, w->orig[c ^ 1].p, w->gid[c ^ 1].p,
           S.out_l[0].p, S.out_l[1].p, S.out_l[2].p, S.gid_l.p, S.out_r[0].p, S.out_r[1].p, S.out_r[2].p, S.gid_r.p);
    c ^= 1;
    w->cur = c;
    NC(g_nccl.GroupStart());
    float4* dst4[3] = {w->pos[c].p, w->vel[c].p, w->vc[c].p};
    if (slab_left(w) >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nl) NC(g_nccl.Send(S.out_l[a].p, (size_t)nl * 16, NCCL_CHAR, slab_left(w), S.comm, w->st));
            if (im_l) NC(g_nccl.Recv(dst4[a] + nk, (size_t)im_l * 16, NCCL_CHAR, slab_left(w), S.comm, w->st));
        }
        if (nl) NC(g_nccl.Send(S.gid_l.p, (size_t)nl * 4, NCCL_CHAR, slab_left(w), S.comm, w->st));
        if (im_l) NC(g_nccl.Recv(w->gid[c].p + nk, (size_t)im_l * 4, NCCL_CHAR, slab_left(w), S.comm, w->st));
    }
    if (slab_right(w) >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nr) NC(g_nccl.Send(S.out_r[a].p, (size_t)nr * 16, NCCL_CHAR, slab_right(w), S.comm, w->st));
            if (im_r) NC(g_nccl.Recv(dst4[a] + nk + im_l, (size_t)im_r * 16, NCCL_CHAR, slab_right(w), S.comm, w->st));
        }
        if (nr) NC(g_nccl.Send(S.gid_r.p, (size_t)nr * 4, NCCL_CHAR, slab_right(w), S.comm, w->st));
        if (im_r) NC(g_nccl.Recv(w->gid[c].p + nk + im_l, (size_t)im_r * 4, NCCL_CHAR, slab_right(w), S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    if (im_l + im_r) LAUNCH(k_iota_from, im_l + im_r, 256, im_l + im_r, nk, w->orig[c].p + nk);
    S.migrated_out = nl + nr;
    S.migrated_in = im_l + im_r;
    // ---- 2. boundary columns -> neighbours' ghosts ------------------------------------------------------------------
    CU(S.flag.ensure(2 * (size_t)n_new + 8));
    uint32_t* gl = S.flag.p;
    uint32_t* gr = S.flag.p + n_new;
    CU(cudaMemsetAsync(S.flag.p, 0, 2 * (size_t)n_new * sizeof(uint32_t), w->st));
    LAUNCH(k_slab_column_flags, n_new, 256, w->pos[c].p, n_new, S.lo, S.hi, S.has_left, S.has_right, gl, gr);
    uint32_t t2[4] = {0, 0, 0, 0};
    if (n_new) {
        CU(cudaMemcpyAsync(&t2[0], gl + n_new - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&t2[1], gr + n_new - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(S.flag2.ensure(2 * (size_t)n_new + 8));
        CU(cudaMemcpyAsync(S.flag2.p, S.flag.p, 2 * (size_t)n_new * sizeof(uint32_t), cudaMemcpyDeviceToDevice, w->st));
        TRY(scan_exclusive(w, gl, n_new));
        TRY(scan_exclusive(w, gr, n_new));
        CU(cudaMemcpyAsync(&t2[2], gl + n_new - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaMemcpyAsync(&t2[3], gr + n_new - 1, 4, cudaMemcpyDeviceToHost, w->st));
        CU(cudaStreamSynchronize(w->st));
    }
    uint32_t sl = t2[0] + t2[2], sr = t2[1] + t2[3];
    uint32_t ghl = 0, ghr = 0;
    TRY(slab_exchange_counts(w, sl, sr, &ghl, &ghr));
    for (int a = 0; a < 3; ++a) {
        CU(S.out_l[a].ensure(std::max<uint32_t>(sl, 1)));
        CU(S.out_r[a].ensure(std::max<uint32_t>(sr, 1)));
    }
    if (n_new) LAUNCH(k_slab_pack_columns, n_new, 256, n_new, S.flag2.p, S.flag2.p + n_new, gl, gr, w->pos[c].p, w->vel[c].p, w->vc[c].p, S.out_l[0].p,
                      S.out_l[1].p, S.out_l[2].p, S.out_r[0].p, S.out_r[1].p, S.out_r[2].p);
    w->Ntot = (size_t)n_new + ghl + ghr;
    TRY(ensure_fluid_buffers(w));  // grows (keeping contents) every per-particle array to Ntot
    dst4[0] = w->pos[c].p; dst4[1] = w->vel[c].p; dst4[2] = w->vc[c].p;
    NC(g_nccl.GroupStart());
    if (slab_left(w) >= 0)
        for (int a = 0; a < 3; ++a) {
            if (sl) NC(g_nccl.Send(S.out_l[a].p, (size_t)sl * 16, NCCL_CHAR, slab_left(w), S.comm, w->st));
            if (ghl) NC(g_nccl.Recv(dst4[a] + n_new, (size_t)ghl * 16, NCCL_CHAR, slab_left(w), S.comm, w->st));
        }
    if (slab_right(w) >= 0)
        for (int a = 0; a < 3; ++a) {
            if (sr) NC(g_nccl.Send(S.out_r[a].p, (size_t)sr * 16, NCCL_CHAR, slab_right(w), S.comm, w->st));
            if (ghr) NC(g_nccl.Recv(dst4[a] + n_new + ghl, (size_t)ghr * 16, NCCL_CHAR, slab_right(w), S.comm, w->st));
        }
    NC(g_nccl.GroupEnd());
    if (ghl + ghr) {  // ghosts carry no original index / id
        CU(cudaMemsetAsync(w->orig[c].p + n_new, 0xFF, (size_t)(ghl + ghr) * 4, w->st));
        CU(cudaMemsetAsync(w->gid[c].p + n_new, 0xFF, (size_t)(ghl + ghr) * 4, w->st));
    }
    S.exp_ghost_l = ghl;
    S.exp_ghost_r = ghr;
    S.exp_send_l = sl;
    S.exp_send_r = sr;
    // global particle count of the fluid (mean errors are global means)
    {
        unsigned long long cnt = n_new;
        CU(cudaMemcpyAsync(S.d_cnt64.p, &cnt, 8, cudaMemcpyHostToDevice, w->st));
        NC(g_nccl.AllReduce(S.d_cnt64.p, S.d_cnt64.p, 1, NCCL_UINT64, NCCL_SUM, S.comm, w->st));
        CU(cudaMemcpyAsync(&cnt, S.d_cnt64.p, 8, cudaMemcpyDeviceToHost, w->st));
        CU(cudaStreamSynchronize(w->st));
        S.global_n = cnt;
    }
    w->own_begin = 0;
    return SPH_OK;
}

// After the sort: locate the owned range and the boundary / ghost column ranges in the sorted arrays.
sph_status slab_after_sort(sph_world* w) {
    SlabState& S = w->slab;
    const Consts& c = w->hc;
    auto col_start_index = [&](long long cx) -> long long {  // index into cstart of the first cell of column cx (absolute)
        long long gx = cx - c.ox;
        if (gx <= 0) return 0;
        if (gx >= c.nx) return (long long)c.nx * c.ny * c.nz;
        return gx * (long long)c.ny * c.nz;
    };
    long long idx[4] = {S.has_left ? col_start_index(S.lo) : 0, S.has_left ? col_start_index((long long)S.lo + 1) : 0,
                        S.has_right ? col_start_index((long long)S.hi - 1) : col_start_index(1LL << 40),
                        S.has_right ? col_start_index(S.hi) : col_start_index(1LL << 40)};
    uint32_t v[4];
    for (int a = 0; a < 4; ++a) CU(cudaMemcpyAsync(&v[a], w->cstart.p + idx[a], 4, cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    uint32_t ntot = (uint32_t)w->Ntot;
    S.gl_count = S.has_left ? v[0] : 0;
    S.sl_begin = v[0];
    S.sl_count = S.has_left ? v[1] - v[0] : 0;
    S.sr_begin = v[2];
    S.sr_count = S.has_right ? v[3] - v[2] : 0;
    S.gr_begin = S.has_right ? v[3] : ntot;
    S.gr_count = ntot - S.gr_begin;
    if (S.gl_count != S.exp_ghost_l || S.gr_count != S.exp_ghost_r || S.sl_count != S.exp_send_l || S.sr_count != S.exp_send_r)
        return w->fail(SPH_ERR_NCCL, "slab layout mismatch after sort: ghosts %u/%u (expected %u/%u), columns %u/%u (expected %u/%u)", S.gl_count,
                       S.gr_count, S.exp_ghost_l, S.exp_ghost_r, S.sl_count, S.sr_count, S.exp_send_l, S.exp_send_r);
    w->own_begin = S.gl_count;
    if (S.gr_begin - S.gl_count != (uint32_t)w->N) return w->fail(SPH_ERR_NCCL, "slab layout mismatch: owned range %u != %zu", S.gr_begin - S.gl_count, w->N);
    return SPH_OK;
}

}  // namespace
