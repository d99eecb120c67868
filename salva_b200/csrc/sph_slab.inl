// sph_slab.inl — multi-GPU 1-D slab decomposition along x.  Included by sph_engine.cu.
//
// One process per GPU.  Rank r owns the particles whose cell x-coordinate floor(x/h) lies in [lo_r, hi_r); every
// step it receives the neighbours' boundary cell columns as GHOST particles (one cell = one kernel radius wide).
// The x-major sort keeps a whole yz-plane of cells contiguous, so after the sort the arrays are laid out as
//     [ left ghosts | my left boundary column | interior | my right boundary column | right ghosts ]
// and every per-iteration exchange (rho, kappa, v*, normals) is an ncclSend/ncclRecv of a contiguous array range
// with NO pack kernel.  Sender and receiver agree on the order inside a column because the in-cell order of the
// counting sort is canonical (ascending particle id, k_cell_sort) and ghost particles travel WITH their ids
// (deterministic mode is forced on).
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
    if (S.comm_st) cudaStreamSynchronize(S.comm_st);
    if (S.ev_ready) cudaEventDestroy(S.ev_ready);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
    if (S.comm_st) cudaStreamDestroy(S.comm_st);
    S.comm_st = nullptr; S.ev_ready = nullptr; S.ev_done = nullptr;
    S.d_cnt.release(); S.flag.release(); S.flag_o.release(); S.gid_l.release(); S.gid_r.release(); S.gid_cl.release(); S.gid_cr.release();
    S.d_cnt64.release();
    for (int a = 0; a < 3; ++a) {
        S.out_l[a].release();
        S.out_r[a].release();
        S.col_l[a].release();
        S.col_r[a].release();
    }
    S.active = false;
}

inline bool slab_on(const sph_world* w) { return w->slab.active; }
inline int slab_left(const sph_world* w) { return w->slab.rank > 0 ? w->slab.rank - 1 : -1; }
inline int slab_right(const sph_world* w) { return w->slab.rank + 1 < w->slab.nranks ? w->slab.rank + 1 : -1; }

// Per-iteration ghost refresh of up to 4 per-particle arrays in ONE NCCL group (elem = bytes per particle): my boundary
// columns go to the neighbours, their boundary columns land in my ghost ranges.  Contiguous ranges, no packing.
sph_status slab_refresh_n(sph_world* w, const SlabArray* arrays, int n_arrays, cudaStream_t st) {
    if (!slab_on(w)) return SPH_OK;
    SlabState& S = w->slab;
    if (!st) st = w->st;
    NC(g_nccl.GroupStart());
    for (int k = 0; k < n_arrays; ++k) {
        char* a = static_cast<char*>(arrays[k].p);
        const size_t elem = arrays[k].elem;
        if (slab_left(w) >= 0) {
            if (S.sl_count) NC(g_nccl.Send(a + (size_t)S.sl_begin * elem, (size_t)S.sl_count * elem, NCCL_CHAR, slab_left(w), S.comm, st));
            if (S.gl_count) NC(g_nccl.Recv(a, (size_t)S.gl_count * elem, NCCL_CHAR, slab_left(w), S.comm, st));
        }
        if (slab_right(w) >= 0) {
            if (S.sr_count) NC(g_nccl.Send(a + (size_t)S.sr_begin * elem, (size_t)S.sr_count * elem, NCCL_CHAR, slab_right(w), S.comm, st));
            if (S.gr_count) NC(g_nccl.Recv(a + (size_t)S.gr_begin * elem, (size_t)S.gr_count * elem, NCCL_CHAR, slab_right(w), S.comm, st));
        }
    }
    NC(g_nccl.GroupEnd());
    w->stats_exchanges++;
    return SPH_OK;
}
sph_status slab_refresh(sph_world* w, void* array, size_t elem) {
    SlabArray a{array, elem};
    return slab_refresh_n(w, &a, 1);
}

// Sum a small device float buffer over all ranks (error means of the Jacobi loops: dfsph_solver.rs:153-158).
sph_status slab_allreduce(sph_world* w, float* buf, size_t n) {
    if (!slab_on(w)) return SPH_OK;
    NC(g_nccl.AllReduce(buf, buf, n, NCCL_FLOAT, NCCL_SUM, w->slab.comm, w->st));
    return SPH_OK;
}

// Step prologue in slab mode — one classification pass, one count exchange, ONE host sync, one data exchange:
//   * particles that left [lo, hi) migrate to the neighbour that now owns them (CFL: at most one cell column per step);
//   * the kept particles of my boundary columns go to the neighbours as their ghosts;
//   * my own emigrants stay here as ghosts (they now sit in the neighbour's boundary column), appended AFTER the
//     neighbour's column so that both sides see that column in the same order: [neighbour's kept | my emigrants].
// On exit the arrays hold [kept | immigrants left | immigrants right | left ghosts | right ghosts]; the counting sort
// follows and leaves [left ghosts | owned | right ghosts] with the boundary columns at the ends of the owned range.
sph_status slab_begin_step(sph_world* w) {
    SlabState& S = w->slab;
    if (w->fluids.size() != 1) return w->fail(SPH_ERR_INVALID, "slab decomposition supports one fluid per world");
    if (w->desc.solver != SPH_SOLVER_DFSPH || w->tile) return w->fail(SPH_ERR_INVALID, "slab decomposition supports DFSPH with gather_backend 0");
    for (auto& fr : w->fluids[0].forces)
        if (fr.d.kind == SPH_FORCE_BECKER2009_ELASTICITY) return w->fail(SPH_ERR_INVALID, "Becker2009 elasticity is not slab-decomposed");
    const uint32_t n_slots = (uint32_t)w->Ntot;  // layout of the previous step: [ghosts | owned | ghosts] (or owned only)
    const uint32_t ob = w->own_begin, on = (uint32_t)w->N;
    int c = w->cur;
    const int L = slab_left(w), R = slab_right(w);
    // ---- classify + count ---------------------------------------------------------------------------------------------
    CU(S.flag.ensure(10 * (size_t)n_slots + 16));
    CU(S.flag_o.ensure((size_t)on + 8));
    CU(S.d_cnt.ensure(32));
    uint32_t* f[5];   // keep, left, right, col-left, col-right flags
    uint32_t* sc[5];  // their exclusive scans
    for (int a = 0; a < 5; ++a) {
        f[a] = S.flag.p + (size_t)a * n_slots;
        sc[a] = S.flag.p + (size_t)(5 + a) * n_slots;
    }
    CU(cudaMemsetAsync(S.flag.p, 0, 5 * (size_t)n_slots * sizeof(uint32_t), w->st));
    CU(cudaMemsetAsync(S.d_cnt.p, 0, 32 * sizeof(uint32_t), w->st));
    LAUNCH(k_slab_classify, on, 256, w->pos[c].p, w->orig[c].p, ob, on, S.lo, S.hi, S.has_left, S.has_right, f[0], f[1], f[2], f[3], f[4], S.flag_o.p,
           S.d_cnt.p);
    // counts go to the neighbours straight from the device: [#emigrants to you, #particles of my column facing you]
    LAUNCH(k_slab_pack_counts, 1, 32, S.d_cnt.p);  // d_cnt[8..9] = {nl, ncl}, d_cnt[10..11] = {nr, ncr}
    NC(g_nccl.GroupStart());
    if (L >= 0) {
        NC(g_nccl.Send(S.d_cnt.p + 8, 8, NCCL_CHAR, L, S.comm, w->st));
        NC(g_nccl.Recv(S.d_cnt.p + 12, 8, NCCL_CHAR, L, S.comm, w->st));
    }
    if (R >= 0) {
        NC(g_nccl.Send(S.d_cnt.p + 10, 8, NCCL_CHAR, R, S.comm, w->st));
        NC(g_nccl.Recv(S.d_cnt.p + 14, 8, NCCL_CHAR, R, S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    uint32_t hc[16];
    CU(cudaMemcpyAsync(hc, S.d_cnt.p, sizeof hc, cudaMemcpyDeviceToHost, w->st));
    CU(cudaMemcpyAsync(sc[0], f[0], 5 * (size_t)n_slots * sizeof(uint32_t), cudaMemcpyDeviceToDevice, w->st));
    {
        ScanSet<5> set;
        for (int a = 0; a < 5; ++a) set.a[a] = sc[a];
        TRY(scan_exclusive_k<5>(w, set, n_slots));  // one 3-launch scan for the five flag arrays
    }
    TRY(scan_exclusive(w, S.flag_o.p, on));  // new original index of the kept particles (stable in the old order)
    CU(cudaStreamSynchronize(w->st));        // the only host sync of the prologue
    const uint32_t nk = hc[0], nl = hc[1], nr = hc[2], ncl = hc[3], ncr = hc[4];
    if (hc[5]) return w->fail(SPH_ERR_INVALID, "%u particles crossed more than one cell column in a step (CFL violated)", hc[5]);
    const uint32_t im_l = L >= 0 ? hc[12] : 0, gcol_l = L >= 0 ? hc[13] : 0;  // from the left rank: its emigrants to me, its column
    const uint32_t im_r = R >= 0 ? hc[14] : 0, gcol_r = R >= 0 ? hc[15] : 0;
    const uint32_t n_new = nk + im_l + im_r;
    const uint32_t ghl = gcol_l + nl, ghr = gcol_r + nr;
    // ---- buffers ------------------------------------------------------------------------------------------------------------
    for (int a = 0; a < 3; ++a) {
        CU(S.out_l[a].ensure(std::max<uint32_t>(nl, 1)));
        CU(S.out_r[a].ensure(std::max<uint32_t>(nr, 1)));
        CU(S.col_l[a].ensure(std::max<uint32_t>(ncl, 1)));
        CU(S.col_r[a].ensure(std::max<uint32_t>(ncr, 1)));
    }
    CU(S.gid_l.ensure(std::max<uint32_t>(nl, 1)));
    CU(S.gid_r.ensure(std::max<uint32_t>(nr, 1)));
    CU(S.gid_cl.ensure(std::max<uint32_t>(ncl, 1)));
    CU(S.gid_cr.ensure(std::max<uint32_t>(ncr, 1)));
    w->N = n_new;
    w->Ntot = (size_t)n_new + ghl + ghr;
    w->fluids[0].n = n_new;
    w->fluids[0].pending_delete.assign(n_new, 0);
    recompute_offsets(w);
    w->cur = c ^ 1;          // the compacted state is built in the other buffer ...
    w->protect_buf = c;      // ... while the old one is still being read by the scatter
    TRY(ensure_fluid_buffers(w));
    const int d = c ^ 1;
    SlabOut keep{w->pos[d].p, w->vel[d].p, w->vc[d].p, w->gid[d].p};
    SlabOut ol{S.out_l[0].p, S.out_l[1].p, S.out_l[2].p, S.gid_l.p}, orr{S.out_r[0].p, S.out_r[1].p, S.out_r[2].p, S.gid_r.p};
    SlabOut cl{S.col_l[0].p, S.col_l[1].p, S.col_l[2].p, S.gid_cl.p}, cr{S.col_r[0].p, S.col_r[1].p, S.col_r[2].p, S.gid_cr.p};
    LAUNCH(k_slab_scatter, n_slots, 256, n_slots, f[0], f[1], f[2], f[3], f[4], sc[0], sc[1], sc[2], sc[3], sc[4], S.flag_o.p, w->pos[c].p, w->vel[c].p,
           w->vc[c].p, w->orig[c].p, w->gid[c].p, keep, w->orig[d].p, ol, orr, cl, cr);
    // ---- one data exchange: emigrants + boundary columns out, immigrants + ghost columns in --------------------------------
    float4* dst4[3] = {w->pos[d].p, w->vel[d].p, w->vc[d].p};
    const uint32_t g0 = n_new, g1 = n_new + ghl;  // first left / right ghost slot
    NC(g_nccl.GroupStart());
    if (L >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nl) NC(g_nccl.Send(S.out_l[a].p, (size_t)nl * 16, NCCL_CHAR, L, S.comm, w->st));
            if (ncl) NC(g_nccl.Send(S.col_l[a].p, (size_t)ncl * 16, NCCL_CHAR, L, S.comm, w->st));
            if (im_l) NC(g_nccl.Recv(dst4[a] + nk, (size_t)im_l * 16, NCCL_CHAR, L, S.comm, w->st));
            if (gcol_l) NC(g_nccl.Recv(dst4[a] + g0, (size_t)gcol_l * 16, NCCL_CHAR, L, S.comm, w->st));
        }
        if (nl) NC(g_nccl.Send(S.gid_l.p, (size_t)nl * 4, NCCL_CHAR, L, S.comm, w->st));
        if (ncl) NC(g_nccl.Send(S.gid_cl.p, (size_t)ncl * 4, NCCL_CHAR, L, S.comm, w->st));
        if (im_l) NC(g_nccl.Recv(w->gid[d].p + nk, (size_t)im_l * 4, NCCL_CHAR, L, S.comm, w->st));
        if (gcol_l) NC(g_nccl.Recv(w->gid[d].p + g0, (size_t)gcol_l * 4, NCCL_CHAR, L, S.comm, w->st));
    }
    if (R >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nr) NC(g_nccl.Send(S.out_r[a].p, (size_t)nr * 16, NCCL_CHAR, R, S.comm, w->st));
            if (ncr) NC(g_nccl.Send(S.col_r[a].p, (size_t)ncr * 16, NCCL_CHAR, R, S.comm, w->st));
            if (im_r) NC(g_nccl.Recv(dst4[a] + nk + im_l, (size_t)im_r * 16, NCCL_CHAR, R, S.comm, w->st));
            if (gcol_r) NC(g_nccl.Recv(dst4[a] + g1, (size_t)gcol_r * 16, NCCL_CHAR, R, S.comm, w->st));
        }
        if (nr) NC(g_nccl.Send(S.gid_r.p, (size_t)nr * 4, NCCL_CHAR, R, S.comm, w->st));
        if (ncr) NC(g_nccl.Send(S.gid_cr.p, (size_t)ncr * 4, NCCL_CHAR, R, S.comm, w->st));
        if (im_r) NC(g_nccl.Recv(w->gid[d].p + nk + im_l, (size_t)im_r * 4, NCCL_CHAR, R, S.comm, w->st));
        if (gcol_r) NC(g_nccl.Recv(w->gid[d].p + g1, (size_t)gcol_r * 4, NCCL_CHAR, R, S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    // my own emigrants are my ghosts now (after the neighbour's column, see the header comment)
    for (int a = 0; a < 3; ++a) {
        if (nl) CU(cudaMemcpyAsync(dst4[a] + g0 + gcol_l, S.out_l[a].p, (size_t)nl * 16, cudaMemcpyDeviceToDevice, w->st));
        if (nr) CU(cudaMemcpyAsync(dst4[a] + g1 + gcol_r, S.out_r[a].p, (size_t)nr * 16, cudaMemcpyDeviceToDevice, w->st));
    }
    if (im_l + im_r) LAUNCH(k_iota_from, im_l + im_r, 256, im_l + im_r, nk, w->orig[d].p + nk);
    if (nl) CU(cudaMemcpyAsync(w->gid[d].p + g0 + gcol_l, S.gid_l.p, (size_t)nl * 4, cudaMemcpyDeviceToDevice, w->st));
    if (nr) CU(cudaMemcpyAsync(w->gid[d].p + g1 + gcol_r, S.gid_r.p, (size_t)nr * 4, cudaMemcpyDeviceToDevice, w->st));
    if (ghl + ghr) CU(cudaMemsetAsync(w->orig[d].p + n_new, 0xFF, (size_t)(ghl + ghr) * 4, w->st));  // ghosts carry no original index (ids they do: the sort key)
    S.migrated_out = nl + nr;
    S.migrated_in = im_l + im_r;
    // slot ranges after the sort follow from the counts alone (CFL: immigrants land in my boundary columns)
    S.gl_count = ghl;
    S.sl_begin = ghl;
    S.sl_count = L >= 0 ? ncl + im_l : 0;
    S.sr_count = R >= 0 ? ncr + im_r : 0;
    S.sr_begin = ghl + n_new - S.sr_count;
    S.gr_begin = ghl + n_new;
    S.gr_count = ghr;
    if (!S.global_valid) {  // particles are conserved by migration: the global count only changes through the host API
        unsigned long long cnt = n_new;
        CU(cudaMemcpyAsync(S.d_cnt64.p, &cnt, 8, cudaMemcpyHostToDevice, w->st));
        NC(g_nccl.AllReduce(S.d_cnt64.p, S.d_cnt64.p, 1, NCCL_UINT64, NCCL_SUM, S.comm, w->st));
        CU(cudaMemcpyAsync(&cnt, S.d_cnt64.p, 8, cudaMemcpyDeviceToHost, w->st));
        CU(cudaStreamSynchronize(w->st));
        S.global_n = cnt;
        S.global_valid = true;
    }
    w->own_begin = 0;  // until the sort
    w->protect_buf = -1;
    TRY(ensure_fluid_buffers(w));  // now the old buffer may grow too (it is the sort's destination)
    return SPH_OK;
}

// After the sort the owned range starts behind the left ghosts.  With SALVA_B200_SLAB_CHECK=1 the ranges derived from
// the exchanged counts are verified against the sorted cell table.
sph_status slab_after_sort(sph_world* w) {
    SlabState& S = w->slab;
    w->own_begin = S.gl_count;
    static const bool check = getenv("SALVA_B200_SLAB_CHECK") && atoi(getenv("SALVA_B200_SLAB_CHECK")) != 0;
    if (!check) return SPH_OK;
    const Consts& c = w->hc;
    auto col_start_index = [&](long long cx) -> long long {
        long long gx = cx - c.ox;
        if (gx <= 0) return 0;
        if (gx >= c.nx) return (long long)c.nx * c.ny * c.nz;
        return gx * (long long)c.ny * c.nz;
    };
    const long long far = 1LL << 40;
    long long idx[4] = {S.has_left ? col_start_index(S.lo) : 0, S.has_left ? col_start_index((long long)S.lo + 1) : 0,
                        S.has_right ? col_start_index((long long)S.hi - 1) : col_start_index(far), S.has_right ? col_start_index(S.hi) : col_start_index(far)};
    uint32_t v[4];
    for (int a = 0; a < 4; ++a) CU(cudaMemcpyAsync(&v[a], w->cstart.p + idx[a], 4, cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    uint32_t ntot = (uint32_t)w->Ntot;
    uint32_t gl = S.has_left ? v[0] : 0, sl = S.has_left ? v[1] - v[0] : 0, sr = S.has_right ? v[3] - v[2] : 0, grb = S.has_right ? v[3] : ntot;
    if (gl != S.gl_count || sl != S.sl_count || sr != S.sr_count || grb != S.gr_begin || (S.has_right && v[2] != S.sr_begin))
        return w->fail(SPH_ERR_NCCL, "slab layout mismatch after sort: ghosts-left %u (expected %u), columns %u/%u (expected %u/%u), right ghosts at %u (expected %u)",
                       gl, S.gl_count, sl, sr, S.sl_count, S.sr_count, grb, S.gr_begin);
    return SPH_OK;
}

}  // namespace
