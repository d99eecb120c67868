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
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int NCCL_CHAR = 0, NCCL_INT = 2, NCCL_FLOAT = 7, NCCL_UINT64 = 5, NCCL_SUM = 0, NCCL_MIN = 3;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

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
    NSYM(AllGather, "ncclAllGather")
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

void p2p_release(sph_world* w) {
    P2PState& P = w->slab.p2p;
    for (int p = 0; p < 8; ++p)
        if (P.peer_base[p] && P.peer_base[p] != P.base) cudaIpcCloseMemHandle(P.peer_base[p]);
    if (P.base) cudaFree(P.base);
    P.tickets.release();
    P = P2PState();
}

void slab_release(sph_world* w) {
    SlabState& S = w->slab;
    p2p_release(w);
    if (S.comm && S.own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(S.comm);
    S.comm = nullptr;
    if (S.comm_st) cudaStreamSynchronize(S.comm_st);
    if (S.ev_ready) cudaEventDestroy(S.ev_ready);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
    if (S.comm_st) cudaStreamDestroy(S.comm_st);
    S.comm_st = nullptr; S.ev_ready = nullptr; S.ev_done = nullptr;
    S.d_cnt.release(); S.flag.release(); S.gid_l.release(); S.gid_r.release(); S.gid_cl.release(); S.gid_cr.release();
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

// ---- NVLink peer-memory exchange (k_p2p_push / k_p2p_pull, sph_kernels.cuh) ---------------------------------------------
// One device allocation per rank, exported with cudaIpc to every other rank of the node:
//   [box L buf0 | box L buf1 | box R buf0 | box R buf1 | 16 flags | reduction table 2 x 8 x MAX_FLUIDS | reduction flags 2 x 8]
// "box L" receives what my LEFT neighbour sends me.  Messages are double buffered by the parity of their sequence number
// (a sender can run at most one exchange ahead of the receiver's copy-out).
inline size_t p2p_off_flags(const P2PState& P) { return 4 * P.box_bytes; }
inline size_t p2p_off_red(const P2PState& P) { return 4 * P.box_bytes + 256; }
inline size_t p2p_off_redflag(const P2PState& P) { return p2p_off_red(P) + 2 * 8 * MAX_FLUIDS * sizeof(float); }
inline size_t p2p_total(const P2PState& P) { return p2p_off_redflag(P) + 2 * 8 * sizeof(uint32_t) + 256; }

// Collective over the slab communicator.  On any failure (IPC not permitted, > 8 ranks, ...) every rank keeps the NCCL path.
sph_status p2p_setup(sph_world* w) {
    SlabState& S = w->slab;
    P2PState& P = S.p2p;
    int want = 1;
    if (const char* t = getenv("SALVA_B200_P2P")) want = atoi(t) != 0;
    if (S.nranks > 8 || S.nranks < 2) want = 0;
    size_t box_mb = 16;
    if (const char* t = getenv("SALVA_B200_P2P_BOX_MB")) box_mb = (size_t)std::max(1, atoi(t));
    P.box_bytes = box_mb << 20;
    int ok = want;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok) {
        ok = cudaMalloc(&P.base, p2p_total(P)) == cudaSuccess && cudaMemset(P.base, 0, p2p_total(P)) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine, P.base) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    // handles of all ranks (64 bytes each) + a "still fine" vote, both through NCCL (the plumbing that exists anyway)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    char* d_h = nullptr;
    int* d_ok = nullptr;
    CU(cudaMalloc(&d_h, 64 * (size_t)(S.nranks + 1)));
    CU(cudaMalloc(&d_ok, sizeof(int)));
    CU(cudaMemcpyAsync(d_h, &mine, 64, cudaMemcpyHostToDevice, w->st));
    NC(g_nccl.AllGather(d_h, d_h + 64, 64, NCCL_CHAR, S.comm, w->st));
    std::vector<cudaIpcMemHandle_t> all(S.nranks);
    CU(cudaMemcpyAsync(all.data(), d_h + 64, 64 * (size_t)S.nranks, cudaMemcpyDeviceToHost, w->st));
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, w->st));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, NCCL_INT, NCCL_MIN, S.comm, w->st));
    CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    if (ok) {
        for (int p = 0; p < S.nranks && ok; ++p) {
            if (p == S.rank) {
                P.peer_base[p] = P.base;
                continue;
            }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            }
            P.peer_base[p] = static_cast<char*>(q);
        }
        CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, w->st));
        NC(g_nccl.AllReduce(d_ok, d_ok, 1, NCCL_INT, NCCL_MIN, S.comm, w->st));
        CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, w->st));
        CU(cudaStreamSynchronize(w->st));
    }
    cudaFree(d_h);
    cudaFree(d_ok);
    if (!ok) {
        p2p_release(w);
        return SPH_OK;  // NCCL path stays in charge
    }
    CU(P.tickets.ensure(4));
    CU(cudaMemset(P.tickets.p, 0, 4 * sizeof(uint32_t)));
    P.on = true;
    return SPH_OK;
}

// Per-iteration ghost refresh of up to 4 per-particle arrays (elem = bytes per particle): my boundary columns go to the
// neighbours, their boundary columns land in my ghost ranges.  Contiguous ranges, no packing.  Messages that fit the peer
// boxes travel as direct NVLink stores (one push + one pull kernel for both directions); the rest (or everything, when the
// peer mapping is unavailable) as ONE NCCL send/recv group.
sph_status slab_refresh_n(sph_world* w, const SlabArray* arrays, int n_arrays, cudaStream_t st) {
    if (!slab_on(w)) return SPH_OK;
    SlabState& S = w->slab;
    P2PState& P = S.p2p;
    if (!st) st = w->st;
    const int nb[2] = {slab_left(w), slab_right(w)};
    const uint32_t send_begin[2] = {S.sl_begin, S.sr_begin}, send_count[2] = {S.sl_count, S.sr_count};
    const uint32_t recv_begin[2] = {0u, S.gr_begin}, recv_count[2] = {S.gl_count, S.gr_count};
    bool p2p_send[2] = {false, false}, p2p_recv[2] = {false, false};
    if (P.on && n_arrays <= 4) {
        for (int side = 0; side < 2; ++side) {
            if (nb[side] < 0) continue;
            size_t sb = 0, rb = 0;
            for (int k = 0; k < n_arrays; ++k) {
                sb += ((size_t)send_count[side] * arrays[k].elem + 15) & ~(size_t)15;
                rb += ((size_t)recv_count[side] * arrays[k].elem + 15) & ~(size_t)15;
            }
            p2p_send[side] = send_count[side] && sb <= P.box_bytes;
            p2p_recv[side] = recv_count[side] && rb <= P.box_bytes;
        }
    }
    // ---- NCCL for whatever does not go through the boxes ----------------------------------------------------------------
    bool any_nccl = false;
    for (int side = 0; side < 2; ++side)
        if (nb[side] >= 0 && ((send_count[side] && !p2p_send[side]) || (recv_count[side] && !p2p_recv[side]))) any_nccl = true;
    if (any_nccl) {
        NC(g_nccl.GroupStart());
        for (int k = 0; k < n_arrays; ++k) {
            char* a = static_cast<char*>(arrays[k].p);
            const size_t elem = arrays[k].elem;
            for (int side = 0; side < 2; ++side) {
                if (nb[side] < 0) continue;
                if (send_count[side] && !p2p_send[side])
                    NC(g_nccl.Send(a + (size_t)send_begin[side] * elem, (size_t)send_count[side] * elem, NCCL_CHAR, nb[side], S.comm, st));
                if (recv_count[side] && !p2p_recv[side])
                    NC(g_nccl.Recv(a + (size_t)recv_begin[side] * elem, (size_t)recv_count[side] * elem, NCCL_CHAR, nb[side], S.comm, st));
            }
        }
        NC(g_nccl.GroupEnd());
    }
    // ---- peer-memory path ------------------------------------------------------------------------------------------------
    if (p2p_send[0] || p2p_send[1] || p2p_recv[0] || p2p_recv[1]) {
        P2PMsg out[2], in[2];
        memset(out, 0, sizeof out);
        memset(in, 0, sizeof in);
        uint32_t seq_out[2] = {0, 0}, seq_in[2] = {0, 0}, max_words = 0;
        for (int side = 0; side < 2; ++side) {
            if (p2p_send[side]) {
                const uint32_t seq = ++P.seq_send[side];
                const int buf = (int)(seq & 1u);
                const int their_side = 1 - side;  // I am my left neighbour's RIGHT neighbour
                char* peer = P.peer_base[nb[side]];
                out[side].box = peer + ((size_t)their_side * 2 + buf) * P.box_bytes;
                out[side].flag = reinterpret_cast<uint32_t*>(peer + p2p_off_flags(P)) + their_side * 2 + buf;
                uint32_t off = 0;
                for (int k = 0; k < n_arrays; ++k) {
                    const uint32_t bytes = (uint32_t)((size_t)send_count[side] * arrays[k].elem);
                    out[side].seg[k] = P2PSeg{static_cast<const char*>(arrays[k].p) + (size_t)send_begin[side] * arrays[k].elem, nullptr, off, bytes};
                    off += (bytes + 15u) & ~15u;
                    out[side].total_words += bytes >> 2;
                }
                out[side].n_seg = n_arrays;
                seq_out[side] = seq;
                max_words = std::max(max_words, out[side].total_words);
            }
            if (p2p_recv[side]) {
                const uint32_t seq = ++P.seq_recv[side];
                const int buf = (int)(seq & 1u);
                in[side].box = P.base + ((size_t)side * 2 + buf) * P.box_bytes;
                in[side].flag = reinterpret_cast<uint32_t*>(P.base + p2p_off_flags(P)) + side * 2 + buf;
                uint32_t off = 0;
                for (int k = 0; k < n_arrays; ++k) {
                    const uint32_t bytes = (uint32_t)((size_t)recv_count[side] * arrays[k].elem);
                    in[side].seg[k] = P2PSeg{nullptr, static_cast<char*>(arrays[k].p) + (size_t)recv_begin[side] * arrays[k].elem, off, bytes};
                    off += (bytes + 15u) & ~15u;
                    in[side].total_words += bytes >> 2;
                }
                in[side].n_seg = n_arrays;
                seq_in[side] = seq;
                max_words = std::max(max_words, in[side].total_words);
            }
        }
        const uint32_t blocks = std::max(1u, std::min(64u, cdiv(max_words / 4 + 1, 1024)));
        if (p2p_send[0] || p2p_send[1]) {
            k_p2p_push<<<dim3(blocks, 2), 256, 0, st>>>(out[0], out[1], seq_out[0], seq_out[1], P.tickets.p);
            w->launches++;
        }
        if (p2p_recv[0] || p2p_recv[1]) {
            k_p2p_pull<<<dim3(blocks, 2), 256, 0, st>>>(in[0], in[1], seq_in[0], seq_in[1], w->d_scal.p + 7);
            w->launches++;
        }
        CU(cudaGetLastError());
    }
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
    SlabState& S = w->slab;
    if (S.p2p.on && n <= (size_t)MAX_FLUIDS) {  // a handful of floats: one 1-block kernel writing into every rank's table over NVLink
        P2PPeers peers;
        memset(&peers, 0, sizeof peers);
        for (int p = 0; p < S.nranks; ++p) {
            peers.red[p] = reinterpret_cast<float*>(S.p2p.peer_base[p] + p2p_off_red(S.p2p));
            peers.red_flag[p] = reinterpret_cast<uint32_t*>(S.p2p.peer_base[p] + p2p_off_redflag(S.p2p));
        }
        k_p2p_allreduce<<<1, 128, 0, w->st>>>(buf, (int)n, S.rank, S.nranks, peers, ++S.p2p.red_seq, w->d_scal.p + 7);
        w->launches++;
        CU(cudaGetLastError());
        return SPH_OK;
    }
    NC(g_nccl.AllReduce(buf, buf, n, NCCL_FLOAT, NCCL_SUM, S.comm, w->st));
    return SPH_OK;
}

// Step prologue in slab mode — one classification pass, one count exchange, ONE host sync, one data exchange, NO compaction:
//   * particles that left [lo, hi) are flagged dead (the counting sort that follows simply does not pick them up) and go to
//     the neighbour that now owns them (CFL: at most one cell column per step);
//   * the kept particles of my boundary columns go to the neighbours as their ghosts;
//   * my own emigrants stay here as ghosts (they now sit in the neighbour's boundary column).
// Immigrants and the new ghosts are appended behind the owned range of the LIVE arrays (over last step's right ghosts), so
// the sort's input is the slot range [own_begin, own_begin + on + immigrants + ghosts); its output is
// [left ghosts | owned | right ghosts] with the boundary columns at the ends of the owned range.  Only the handful of
// particles that cross a rank boundary are ever copied by the prologue (round 1 compacted the whole state: 0.6 ms of a 4 ms
// step at 4M particles per GPU, profiles/r2_multi_gpu.md).
// Index order in a slab world: fluid.positions[i] of a slab world is the engine's sorted order of the moment (it changes
// with every step); particles are tracked by their ids (sph_fluid_read_ids), which is what the migration preserves.
sph_status slab_begin_step(sph_world* w) {
    SlabState& S = w->slab;
    if (w->fluids.size() != 1) return w->fail(SPH_ERR_INVALID, "slab decomposition supports one fluid per world");
    if (w->desc.solver != SPH_SOLVER_DFSPH || w->tile) return w->fail(SPH_ERR_INVALID, "slab decomposition supports DFSPH with gather_backend 0");
    for (auto& fr : w->fluids[0].forces)
        if (fr.d.kind == SPH_FORCE_BECKER2009_ELASTICITY) return w->fail(SPH_ERR_INVALID, "Becker2009 elasticity is not slab-decomposed");
    const uint32_t ob = w->own_begin, on = (uint32_t)w->N;
    const int c = w->cur;
    const int L = slab_left(w), R = slab_right(w);
    CU(S.flag.ensure((size_t)on + 16));
    CU(S.d_cnt.ensure(32));
    uint32_t hc[16];
    for (int attempt = 0;; ++attempt) {
        for (int a = 0; a < 3; ++a) {
            CU(S.out_l[a].ensure(S.cap_out));
            CU(S.out_r[a].ensure(S.cap_out));
            CU(S.col_l[a].ensure(S.cap_col));
            CU(S.col_r[a].ensure(S.cap_col));
        }
        CU(S.gid_l.ensure(S.cap_out));
        CU(S.gid_r.ensure(S.cap_out));
        CU(S.gid_cl.ensure(S.cap_col));
        CU(S.gid_cr.ensure(S.cap_col));
        SlabOut ol{S.out_l[0].p, S.out_l[1].p, S.out_l[2].p, S.gid_l.p}, orr{S.out_r[0].p, S.out_r[1].p, S.out_r[2].p, S.gid_r.p};
        SlabOut cl{S.col_l[0].p, S.col_l[1].p, S.col_l[2].p, S.gid_cl.p}, cr{S.col_r[0].p, S.col_r[1].p, S.col_r[2].p, S.gid_cr.p};
        CU(cudaMemsetAsync(S.d_cnt.p, 0, 12 * sizeof(uint32_t), w->st));
        LAUNCH(k_slab_classify, on, 256, w->pos[c].p, w->vel[c].p, w->vc[c].p, w->gid[c].p, ob, on, S.lo, S.hi, S.has_left, S.has_right, S.flag.p, ol, orr, cl,
               cr, S.cap_out, S.cap_col, S.d_cnt.p);
        if (attempt == 0) {
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
            CU(cudaMemcpyAsync(hc, S.d_cnt.p, sizeof hc, cudaMemcpyDeviceToHost, w->st));
            CU(cudaStreamSynchronize(w->st));  // the only host sync of the prologue
        }
        if (hc[1] <= S.cap_out && hc[2] <= S.cap_out && hc[3] <= S.cap_col && hc[4] <= S.cap_col) break;
        if (attempt) return w->fail(SPH_ERR_INVALID, "slab staging buffers did not grow");
        // the staging buffers were too small (first step of a big scene): grow them and classify once more — the counts, and
        // therefore what the neighbours were told, do not change
        S.cap_out = std::max(S.cap_out, std::max(hc[1], hc[2]) + std::max(hc[1], hc[2]) / 2 + 1024);
        S.cap_col = std::max(S.cap_col, std::max(hc[3], hc[4]) + std::max(hc[3], hc[4]) / 2 + 1024);
    }
    const uint32_t nl = hc[1], nr = hc[2], ncl = hc[3], ncr = hc[4];
    if (hc[5]) return w->fail(SPH_ERR_INVALID, "%u particles crossed more than one cell column in a step (CFL violated)", hc[5]);
    const uint32_t im_l = L >= 0 ? hc[12] : 0, gcol_l = L >= 0 ? hc[13] : 0;  // from the left rank: its emigrants to me, its column
    const uint32_t im_r = R >= 0 ? hc[14] : 0, gcol_r = R >= 0 ? hc[15] : 0;
    const uint32_t n_new = on - nl - nr + im_l + im_r;
    const uint32_t ghl = gcol_l + nl, ghr = gcol_r + nr;
    const uint32_t n_in = on + im_l + im_r + ghl + ghr;  // slots the sort reads: owned (incl. the dead ones) + appended
    // ---- grow the live arrays behind the owned range if needed (contents kept) ----------------------------------------------
    w->Ntot = (size_t)ob + n_in;  // sizing only; the real value follows below
    TRY(ensure_fluid_buffers(w));
    float4* dst4[3] = {w->pos[c].p, w->vel[c].p, w->vc[c].p};
    const uint32_t a0 = ob + on;                    // first appended slot: immigrants from the left, then from the right,
    const uint32_t g0 = a0 + im_l + im_r;           // then the left ghosts, then the right ghosts
    const uint32_t g1 = g0 + ghl;
    // ---- one data exchange: emigrants + boundary columns out, immigrants + ghost columns in --------------------------------
    NC(g_nccl.GroupStart());
    if (L >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nl) NC(g_nccl.Send(S.out_l[a].p, (size_t)nl * 16, NCCL_CHAR, L, S.comm, w->st));
            if (ncl) NC(g_nccl.Send(S.col_l[a].p, (size_t)ncl * 16, NCCL_CHAR, L, S.comm, w->st));
            if (im_l) NC(g_nccl.Recv(dst4[a] + a0, (size_t)im_l * 16, NCCL_CHAR, L, S.comm, w->st));
            if (gcol_l) NC(g_nccl.Recv(dst4[a] + g0, (size_t)gcol_l * 16, NCCL_CHAR, L, S.comm, w->st));
        }
        if (nl) NC(g_nccl.Send(S.gid_l.p, (size_t)nl * 4, NCCL_CHAR, L, S.comm, w->st));
        if (ncl) NC(g_nccl.Send(S.gid_cl.p, (size_t)ncl * 4, NCCL_CHAR, L, S.comm, w->st));
        if (im_l) NC(g_nccl.Recv(w->gid[c].p + a0, (size_t)im_l * 4, NCCL_CHAR, L, S.comm, w->st));
        if (gcol_l) NC(g_nccl.Recv(w->gid[c].p + g0, (size_t)gcol_l * 4, NCCL_CHAR, L, S.comm, w->st));
    }
    if (R >= 0) {
        for (int a = 0; a < 3; ++a) {
            if (nr) NC(g_nccl.Send(S.out_r[a].p, (size_t)nr * 16, NCCL_CHAR, R, S.comm, w->st));
            if (ncr) NC(g_nccl.Send(S.col_r[a].p, (size_t)ncr * 16, NCCL_CHAR, R, S.comm, w->st));
            if (im_r) NC(g_nccl.Recv(dst4[a] + a0 + im_l, (size_t)im_r * 16, NCCL_CHAR, R, S.comm, w->st));
            if (gcol_r) NC(g_nccl.Recv(dst4[a] + g1, (size_t)gcol_r * 16, NCCL_CHAR, R, S.comm, w->st));
        }
        if (nr) NC(g_nccl.Send(S.gid_r.p, (size_t)nr * 4, NCCL_CHAR, R, S.comm, w->st));
        if (ncr) NC(g_nccl.Send(S.gid_cr.p, (size_t)ncr * 4, NCCL_CHAR, R, S.comm, w->st));
        if (im_r) NC(g_nccl.Recv(w->gid[c].p + a0 + im_l, (size_t)im_r * 4, NCCL_CHAR, R, S.comm, w->st));
        if (gcol_r) NC(g_nccl.Recv(w->gid[c].p + g1, (size_t)gcol_r * 4, NCCL_CHAR, R, S.comm, w->st));
    }
    NC(g_nccl.GroupEnd());
    // my own emigrants are my ghosts now (they sit in the neighbour's boundary column)
    for (int a = 0; a < 3; ++a) {
        if (nl) CU(cudaMemcpyAsync(dst4[a] + g0 + gcol_l, S.out_l[a].p, (size_t)nl * 16, cudaMemcpyDeviceToDevice, w->st));
        if (nr) CU(cudaMemcpyAsync(dst4[a] + g1 + gcol_r, S.out_r[a].p, (size_t)nr * 16, cudaMemcpyDeviceToDevice, w->st));
    }
    if (nl) CU(cudaMemcpyAsync(w->gid[c].p + g0 + gcol_l, S.gid_l.p, (size_t)nl * 4, cudaMemcpyDeviceToDevice, w->st));
    if (nr) CU(cudaMemcpyAsync(w->gid[c].p + g1 + gcol_r, S.gid_r.p, (size_t)nr * 4, cudaMemcpyDeviceToDevice, w->st));
    S.migrated_out = nl + nr;
    S.migrated_in = im_l + im_r;
    // what the sort reads, and which of its first `on` input slots it must drop
    S.sort_off = ob;
    S.sort_n = n_in;
    S.sort_dead_n = on;
    w->N = n_new;
    w->Ntot = (size_t)n_new + ghl + ghr;
    w->fluids[0].n = n_new;
    w->fluids[0].pending_delete.assign(n_new, 0);
    recompute_offsets(w);
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
    return SPH_OK;
}

// After the sort the owned range starts behind the left ghosts.  With SALVA_B200_SLAB_CHECK=1 the ranges derived from
// the exchanged counts are verified against the sorted cell table.
sph_status slab_after_sort(sph_world* w) {
    SlabState& S = w->slab;
    w->own_begin = S.gl_count;
    {   // index order of a slab world = sorted order of the moment: orig is the identity over the owned range
        const int c = w->cur;
        if (w->N) LAUNCH(k_iota_from, w->N, 256, (uint32_t)w->N, 0u, w->orig[c].p + w->own_begin);
        if (S.gl_count) CU(cudaMemsetAsync(w->orig[c].p, 0xFF, (size_t)S.gl_count * 4, w->st));
        if (S.gr_count) CU(cudaMemsetAsync(w->orig[c].p + S.gr_begin, 0xFF, (size_t)S.gr_count * 4, w->st));
    }
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
