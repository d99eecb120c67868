// sph_tile.cuh — tile-staged neighbour-gather kernels (the hot path).
//
// Why: profiles/r1_v0_* show the v0 gather passes bound by the L1 data pipe (two float4 global gathers per
// contact ≈ 10 L1 cycles each per warp).  Here a thread block owns a TILE of 2 x 2 cell columns x TILE_Z cells;
// the particles of the tile's halo (4 x 4 columns x TILE_Z+2 cells = 16 contiguous runs of the x-major sorted
// arrays) are staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), and the
// passes gather neighbour data with LDS.128 from there.  Neighbour lists hold 16-bit indices into the tile's
// halo index space (half the bytes of v0's 32-bit global indices).
//
// Halo index space of a tile: column c (0..15) occupies local indices [col_first[c], col_first[c] + count_c);
// col_first[c] = base4_c + (col_start[c] & 3) with base4_c a multiple of 4, so that 4-byte-per-particle arrays can be
// bulk-copied from a 16-byte aligned global address (col_start[c] & ~3) to a 16-byte aligned shared address and still
// use the SAME local index as the float4 arrays.
#pragma once
#include "sph_kernels.cuh"

namespace sphk {

constexpr int TILE_X = 2, TILE_Y = 2, TILE_Z = 6;
constexpr int HALO_X = TILE_X + 2, HALO_Y = TILE_Y + 2, HALO_Z = TILE_Z + 2;
constexpr int HALO_COLS = HALO_X * HALO_Y;  // 16
constexpr int OWN_COLS = TILE_X * TILE_Y;   // 4
constexpr int TILE_T = 256;                 // threads per tile block

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tile context (static shared memory) ---------------------------------------------------------
struct TileCtx {
    uint32_t col_start[HALO_COLS];  // global sorted index of the first particle of the column run
    uint32_t col_count[HALO_COLS];
    uint32_t col_first[HALO_COLS];  // local index of that particle
    uint32_t own_first[OWN_COLS];   // local index of the first OWN particle of each own column
    uint32_t own_gfirst[OWN_COLS];  // its global sorted index
    uint32_t own_prefix[OWN_COLS + 1];
    uint32_t cell_first[HALO_COLS][HALO_Z + 1];  // neighbour build only: local index of each halo cell's first particle
    uint32_t n_halo;                // size of the local index space
    int use_smem;                   // 0: halo does not fit the dynamic shared memory -> gather from global memory
    int cx0, cy0, cz0;              // grid-array coordinates of the first own cell
    alignas(8) uint64_t mbar;
};

__device__ __forceinline__ int grid_cell(int gx, int gy, int gz) { return (gx * C.ny + gy) * C.nz + gz; }

// Computes the tile's column table.  Returns the number of own particles (0 => nothing to do).
// cap = number of local slots that fit the dynamic shared memory of this launch.
__device__ __forceinline__ uint32_t tile_setup(TileCtx& T, const uint32_t* __restrict__ cstart, uint32_t cap, bool want_cells) {
    int t = blockIdx.x;
    int tz = t % C.ntz;
    int ty = (t / C.ntz) % C.nty;
    int tx = t / (C.ntz * C.nty);
    int x0 = 1 + TILE_X * tx, y0 = 1 + TILE_Y * ty, z0 = 1 + TILE_Z * tz;
    int z1 = min(z0 + TILE_Z - 1, C.nz - 2);  // last own cell (interior)
    int zlo = z0 - 1, zhi = z1 + 1;           // halo cells (padding cells exist at 0 and nz-1)
    if (threadIdx.x < HALO_COLS) {
        int c = threadIdx.x;
        int gx = x0 - 1 + c / HALO_Y, gy = y0 - 1 + c % HALO_Y;
        uint32_t s = 0, e = 0;
        if (gx < C.nx && gy < C.ny) {
            s = cstart[grid_cell(gx, gy, zlo)];
            e = cstart[grid_cell(gx, gy, zhi) + 1];
        }
        T.col_start[c] = s;
        T.col_count[c] = e - s;
        int hx = c / HALO_Y, hy = c % HALO_Y;
        if (hx >= 1 && hx <= TILE_X && hy >= 1 && hy <= TILE_Y) {
            int oc = (hx - 1) * TILE_Y + (hy - 1);
            uint32_t os = 0, oe = 0;
            if (gx <= C.nx - 2 && gy <= C.ny - 2) {
                os = cstart[grid_cell(gx, gy, z0)];
                oe = cstart[grid_cell(gx, gy, z1) + 1];
            }
            T.own_gfirst[oc] = os;
            T.own_prefix[oc + 1] = oe - os;  // counts, prefixed below
            T.own_first[oc] = os - s;        // relative to the column run, fixed up below
        }
    }
    if (threadIdx.x == 0) {
        T.cx0 = x0; T.cy0 = y0; T.cz0 = z0;
        mbar_init(&T.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = 0; c < HALO_COLS; ++c) {
            uint32_t lead = T.col_start[c] & 3u;
            T.col_first[c] = run + lead;
            run = (run + lead + T.col_count[c] + 3u) & ~3u;
        }
        T.n_halo = run;
        T.use_smem = run <= cap;
        T.own_prefix[0] = 0;
        for (int oc = 0; oc < OWN_COLS; ++oc) {
            int c = (oc / TILE_Y + 1) * HALO_Y + (oc % TILE_Y + 1);
            T.own_first[oc] += T.col_first[c];
            T.own_prefix[oc + 1] += T.own_prefix[oc];
        }
    }
    __syncthreads();
    if (want_cells && threadIdx.x < HALO_COLS * (HALO_Z + 1)) {
        int c = threadIdx.x / (HALO_Z + 1), zz = threadIdx.x % (HALO_Z + 1);
        int gx = x0 - 1 + c / HALO_Y, gy = y0 - 1 + c % HALO_Y;
        int gz = min(zlo + zz, C.nz);  // zz == HALO_Z -> one past the last halo cell
        uint32_t v = T.col_first[c] + T.col_count[c];
        if (gx < C.nx && gy < C.ny && zlo + zz <= zhi) v = T.col_first[c] + (cstart[grid_cell(gx, gy, gz)] - T.col_start[c]);
        T.cell_first[c][zz] = v;
    }
    return T.own_prefix[OWN_COLS];
}

// Stage N4 float4 arrays and N1 4-byte arrays of the tile halo into shared memory (TMA bulk copies).
// s4[a] / s1[a] point to `cap`-slot shared arrays.  All threads must call; ends with the data visible.
template <int N4, int N1>
__device__ __forceinline__ void tile_stage(TileCtx& T, const float4* const* g4, float4* const* s4, const float* const* g1, float* const* s1) {
    if (T.use_smem) {
        if (threadIdx.x == 0) {
            uint32_t bytes = 0;
            for (int c = 0; c < HALO_COLS; ++c) {
                uint32_t n = T.col_count[c];
                if (!n) continue;
                bytes += N4 * n * 16u;
                bytes += N1 * (((T.col_start[c] & 3u) + n + 3u) & ~3u) * 4u;
            }
            mbar_arrive_expect_tx(&T.mbar, bytes);
        }
        __syncthreads();
        if (threadIdx.x < HALO_COLS) {
            int c = threadIdx.x;
            uint32_t n = T.col_count[c];
            if (n) {
                uint32_t gs = T.col_start[c], lf = T.col_first[c], lead = gs & 3u;
#pragma unroll
                for (int a = 0; a < N4; ++a) bulk_g2s(s4[a] + lf, g4[a] + gs, n * 16u, &T.mbar);
#pragma unroll
                for (int a = 0; a < N1; ++a) bulk_g2s(s1[a] + (lf - lead), g1[a] + (gs - lead), ((lead + n + 3u) & ~3u) * 4u, &T.mbar);
            }
        }
        while (!mbar_try_wait(&T.mbar, 0)) {}
    }
}

// local halo index -> global sorted index (slow path for tiles that do not fit shared memory)
__device__ __forceinline__ uint32_t tile_to_global(const TileCtx& T, uint32_t l) {
    int c = 0;
#pragma unroll
    for (int k = 1; k < HALO_COLS; ++k)
        if (T.col_count[k] && l >= T.col_first[k]) c = k;
    return T.col_start[c] + (l - T.col_first[c]);
}
__device__ __forceinline__ float4 tile_get4(const TileCtx& T, const float4* s, const float4* __restrict__ g, uint32_t l) {
    return T.use_smem ? s[l] : __ldg(&g[tile_to_global(T, l)]);
}
__device__ __forceinline__ float tile_get1(const TileCtx& T, const float* s, const float* __restrict__ g, uint32_t l) {
    return T.use_smem ? s[l] : __ldg(&g[tile_to_global(T, l)]);
}

// own particle t (0 <= t < n_own) -> local index and global sorted index
__device__ __forceinline__ void tile_own(const TileCtx& T, uint32_t t, uint32_t& li, uint32_t& gi) {
    int oc = 0;
#pragma unroll
    for (int k = 1; k < OWN_COLS; ++k)
        if (t >= T.own_prefix[k]) oc = k;
    uint32_t r = t - T.own_prefix[oc];
    li = T.own_first[oc] + r;
    gi = T.own_gfirst[oc] + r;
}

struct TileLists {
    const uint16_t* nbr_f;  // nbr_f[k * stride + i]: tile-local halo index of the k-th fluid contact of i
    const uint32_t* nbr_b;  // boundary contacts: global sorted boundary index (as v0)
    const uint32_t* cnt_f;
    const uint32_t* cnt_b;
};

template <bool W, bool G, class FF>
__device__ __forceinline__ void tile_fluid_contacts(uint32_t i, const float4& pi, const TileLists& L, const TileCtx& T, const float4* s_pos,
                                                    const float4* __restrict__ g_pos, FF ff) {
    uint32_t n = min(L.cnt_f[i], C.cap_f);
    const uint16_t* col = L.nbr_f + i;
#pragma unroll 4
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t lj = col[(size_t)k * C.stride];
        float4 pj = tile_get4(T, s_pos, g_pos, lj);
        Pair p = make_pair<W, G>(pi, pj);
        ff(lj, p, pj);
    }
}
template <bool W, bool G, class FB>
__device__ __forceinline__ void tile_boundary_contacts(uint32_t i, const float4& pi, const TileLists& L, const float4* __restrict__ bpos, FB fb) {
    uint32_t n = min(L.cnt_b[i], C.cap_b);
    const uint32_t* col = L.nbr_b + i;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t j = col[(size_t)k * C.stride];
        float4 pj = __ldg(&bpos[j]);
        Pair p = make_pair<W, G>(pi, pj);
        fb(j, p, pj);
    }
}

// per-fluid deterministic error partials of a tile block: partial[block * n_fluids + f]
template <bool MULTI>
__device__ __forceinline__ void tile_reduce_error(const float* e /* [MULTI ? MAX_FLUIDS : 1] */, float* __restrict__ partial, float* sm) {
    if (!MULTI) {
        float s = block_sum(e[0], sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    } else {
        for (int f = 0; f < C.n_fluids; ++f) {
            float s = block_sum(e[f], sm);
            if (threadIdx.x == 0) partial[(size_t)blockIdx.x * C.n_fluids + f] = s;
        }
    }
}

#define TILE_SMEM_DECL                                       \
    extern __shared__ __align__(128) unsigned char tile_dyn[]; \
    __shared__ TileCtx T;                                      \
    __shared__ float red_sm[32];

// ------------------------------------------------------------------------------------------------
// K2 (tile): neighbour search contacts.rs:154-400, candidates read from the staged halo.
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(TILE_T)
k_tile_neighbors(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint32_t* __restrict__ cstart, uint32_t cap,
                 uint16_t* __restrict__ nbr_f, uint32_t* __restrict__ cnt_f, uint32_t* __restrict__ maxcnt /* [0] widest list, [2] widest halo */) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, true);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float4* s_vel = s_pos + cap;
    {
        const float4* g4[2] = {pos, vel};
        float4* s4[2] = {s_pos, s_vel};
        tile_stage<MULTI ? 2 : 1, 0>(T, g4, s4, nullptr, nullptr);
    }
    __syncthreads();  // cell_first table complete
    if (threadIdx.x == 0) atomicMax(&maxcnt[2], T.n_halo);
    uint32_t widest = 0;
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, gi;
        tile_own(T, t, li, gi);
        float4 pi = tile_get4(T, s_pos, pos, li);
        uint32_t fi = MULTI ? fid_of(tile_get4(T, s_vel, vel, li)) : 0u;
        // own column / cell of this particle inside the halo (cell coordinates as the reference: floor(x / h))
        int hx = cell_coord(pi.x) - C.ox - (T.cx0 - 1), hy = cell_coord(pi.y) - C.oy - (T.cy0 - 1), hz = cell_coord(pi.z) - C.oz - (T.cz0 - 1);
        uint32_t nf = 0;
        for (int ax = -1; ax <= 1; ++ax)
            for (int ay = -1; ay <= 1; ++ay) {
                int c = (hx + ax) * HALO_Y + (hy + ay);
                uint32_t s = T.cell_first[c][hz - 1], e = T.cell_first[c][hz + 2];
                for (uint32_t lj = s; lj < e; ++lj) {
                    float4 pj = tile_get4(T, s_pos, pos, lj);
                    float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                    bool ok = d2 <= C.h2;
                    if (MULTI && ok) {  // contacts.rs:355-362
                        uint32_t fj = fid_of(tile_get4(T, s_vel, vel, lj));
                        ok = fi == fj || groups_test(C.fluids[fi].memberships, C.fluids[fi].filter, C.fluids[fj].memberships, C.fluids[fj].filter);
                    }
                    if (ok) {
                        if (nf < C.cap_f) nbr_f[(size_t)nf * C.stride + gi] = (uint16_t)lj;
                        ++nf;
                    }
                }
            }
        cnt_f[gi] = nf;
        widest = max(widest, nf);
    }
    for (int o = 16; o > 0; o >>= 1) widest = max(widest, __shfl_xor_sync(0xffffffffu, widest, o));
    if ((threadIdx.x & 31) == 0 && widest) atomicMax(&maxcnt[0], widest);
}

// boundary contacts of fluid particles (contacts.rs:309-352), global indices as in v0
__global__ void __launch_bounds__(128)
k_neighbors_boundary(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const float4* __restrict__ bvel,
                     const uint32_t* __restrict__ bstart, uint32_t* __restrict__ nbr_b, uint32_t* __restrict__ cnt_b, uint32_t* __restrict__ maxcnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nb = 0;
    if (i < C.n_fluid) {
        if (C.n_bound) {
            float4 pi = pos[i];
            uint32_t fi = fid_of(vel[i]);
            int cx = cell_coord(pi.x), cy = cell_coord(pi.y), cz = cell_coord(pi.z);
            for (int ax = -1; ax <= 1; ++ax)
                for (int ay = -1; ay <= 1; ++ay) {
                    int base = cell_id(cx + ax, cy + ay, cz);
                    uint32_t sb = bstart[base - 1], eb = bstart[base + 2];
                    for (uint32_t j = sb; j < eb; ++j) {
                        float4 pj = __ldg(&bpos[j]);
                        float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                        if (d2 <= C.h2) {
                            uint32_t bj = fid_of(__ldg(&bvel[j]));
                            if (groups_test(C.fluids[fi].memberships, C.fluids[fi].filter, C.bounds[bj].memberships, C.bounds[bj].filter)) {
                                if (nb < C.cap_b) nbr_b[(size_t)nb * C.stride + i] = j;
                                ++nb;
                            }
                        }
                    }
                }
        }
        cnt_b[i] = nb;
    }
    uint32_t mb = nb;
    for (int o = 16; o > 0; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    if ((threadIdx.x & 31) == 0 && mb) atomicMax(&maxcnt[1], mb);
}

// ------------------------------------------------------------------------------------------------
// K3 (tile): densities + alphas (dfsph_solver.rs:628-665, 165-216, helper.rs:9-65)
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(TILE_T)
k_tile_density_alpha(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const uint32_t* __restrict__ cstart,
                     uint32_t cap, TileLists L, float* __restrict__ dens, float* __restrict__ alpha, int* __restrict__ err) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    {
        const float4* g4[1] = {pos};
        float4* s4[1] = {s_pos};
        tile_stage<1, 0>(T, g4, s4, nullptr, nullptr);
    }
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        float4 pi = tile_get4(T, s_pos, pos, li);
        float rho0 = C.fluids[MULTI ? fid_of(vel[i]) : 0].density0;
        float rho = 0.f, sq = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        tile_fluid_contacts<true, true>(i, pi, L, T, s_pos, pos, [&](uint32_t, const Pair& p, const float4& pj) {
            rho = fmaf(pj.w, p.w, rho);
            float s = p.g * pj.w;
            float ax = s * p.dx, ay = s * p.dy, az = s * p.dz;
            sq += ax * ax + ay * ay + az * az;
            gx += ax; gy += ay; gz += az;
        });
        tile_boundary_contacts<true, true>(i, pi, L, bpos, [&](uint32_t, const Pair& p, const float4& pj) {
            float mb = pj.w * rho0;
            rho = fmaf(mb, p.w, rho);
            float s = p.g * mb;
            float ax = s * p.dx, ay = s * p.dy, az = s * p.dz;
            sq += ax * ax + ay * ay + az * az;
            gx += ax; gy += ay; gz += az;
        });
        if (rho == 0.f) atomicOr(err, 1);
        float den = sq + (gx * gx + gy * gy + gz * gz);
        dens[i] = rho;
        alpha[i] = den <= 1.0e-5f ? 0.f : 1.0f / den;
    }
}

// ------------------------------------------------------------------------------------------------
// K4a / K8a (tile): compute_divergences dfsph_solver.rs:279-356 and compute_predicted_densities :98-162
// share one kernel: sum_j m_j (v*_i - v*_j) . gradW_ij  (+ boundary term with / without boundary velocity).
// ------------------------------------------------------------------------------------------------
template <bool MULTI, bool PREDICT>
__global__ void __launch_bounds__(TILE_T)
k_tile_vel_divergence(const float4* __restrict__ pos, const float4* __restrict__ vs, const float4* __restrict__ vel, const float4* __restrict__ bpos,
                      const float4* __restrict__ bvel, const uint32_t* __restrict__ cstart, uint32_t cap, TileLists L, const float* __restrict__ dens,
                      const float* __restrict__ alpha, float* __restrict__ out /* divv or pred */, float* __restrict__ kappa, float* __restrict__ partial,
                      float dt, int* __restrict__ err) {
    TILE_SMEM_DECL
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    float e[MULTI ? MAX_FLUIDS : 1];
#pragma unroll
    for (int f = 0; f < (MULTI ? MAX_FLUIDS : 1); ++f) e[f] = 0.f;
    if (n_own) {
        float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
        float4* s_vs = s_pos + cap;
        {
            const float4* g4[2] = {pos, vs};
            float4* s4[2] = {s_pos, s_vs};
            tile_stage<2, 0>(T, g4, s4, nullptr, nullptr);
        }
        for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
            uint32_t li, i;
            tile_own(T, t, li, i);
            float4 pi = tile_get4(T, s_pos, pos, li);
            float4 vi = tile_get4(T, s_vs, vs, li);
            uint32_t fi = MULTI ? fid_of(vel[i]) : 0u;
            float rho0 = C.fluids[fi].density0;
            float d = 0.f;
            if (PREDICT || L.cnt_f[i] + L.cnt_b[i] >= 20u) {  // min_neighbors_for_divergence_solve :62,301-314
                tile_fluid_contacts<false, true>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
                    float4 vj = tile_get4(T, s_vs, vs, lj);
                    float dv = (vi.x - vj.x) * p.dx + (vi.y - vj.y) * p.dy + (vi.z - vj.z) * p.dz;
                    d = fmaf(dv * p.g, pj.w, d);
                });
                tile_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
                    float dv;
                    if (PREDICT) {  // :136-141 uses the boundary velocity, the divergence does not (:336-338)
                        float4 vj = __ldg(&bvel[j]);
                        dv = (vi.x - vj.x) * p.dx + (vi.y - vj.y) * p.dy + (vi.z - vj.z) * p.dz;
                    } else {
                        dv = vi.x * p.dx + vi.y * p.dy + vi.z * p.dz;
                    }
                    d = fmaf(dv * p.g, pj.w * rho0, d);
                });
            }
            float ee;
            if (PREDICT) {
                float pd = fmaf(d, dt, dens[i]);
                if (pd == 0.f) atomicOr(err, 1);  // assert :145
                out[i] = pd;
                kappa[i] = fmaxf((pd - rho0) * alpha[i], 0.f);
                ee = pd < rho0 ? 0.f : pd / rho0 - 1.0f;
            } else {
                d = fmaxf(d, 0.f);
                out[i] = d;
                kappa[i] = d * alpha[i];
                ee = d / rho0;
            }
            if (MULTI) {
#pragma unroll
                for (int f = 0; f < MAX_FLUIDS; ++f)
                    if (fi == (uint32_t)f) e[f] += ee;
            } else {
                e[0] += ee;
            }
        }
    }
    tile_reduce_error<MULTI>(e, partial, red_sm);
}

// ------------------------------------------------------------------------------------------------
// K4b / K8b (tile): compute_velocity_changes_for_divergence :358-409 and compute_velocity_changes :218-277:
// vc_i -= scale * sum_j (k_i + k_j) m_j gradW_ij (+ boundary term), v* = vel + vc.
// PRESSURE: k = kappa+ (>= 0), scale = inv_dt, boundary term only if k_i > 0; else k = div*alpha, scale = 1.
// ------------------------------------------------------------------------------------------------
template <bool MULTI, bool BFORCE, bool PRESSURE>
__global__ void __launch_bounds__(TILE_T)
k_tile_vel_update(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const uint32_t* __restrict__ cstart,
                  uint32_t cap, TileLists L, const float* __restrict__ kappa, float4* __restrict__ vc, float4* __restrict__ vs, float* __restrict__ bforce,
                  float inv_dt) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float* s_k = reinterpret_cast<float*>(s_pos + cap);
    {
        const float4* g4[1] = {pos};
        float4* s4[1] = {s_pos};
        const float* g1[1] = {kappa};
        float* s1[1] = {s_k};
        tile_stage<1, 1>(T, g4, s4, g1, s1);
    }
    const float scale = PRESSURE ? inv_dt : 1.0f;
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        float4 pi = tile_get4(T, s_pos, pos, li);
        float ki = tile_get1(T, s_k, kappa, li);
        float4 v = vel[i];
        float rho0 = C.fluids[MULTI ? fid_of(v) : 0].density0;
        float ax = 0.f, ay = 0.f, az = 0.f;
        tile_fluid_contacts<false, true>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
            float c = (ki + tile_get1(T, s_k, kappa, lj)) * pj.w * scale * p.g;
            ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
        });
        if (!PRESSURE || ki > 0.f) {  // :257
            tile_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
                float c = ki * pj.w * rho0 * scale * p.g;
                ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
                if (BFORCE) {  // :269-272 / :403-405: both reduce to +c * inv_dt * m_i * x_ij on the boundary particle
                    float s = c * inv_dt * pi.w;
                    atomicAdd(&bforce[3 * (size_t)j + 0], s * p.dx);
                    atomicAdd(&bforce[3 * (size_t)j + 1], s * p.dy);
                    atomicAdd(&bforce[3 * (size_t)j + 2], s * p.dz);
                }
            });
        }
        float4 c4 = vc[i];
        c4.x -= ax; c4.y -= ay; c4.z -= az;
        vc[i] = c4;
        vs[i] = make_float4(v.x + c4.x, v.y + c4.y, v.z + c4.z, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// Nonpressure forces (tile): same-fluid contacts only.
// ------------------------------------------------------------------------------------------------
// a12 XSPHViscosity::solve xsph_viscosity.rs:30-95
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(TILE_T)
k_tile_xsph(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const float4* __restrict__ bvel,
            const uint32_t* __restrict__ cstart, uint32_t cap, TileLists L, const float* __restrict__ dens, float4* __restrict__ acc,
            float* __restrict__ bforce, uint32_t which, float cf, float cb, float inv_dt) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float4* s_vel = s_pos + cap;
    float* s_rho = reinterpret_cast<float*>(s_vel + cap);
    {
        const float4* g4[2] = {pos, vel};
        float4* s4[2] = {s_pos, s_vel};
        const float* g1[1] = {dens};
        float* s1[1] = {s_rho};
        tile_stage<2, 1>(T, g4, s4, g1, s1);
    }
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        float4 vi = tile_get4(T, s_vel, vel, li);
        if (MULTI && fid_of(vi) != which) continue;
        float4 pi = tile_get4(T, s_pos, pos, li);
        float rho0 = C.fluids[which].density0;
        float fx = 0.f, fy = 0.f, fz = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        if (cf != 0.f)
            tile_fluid_contacts<true, false>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
                float4 vj = tile_get4(T, s_vel, vel, lj);
                if (MULTI && fid_of(vj) != which) return;
                float c = cf * p.w * pj.w / tile_get1(T, s_rho, dens, lj);
                fx = fmaf(c, vj.x - vi.x, fx); fy = fmaf(c, vj.y - vi.y, fy); fz = fmaf(c, vj.z - vi.z, fz);
            });
        if (cb != 0.f) {
            float rho_i = tile_get1(T, s_rho, dens, li);
            tile_boundary_contacts<true, false>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
                float4 vj = __ldg(&bvel[j]);
                float c = cb * p.w * pj.w * rho0 / rho_i;
                float dx = c * (vj.x - vi.x), dy = c * (vj.y - vi.y), dz = c * (vj.z - vi.z);
                bx += dx; by += dy; bz += dz;
                if (BFORCE) {
                    float s = -pi.w * inv_dt;
                    atomicAdd(&bforce[3 * (size_t)j + 0], s * dx);
                    atomicAdd(&bforce[3 * (size_t)j + 1], s * dy);
                    atomicAdd(&bforce[3 * (size_t)j + 2], s * dz);
                }
            });
        }
        float4 a = acc[i];
        a.x += fx * inv_dt + bx * inv_dt; a.y += fy * inv_dt + by * inv_dt; a.z += fz * inv_dt + bz * inv_dt;
        acc[i] = a;
    }
}

// a13 ArtificialViscosity::solve artificial_viscosity.rs:40-124
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(TILE_T)
k_tile_artificial(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const float4* __restrict__ bvel,
                  const uint32_t* __restrict__ cstart, uint32_t cap, TileLists L, const float* __restrict__ dens, float4* __restrict__ acc,
                  float* __restrict__ bforce, uint32_t which, float cf, float cb, float alpha, float beta, float cs) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float4* s_vel = s_pos + cap;
    float* s_rho = reinterpret_cast<float*>(s_vel + cap);
    {
        const float4* g4[2] = {pos, vel};
        float4* s4[2] = {s_pos, s_vel};
        const float* g1[1] = {dens};
        float* s1[1] = {s_rho};
        tile_stage<2, 1>(T, g4, s4, g1, s1);
    }
    const float eta2 = C.h * C.h * 0.01f;
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        float4 vi = tile_get4(T, s_vel, vel, li);
        if (MULTI && fid_of(vi) != which) continue;
        float4 pi = tile_get4(T, s_pos, pos, li);
        float rho0 = C.fluids[which].density0;
        float rho_i = tile_get1(T, s_rho, dens, li);
        float fx = 0.f, fy = 0.f, fz = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        if (cf != 0.f)
            tile_fluid_contacts<false, true>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
                float4 vj = tile_get4(T, s_vel, vel, lj);
                if (MULTI && fid_of(vj) != which) return;
                float vr = p.dx * (vi.x - vj.x) + p.dy * (vi.y - vj.y) + p.dz * (vi.z - vj.z);
                if (vr < 0.f) {
                    float davg = (rho_i + tile_get1(T, s_rho, dens, lj)) * 0.5f;
                    float mu = C.h * vr / (p.d2 + eta2);
                    float c = cf * (cs * alpha * mu - beta * mu * mu) * (pj.w / davg) * p.g;
                    fx = fmaf(c, p.dx, fx); fy = fmaf(c, p.dy, fy); fz = fmaf(c, p.dz, fz);
                }
            });
        if (cb != 0.f)
            tile_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
                float4 vj = __ldg(&bvel[j]);
                float vr = p.dx * (vi.x - vj.x) + p.dy * (vi.y - vj.y) + p.dz * (vi.z - vj.z);
                if (vr < 0.f) {
                    float mu = C.h * vr / (p.d2 + eta2);
                    float c = cb * (cs * alpha * mu - beta * mu * mu) * (pj.w * rho0 / rho_i) * p.g;
                    bx = fmaf(c, p.dx, bx); by = fmaf(c, p.dy, by); bz = fmaf(c, p.dz, bz);
                    if (BFORCE) {  // running sum, as the reference (:117)
                        atomicAdd(&bforce[3 * (size_t)j + 0], -pi.w * bx);
                        atomicAdd(&bforce[3 * (size_t)j + 1], -pi.w * by);
                        atomicAdd(&bforce[3 * (size_t)j + 2], -pi.w * bz);
                    }
                }
            });
        float4 a = acc[i];
        a.x += fx + bx; a.y += fy + by; a.z += fz + bz;
        acc[i] = a;
    }
}

// a14 pass 1: compute_normals akinci2013_surface_tension.rs:43-68
template <bool MULTI>
__global__ void __launch_bounds__(TILE_T)
k_tile_akinci_normals(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint32_t* __restrict__ cstart, uint32_t cap, TileLists L,
                      const float* __restrict__ dens, float4* __restrict__ normals, uint32_t which) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float4* s_vel = s_pos + cap;
    float* s_rho = reinterpret_cast<float*>(s_pos + (MULTI ? 2 : 1) * (size_t)cap);
    {
        const float4* g4[2] = {pos, vel};
        float4* s4[2] = {s_pos, s_vel};
        const float* g1[1] = {dens};
        float* s1[1] = {s_rho};
        tile_stage<MULTI ? 2 : 1, 1>(T, g4, s4, g1, s1);
    }
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        if (MULTI && fid_of(tile_get4(T, s_vel, vel, li)) != which) continue;
        float4 pi = tile_get4(T, s_pos, pos, li);
        float nx = 0.f, ny = 0.f, nz = 0.f;
        tile_fluid_contacts<false, true>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
            if (MULTI && fid_of(tile_get4(T, s_vel, vel, lj)) != which) return;
            float c = p.g * (pj.w / tile_get1(T, s_rho, dens, lj));
            nx = fmaf(c, p.dx, nx); ny = fmaf(c, p.dy, ny); nz = fmaf(c, p.dz, nz);
        });
        normals[i] = make_float4(nx * C.h, ny * C.h, nz * C.h, 0.f);
    }
}

// a14 pass 2: Akinci2013SurfaceTension::solve akinci2013_surface_tension.rs:113-192
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(TILE_T)
k_tile_akinci_force(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const uint32_t* __restrict__ cstart,
                    uint32_t cap, TileLists L, const float* __restrict__ dens, const float4* __restrict__ normals, float4* __restrict__ acc,
                    float* __restrict__ bforce, uint32_t which, float gamma, float adh, float coh_norm, float h6_64, float adh_norm) {
    TILE_SMEM_DECL
    (void)red_sm;
    uint32_t n_own = tile_setup(T, cstart, cap, false);
    if (n_own == 0) return;
    float4* s_pos = reinterpret_cast<float4*>(tile_dyn);
    float4* s_nrm = s_pos + cap;
    float4* s_vel = s_nrm + cap;
    float* s_rho = reinterpret_cast<float*>(s_pos + (MULTI ? 3 : 2) * (size_t)cap);
    {
        const float4* g4[3] = {pos, normals, vel};
        float4* s4[3] = {s_pos, s_nrm, s_vel};
        const float* g1[1] = {dens};
        float* s1[1] = {s_rho};
        tile_stage<MULTI ? 3 : 2, 1>(T, g4, s4, g1, s1);
    }
    for (uint32_t t = threadIdx.x; t < n_own; t += blockDim.x) {
        uint32_t li, i;
        tile_own(T, t, li, i);
        if (MULTI && fid_of(tile_get4(T, s_vel, vel, li)) != which) continue;
        float4 pi = tile_get4(T, s_pos, pos, li);
        float rho0 = C.fluids[which].density0;
        float rho_i = tile_get1(T, s_rho, dens, li);
        float4 ni = tile_get4(T, s_nrm, normals, li);
        float ax = 0.f, ay = 0.f, az = 0.f;
        if (gamma != 0.f)
            tile_fluid_contacts<false, false>(i, pi, L, T, s_pos, pos, [&](uint32_t lj, const Pair& p, const float4& pj) {
                if (MULTI && fid_of(tile_get4(T, s_vel, vel, lj)) != which) return;
                float4 nj = tile_get4(T, s_nrm, normals, lj);
                float coh = p.d2 > F32_EPS * F32_EPS ? cohesion_kernel(p.r, coh_norm, h6_64) / p.r : 0.f;
                float cm = coh * (-gamma * pj.w);
                float kij = 2.0f * rho0 / (rho_i + tile_get1(T, s_rho, dens, lj));
                ax += (-gamma * (ni.x - nj.x) + cm * p.dx) * kij;
                ay += (-gamma * (ni.y - nj.y) + cm * p.dy) * kij;
                az += (-gamma * (ni.z - nj.z) + cm * p.dz) * kij;
            });
        if (adh != 0.f)
            tile_boundary_contacts<false, false>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
                float ad = p.d2 > F32_EPS * F32_EPS ? adhesion_kernel(p.r, adh_norm) / p.r : 0.f;
                float c = ad * adh * (pj.w * rho0);
                ax -= c * p.dx; ay -= c * p.dy; az -= c * p.dz;
                if (BFORCE) {
                    atomicAdd(&bforce[3 * (size_t)j + 0], c * p.dx * pi.w);
                    atomicAdd(&bforce[3 * (size_t)j + 1], c * p.dy * pi.w);
                    atomicAdd(&bforce[3 * (size_t)j + 2], c * p.dz * pi.w);
                }
            });
        float4 a = acc[i];
        a.x += ax; a.y += ay; a.z += az;
        acc[i] = a;
    }
}

}  // namespace sphk
