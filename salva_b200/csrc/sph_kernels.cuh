// sph_kernels.cuh — sm_100a CUDA kernels of the SPH step path.
//
// Data layout (all per-particle arrays are in SORTED order: x-major cell order, z fastest, so the
// 27-cell stencil of a particle is 9 contiguous runs of the sorted arrays):
//   pos4  : x, y, z, mass          (mass = volume * density0 of the particle's fluid, fluid.rs:183-185)
//   vel4  : vx, vy, vz, fluid id   (bit pattern of the fluid index in .w)
//   vc4   : velocity_changes       (dfsph_solver.rs:44)
//   vs4   : v* = vel + vc          (materialised so gather passes read ONE vector per neighbour)
//   bpos4 : boundary x, y, z, volume (dfsph_solver.rs:72-96);  bvel4: boundary velocity, boundary id
// Neighbour lists ("contacts", contacts.rs:83-87) are index-only and column-major:
//   nbr_f[k * stride + i] = sorted index of the k-th fluid neighbour of i (self included, ascending j),
//   nbr_b[k * stride + i] likewise for boundary particles; W and grad W are recomputed from pos4 in
//   every pass (cheaper than streaming cached 16-byte contacts from HBM: see DESIGN.md).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sphk {

constexpr int MAX_FLUIDS = 16;
constexpr int MAX_BOUNDARIES = 64;
constexpr float F32_EPS = 1.1920929e-07f;

struct FluidParams {  // per fluid, in __constant__ memory
    float density0;
    uint32_t memberships, filter;
    uint32_t count;
};
struct BoundaryParams {
    uint32_t memberships, filter;
};
struct Consts {
    float h, inv_h, h2;          // h2 = h*h rounded once (contacts.rs:285 `h * h`)
    float sigma;                 // 8 / (pi h^3)            cubic_spline_kernel.rs:18
    float dsigma;                // sigma / h               cubic_spline_kernel.rs:79
    int ox, oy, oz;              // grid origin in cell coordinates (one padding cell each side)
    int nx, ny, nz;
    uint32_t n_fluid, n_bound;   // particle totals
    uint32_t stride;             // neighbour-list column stride (>= n_fluid, multiple of 32)
    uint32_t cap_f, cap_b;       // list capacities (rows)
    int n_fluids, n_bounds;      // object counts
    FluidParams fluids[MAX_FLUIDS];
    BoundaryParams bounds[MAX_BOUNDARIES];
};

__constant__ Consts C;

// ------------------------------------------------------------------------------------------------
// geometry helpers
// ------------------------------------------------------------------------------------------------
// hgrid.rs:41-52: cell = floor(x / h), IEEE division exactly as the reference.
__device__ __forceinline__ int cell_coord(float x) { return (int)floorf(__fdiv_rn(x, C.h)); }

__device__ __forceinline__ int cell_id(int cx, int cy, int cz) { return ((cx - C.ox) * C.ny + (cy - C.oy)) * C.nz + (cz - C.oz); }

// contacts.rs:285,322,366: (dx*dx + dy*dy) + dz*dz <= h*h with no contraction (rustc never fuses).
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// cubic_spline_kernel.rs:12-33: W(r) for q = r/h.
__device__ __forceinline__ float kernel_w(float r) {
    float q = r * C.inv_h;
    float q2 = q * q;
    float a = 1.0f + (q2 * q - q2) * 6.0f;
    float t = 1.0f - q;
    float b = (t * t * t) * 2.0f;
    float rhs = q <= 0.5f ? a : (q <= 1.0f ? b : 0.0f);
    return C.sigma * rhs;
}
// cubic_spline_kernel.rs:55-80 + kernel.rs:18-24: returns g with grad W_ij = g * (x_i - x_j);
// zero if |x_ij|^2 <= eps^2, q <= 1e-5 or q > 1.
__device__ __forceinline__ float kernel_gfac(float d2, float r, float inv_r) {
    float q = r * C.inv_h;
    float t = 1.0f - q;
    float a = (q * 3.0f - 2.0f) * q * 6.0f;
    float b = -t * t * 6.0f;
    float rhs = q <= 0.5f ? a : b;
    bool zero = (q > 1.0f) | (q <= 1.0e-5f) | !(d2 > F32_EPS * F32_EPS);
    return zero ? 0.0f : C.dsigma * rhs * inv_r;
}

struct Pair {        // geometry of one (i, j) contact
    float dx, dy, dz;  // x_i - x_j
    float d2, r;
    float w;           // contact.weight
    float g;           // contact.gradient = g * (dx, dy, dz)
};
template <bool NEED_W, bool NEED_G>
__device__ __forceinline__ Pair make_pair(const float4& pi, const float4& pj) {
    Pair p;
    p.dx = pi.x - pj.x;
    p.dy = pi.y - pj.y;
    p.dz = pi.z - pj.z;
    p.d2 = fmaf(p.dz, p.dz, fmaf(p.dy, p.dy, p.dx * p.dx));
    float inv_r = rsqrtf(fmaxf(p.d2, 1.0e-30f));
    p.r = p.d2 * inv_r;
    p.w = NEED_W ? kernel_w(p.r) : 0.f;
    p.g = NEED_G ? kernel_gfac(p.d2, p.r, inv_r) : 0.f;
    return p;
}

__device__ __forceinline__ uint32_t fid_of(const float4& v) { return __float_as_uint(v.w); }

// interaction_groups.rs:64-69
__device__ __forceinline__ bool groups_test(uint32_t m1, uint32_t f1, uint32_t m2, uint32_t f2) {
    return (m1 & f2) != 0 && (m2 & f1) != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Deterministic block sum (fixed tree); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* sm /* >= 32 floats */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? sm[lane] : 0.f;
        v = warp_sum(v);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// K0: bounds (cell-coordinate AABB) — replaces the unbounded HashMap of hgrid.rs:22-25 by a dense
// grid over the occupied region.
// ------------------------------------------------------------------------------------------------
__global__ void k_bounds(const float4* __restrict__ pos, uint32_t n, int* __restrict__ out /* minx,miny,minz,maxx,maxy,maxz,bad */) {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    int bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pos[i];
        float c[3] = {floorf(__fdiv_rn(p.x, C.h)), floorf(__fdiv_rn(p.y, C.h)), floorf(__fdiv_rn(p.z, C.h))};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(fabsf(c[a]) < 1.0e9f)) { bad = 1; continue; }  // NaN / inf / absurd coordinates
            int v = (int)c[a];
            mn[a] = min(mn[a], v);
            mx[a] = max(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&out[a], mn[a]);
            atomicMax(&out[3 + a], mx[a]);
        }
        if (bad) atomicOr(&out[6], 1);
    }
}

// ------------------------------------------------------------------------------------------------
// K1: counting sort by cell (replaces HGrid::insert hgrid.rs:60-63 / insert_*_to_grid contacts.rs:133-151
// and the dead z_order.rs sort).
// ------------------------------------------------------------------------------------------------
__global__ void k_cell_hist(const float4* __restrict__ pos, uint32_t n, uint32_t* __restrict__ cid, uint32_t* __restrict__ rank,
                            uint32_t* __restrict__ count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    uint32_t id = (uint32_t)cell_id(cell_coord(p.x), cell_coord(p.y), cell_coord(p.z));
    cid[i] = id;
    rank[i] = atomicAdd(&count[id], 1u);
}

__global__ void k_cell_scatter(uint32_t n, const uint32_t* __restrict__ cid, const uint32_t* __restrict__ rank,
                               const uint32_t* __restrict__ start, uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    perm[start[cid[i]] + rank[i]] = i;
}

// Deterministic mode: atomics hand out in-cell ranks in arbitrary order; sort each cell's slice of perm
// ascending so the sorted order (and every f32 summation order downstream) is reproducible.
__global__ void k_cell_sort(uint32_t ncell, const uint32_t* __restrict__ start, uint32_t* __restrict__ perm) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    uint32_t s = start[c], e = start[c + 1];
    for (uint32_t a = s + 1; a < e; ++a) {
        uint32_t v = perm[a];
        uint32_t b = a;
        while (b > s && perm[b - 1] > v) {
            perm[b] = perm[b - 1];
            --b;
        }
        perm[b] = v;
    }
}

struct GatherSet {  // arrays reordered together by the counting sort
    const float4* in4[6];
    float4* out4[6];
    const uint32_t* in1[4];
    uint32_t* out1[4];
    int n4, n1;
};
__global__ void k_gather(uint32_t n, const uint32_t* __restrict__ perm, GatherSet g) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t src = perm[s];
#pragma unroll
    for (int a = 0; a < 6; ++a)
        if (a < g.n4) g.out4[a][s] = g.in4[a][src];
#pragma unroll
    for (int a = 0; a < 4; ++a)
        if (a < g.n1) g.out1[a][s] = g.in1[a][src];
}

// exclusive scan, 2048 items per block (256 threads x 8)
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;
__global__ void k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_tot[SCAN_T / 32];
    uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        tsum += v[k];
    }
    uint32_t incl = tsum;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_T / 32 ? warp_tot[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_T / 32) warp_tot[lane] = wi - w;
        if (lane == SCAN_T / 32 - 1 && block_sums) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_tot[wid] + incl - tsum;
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}
__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ block_offsets) {
    uint32_t i = blockIdx.x * SCAN_B + threadIdx.x;
    uint32_t off = block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        uint32_t j = i + k * SCAN_T;
        if (j < n) out[j] += off;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: neighbour search (contacts.rs:154-400).  One thread per particle walks the 9 z-runs of its
// 27-cell stencil and keeps the indices that pass the reference's exact `d^2 <= h*h` test.
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(128)
k_neighbors(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint32_t* __restrict__ cstart,
            const float4* __restrict__ bpos, const float4* __restrict__ bvel, const uint32_t* __restrict__ bstart,
            uint32_t* __restrict__ nbr_f, uint32_t* __restrict__ nbr_b, uint32_t* __restrict__ cnt_f, uint32_t* __restrict__ cnt_b,
            uint32_t* __restrict__ maxcnt /* [0]=fluid,[1]=boundary */) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nf = 0, nb = 0;
    if (i < C.n_fluid) {
        float4 pi = pos[i];
        uint32_t fi = MULTI ? fid_of(vel[i]) : 0u;
        int cx = cell_coord(pi.x), cy = cell_coord(pi.y), cz = cell_coord(pi.z);
        for (int ax = -1; ax <= 1; ++ax)
            for (int ay = -1; ay <= 1; ++ay) {
                int base = cell_id(cx + ax, cy + ay, cz);
                uint32_t s = cstart[base - 1], e = cstart[base + 2];
                for (uint32_t j = s; j < e; ++j) {
                    float4 pj = __ldg(&pos[j]);
                    float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                    bool ok = d2 <= C.h2;
                    if (MULTI && ok) {  // contacts.rs:355-362: different fluids need the groups test
                        uint32_t fj = fid_of(__ldg(&vel[j]));
                        ok = fi == fj || groups_test(C.fluids[fi].memberships, C.fluids[fi].filter, C.fluids[fj].memberships,
                                                     C.fluids[fj].filter);
                    }
                    if (ok) {
                        if (nf < C.cap_f) nbr_f[(size_t)nf * C.stride + i] = j;
                        ++nf;
                    }
                }
                if (C.n_bound) {
                    uint32_t sb = bstart[base - 1], eb = bstart[base + 2];
                    for (uint32_t j = sb; j < eb; ++j) {
                        float4 pj = __ldg(&bpos[j]);
                        float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                        bool ok = d2 <= C.h2;
                        if (ok) {  // contacts.rs:347-352
                            uint32_t bj = fid_of(__ldg(&bvel[j]));
                            ok = groups_test(C.fluids[fi].memberships, C.fluids[fi].filter, C.bounds[bj].memberships, C.bounds[bj].filter);
                        }
                        if (ok) {
                            if (nb < C.cap_b) nbr_b[(size_t)nb * C.stride + i] = j;
                            ++nb;
                        }
                    }
                }
            }
        cnt_f[i] = nf;
        cnt_b[i] = nb;
    }
    uint32_t mf = nf, mb = nb;
    for (int o = 16; o > 0; o >>= 1) {
        mf = max(mf, __shfl_xor_sync(0xffffffffu, mf, o));
        mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mf) atomicMax(&maxcnt[0], mf);
        if (mb) atomicMax(&maxcnt[1], mb);
    }
}

// a4: compute_boundary_volumes dfsph_solver.rs:72-96 — vol_b = 1 / sum_{b'} W_bb' over boundary-boundary
// contacts (same boundary, or other boundaries passing the groups test; self included).
__global__ void __launch_bounds__(128)
k_boundary_volumes(const float4* __restrict__ bpos, const float4* __restrict__ bvel, const uint32_t* __restrict__ bstart, float* __restrict__ bvol,
                   unsigned long long* __restrict__ ncontacts, int* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (i < C.n_bound) {
        float4 pi = bpos[i];
        uint32_t bi = fid_of(bvel[i]);
        int cx = cell_coord(pi.x), cy = cell_coord(pi.y), cz = cell_coord(pi.z);
        float den = 0.f;
        for (int ax = -1; ax <= 1; ++ax)
            for (int ay = -1; ay <= 1; ++ay) {
                int base = cell_id(cx + ax, cy + ay, cz);
                uint32_t s = bstart[base - 1], e = bstart[base + 2];
                for (uint32_t j = s; j < e; ++j) {
                    float4 pj = __ldg(&bpos[j]);
                    float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    float d2 = dist2_exact(dx, dy, dz);
                    if (d2 <= C.h2) {
                        uint32_t bj = fid_of(__ldg(&bvel[j]));
                        if (bi == bj || groups_test(C.bounds[bi].memberships, C.bounds[bi].filter, C.bounds[bj].memberships, C.bounds[bj].filter)) {
                            den += kernel_w(sqrtf(d2));
                            ++cnt;
                        }
                    }
                }
            }
        if (den == 0.f) atomicOr(err, 1);  // assert!(!denominator.is_zero()) dfsph_solver.rs:92
        bvol[i] = 1.0f / den;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(ncontacts, (unsigned long long)cnt);
}
__global__ void k_set_w(uint32_t n, float4* __restrict__ a, const float* __restrict__ w) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i].w = w[i];
}

// ------------------------------------------------------------------------------------------------
// Neighbour-gather pass skeleton: calls ff(j, pair, pj) for every fluid contact and fb(j, pair, pj) for
// every boundary contact of particle i.
// ------------------------------------------------------------------------------------------------
struct Lists {
    const uint32_t* nbr_f;
    const uint32_t* nbr_b;
    const uint32_t* cnt_f;
    const uint32_t* cnt_b;
};

template <bool W, bool G, class FF>
__device__ __forceinline__ void for_fluid_contacts(uint32_t i, const float4& pi, const Lists& L, const float4* __restrict__ pos, FF ff) {
    uint32_t n = min(L.cnt_f[i], C.cap_f);
    const uint32_t* col = L.nbr_f + i;
#pragma unroll 4
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t j = col[(size_t)k * C.stride];
        float4 pj = __ldg(&pos[j]);
        Pair p = make_pair<W, G>(pi, pj);
        ff(j, p, pj);
    }
}
template <bool W, bool G, class FB>
__device__ __forceinline__ void for_boundary_contacts(uint32_t i, const float4& pi, const Lists& L, const float4* __restrict__ bpos, FB fb) {
    uint32_t n = min(L.cnt_b[i], C.cap_b);
    const uint32_t* col = L.nbr_b + i;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t j = col[(size_t)k * C.stride];
        float4 pj = __ldg(&bpos[j]);
        Pair p = make_pair<W, G>(pi, pj);
        fb(j, p, pj);
    }
}

// Per-fluid deterministic error reduction: partial[block * n_fluids + f].
template <bool MULTI>
__device__ __forceinline__ void reduce_error(float e, uint32_t fi, bool valid, float* __restrict__ partial, float* sm) {
    if (!MULTI) {
        float s = block_sum(valid ? e : 0.f, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    } else {
        for (int f = 0; f < C.n_fluids; ++f) {
            float s = block_sum((valid && fi == (uint32_t)f) ? e : 0.f, sm);
            if (threadIdx.x == 0) partial[(size_t)blockIdx.x * C.n_fluids + f] = s;
        }
    }
}
// One block per fluid: fixed-order sum of the per-block partials.
__global__ void k_reduce_partials(const float* __restrict__ partial, uint32_t nblocks, int n_fluids, float* __restrict__ out) {
    __shared__ float sm[32];
    int f = blockIdx.x;
    float s = 0.f;
    for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t)b * n_fluids + f];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[f] = s;
}

constexpr int PASS_T = 128;  // threads per block of the gather passes

// ------------------------------------------------------------------------------------------------
// K3: densities (dfsph_solver.rs:628-665) fused with alphas (dfsph_solver.rs:165-216) and the per-contact
// kernel evaluation of helper.rs:9-65.
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(PASS_T)
k_density_alpha(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, Lists L,
                float* __restrict__ dens, float* __restrict__ alpha, int* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 pi = pos[i];
    float rho0 = C.fluids[MULTI ? fid_of(vel[i]) : 0].density0;
    float rho = 0.f, sq = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    for_fluid_contacts<true, true>(i, pi, L, pos, [&](uint32_t, const Pair& p, const float4& pj) {
        rho = fmaf(pj.w, p.w, rho);
        float s = p.g * pj.w;  // m_j * gradient
        float ax = s * p.dx, ay = s * p.dy, az = s * p.dz;
        sq += ax * ax + ay * ay + az * az;
        gx += ax; gy += ay; gz += az;
    });
    for_boundary_contacts<true, true>(i, pi, L, bpos, [&](uint32_t, const Pair& p, const float4& pj) {
        float mb = pj.w * rho0;  // boundary pseudo mass: vol_b * rho0_i
        rho = fmaf(mb, p.w, rho);
        float s = p.g * mb;
        float ax = s * p.dx, ay = s * p.dy, az = s * p.dz;
        sq += ax * ax + ay * ay + az * az;
        gx += ax; gy += ay; gz += az;
    });
    if (rho == 0.f) atomicOr(err, 1);  // assert!(!density.is_zero()) dfsph_solver.rs:662
    float den = sq + (gx * gx + gy * gy + gz * gz);
    dens[i] = rho;
    alpha[i] = den <= 1.0e-5f ? 0.f : 1.0f / den;  // dfsph_solver.rs:209-213
}

// ------------------------------------------------------------------------------------------------
// K4a: compute_divergences dfsph_solver.rs:279-356.  Writes div_i and kdiv_i = div_i * alpha_i.
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(PASS_T)
k_divergence(const float4* __restrict__ pos, const float4* __restrict__ vs, const float4* __restrict__ vel, const float4* __restrict__ bpos, Lists L,
             const float* __restrict__ alpha, float* __restrict__ divv, float* __restrict__ kappa, float* __restrict__ partial) {
    __shared__ float sm[32];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < C.n_fluid;
    float e = 0.f;
    uint32_t fi = 0;
    if (valid) {
        float4 pi = pos[i];
        float4 vi = vs[i];
        fi = MULTI ? fid_of(vel[i]) : 0u;
        float rho0 = C.fluids[fi].density0;
        float d = 0.f;
        if (L.cnt_f[i] + L.cnt_b[i] >= 20u) {  // min_neighbors_for_divergence_solve dfsph_solver.rs:62,301-314
            for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
                float4 vj = __ldg(&vs[j]);
                float dv = (vi.x - vj.x) * p.dx + (vi.y - vj.y) * p.dy + (vi.z - vj.z) * p.dz;
                d = fmaf(dv * p.g, pj.w, d);
            });
            for_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t, const Pair& p, const float4& pj) {
                float dv = vi.x * p.dx + vi.y * p.dy + vi.z * p.dz;  // boundary velocity ignored (:336-338)
                d = fmaf(dv * p.g, pj.w * rho0, d);
            });
            d = fmaxf(d, 0.f);
        }
        divv[i] = d;
        kappa[i] = d * alpha[i];
        e = d / rho0;
    }
    reduce_error<MULTI>(e, fi, valid, partial, sm);
}

// ------------------------------------------------------------------------------------------------
// K4b: compute_velocity_changes_for_divergence dfsph_solver.rs:358-409 (+ v* = vel + vc).
// ------------------------------------------------------------------------------------------------
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(PASS_T)
k_divergence_update(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, Lists L,
                    const float* __restrict__ kappa, float4* __restrict__ vc, float4* __restrict__ vs, float* __restrict__ bforce, float inv_dt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 pi = pos[i];
    float4 v = vel[i];
    float rho0 = C.fluids[MULTI ? fid_of(v) : 0].density0;
    float ki = kappa[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
        float c = -(ki + __ldg(&kappa[j])) * pj.w * p.g;
        ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
    });
    for_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
        float c = -ki * pj.w * rho0 * p.g;
        ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
        if (BFORCE) {  // boundary2.apply_force(c.j, delta * (-inv_dt * m_i)) :405
            float s = c * (-inv_dt * pi.w);
            atomicAdd(&bforce[3 * (size_t)j + 0], s * p.dx);
            atomicAdd(&bforce[3 * (size_t)j + 1], s * p.dy);
            atomicAdd(&bforce[3 * (size_t)j + 2], s * p.dz);
        }
    });
    float4 c4 = vc[i];
    c4.x += ax; c4.y += ay; c4.z += az;
    vc[i] = c4;
    vs[i] = make_float4(v.x + c4.x, v.y + c4.y, v.z + c4.z, 0.f);
}

// v* = vel + vc after the reorder (the divergence solve works on vel + vc carried over from the previous step, Appendix A.3.2)
__global__ void k_make_vstar(const float4* __restrict__ vel, const float4* __restrict__ vc, float4* __restrict__ vs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 v = vel[i], c = vc[i];
    vs[i] = make_float4(v.x + c.x, v.y + c.y, v.z + c.z, 0.f);
}

// a10: update_velocities dfsph_solver.rs:422-430 + zero vc :689-691 + acc = gravity (predict_advection :574-578)
__global__ void k_fold_velocities(float4* __restrict__ vel, float4* __restrict__ vc, float4* __restrict__ vs, float4* __restrict__ acc, float gx, float gy,
                                  float gz) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 v = vel[i], c = vc[i];
    v.x += c.x; v.y += c.y; v.z += c.z;
    vel[i] = v;
    vc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    vs[i] = make_float4(v.x, v.y, v.z, 0.f);
    acc[i] = make_float4(gx, gy, gz, 0.f);
}
// IISPH variant: accelerations += gravity only (vc is already zero, velocities untouched).
__global__ void k_set_gravity(const float4* __restrict__ vel, float4* __restrict__ vs, float4* __restrict__ acc, float gx, float gy, float gz) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 v = vel[i];
    vs[i] = make_float4(v.x, v.y, v.z, 0.f);
    acc[i] = make_float4(gx, gy, gz, 0.f);
}

// a18: integrate_and_clear_accelerations dfsph_solver.rs:505-518 (+ v* = vel + vc)
__global__ void k_integrate_acc(const float4* __restrict__ vel, float4* __restrict__ vc, float4* __restrict__ vs, float4* __restrict__ acc, float dt,
                                float4* __restrict__ dbg_acc) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 a = acc[i], c = vc[i], v = vel[i];
    if (dbg_acc) dbg_acc[i] = a;
    c.x += a.x * dt; c.y += a.y * dt; c.z += a.z * dt;
    vc[i] = c;
    vs[i] = make_float4(v.x + c.x, v.y + c.y, v.z + c.z, 0.f);
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// K8a: compute_predicted_densities dfsph_solver.rs:98-162.  Writes rho*_i and kappa+_i = max((rho*-rho0) alpha, 0).
// ------------------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(PASS_T)
k_predict_density(const float4* __restrict__ pos, const float4* __restrict__ vs, const float4* __restrict__ vel, const float4* __restrict__ bpos,
                  const float4* __restrict__ bvel, Lists L, const float* __restrict__ dens, const float* __restrict__ alpha, float* __restrict__ pred,
                  float* __restrict__ kappa, float* __restrict__ partial, float dt, int* __restrict__ err) {
    __shared__ float sm[32];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < C.n_fluid;
    float e = 0.f;
    uint32_t fi = 0;
    if (valid) {
        float4 pi = pos[i];
        float4 vi = vs[i];
        fi = MULTI ? fid_of(vel[i]) : 0u;
        float rho0 = C.fluids[fi].density0;
        float delta = 0.f;
        for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&vs[j]);
            float dv = (vi.x - vj.x) * p.dx + (vi.y - vj.y) * p.dy + (vi.z - vj.z) * p.dz;
            delta = fmaf(dv * p.g, pj.w, delta);
        });
        for_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&bvel[j]);
            float dv = (vi.x - vj.x) * p.dx + (vi.y - vj.y) * p.dy + (vi.z - vj.z) * p.dz;
            delta = fmaf(dv * p.g, pj.w * rho0, delta);
        });
        float pd = fmaf(delta, dt, dens[i]);
        if (pd == 0.f) atomicOr(err, 1);  // assert dfsph_solver.rs:145
        pred[i] = pd;
        kappa[i] = fmaxf((pd - rho0) * alpha[i], 0.f);
        e = pd < rho0 ? 0.f : pd / rho0 - 1.0f;
    }
    reduce_error<MULTI>(e, fi, valid, partial, sm);
}

// ------------------------------------------------------------------------------------------------
// K8b: compute_velocity_changes dfsph_solver.rs:218-277 (+ v* = vel + vc).
// ------------------------------------------------------------------------------------------------
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(PASS_T)
k_pressure_update(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, Lists L,
                  const float* __restrict__ kappa, float4* __restrict__ vc, float4* __restrict__ vs, float* __restrict__ bforce, float inv_dt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 pi = pos[i];
    float4 v = vel[i];
    float rho0 = C.fluids[MULTI ? fid_of(v) : 0].density0;
    float ki = kappa[i];  // already clamped to >= 0
    float ax = 0.f, ay = 0.f, az = 0.f;
    for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
        float kij = ki + __ldg(&kappa[j]);  // max(ki,0) + max(kj,0); contributes only if > 0 (:248-254)
        float c = kij * pj.w * inv_dt * p.g;
        ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
    });
    if (ki > 0.f) {  // :257
        for_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float c = ki * pj.w * rho0 * inv_dt * p.g;
            ax = fmaf(c, p.dx, ax); ay = fmaf(c, p.dy, ay); az = fmaf(c, p.dz, az);
            if (BFORCE) {  // apply_force(c.j, delta * (inv_dt * m_i)) :269-272
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

// a22: update_positions dfsph_solver.rs:411-420: pos += (vel + vc) * dt
__global__ void k_update_positions(float4* __restrict__ pos, const float4* __restrict__ vs, float dt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 p = pos[i], v = vs[i];
    p.x += v.x * dt; p.y += v.y * dt; p.z += v.z * dt;
    pos[i] = p;
}

// ------------------------------------------------------------------------------------------------
// Nonpressure forces (predict_advection dfsph_solver.rs:565-604).  Only contacts of the SAME fluid
// count (c.i_model == c.j_model); `which` selects the fluid a force instance belongs to.
// ------------------------------------------------------------------------------------------------
// a12: XSPHViscosity::solve xsph_viscosity.rs:30-95
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(PASS_T)
k_force_xsph(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const float4* __restrict__ bvel, Lists L,
             const float* __restrict__ dens, float4* __restrict__ acc, float* __restrict__ bforce, uint32_t which, float cf, float cb, float inv_dt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 vi = vel[i];
    if (MULTI && fid_of(vi) != which) return;
    float4 pi = pos[i];
    float rho0 = C.fluids[which].density0;
    float fx = 0.f, fy = 0.f, fz = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    if (cf != 0.f)
        for_fluid_contacts<true, false>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&vel[j]);
            if (MULTI && fid_of(vj) != which) return;
            float c = cf * p.w * pj.w / __ldg(&dens[j]);  // coeff * W * (vol_j * rho0) / rho_j
            fx = fmaf(c, vj.x - vi.x, fx); fy = fmaf(c, vj.y - vi.y, fy); fz = fmaf(c, vj.z - vi.z, fz);
        });
    if (cb != 0.f) {
        float rho_i = dens[i];
        for_boundary_contacts<true, false>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&bvel[j]);
            float c = cb * p.w * pj.w * rho0 / rho_i;
            float dx = c * (vj.x - vi.x), dy = c * (vj.y - vi.y), dz = c * (vj.z - vi.z);
            bx += dx; by += dy; bz += dz;
            if (BFORCE) {  // apply_force(c.j, delta * (-m_i * inv_dt)) :87-88
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

// a13: ArtificialViscosity::solve artificial_viscosity.rs:40-124
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(PASS_T)
k_force_artificial(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, const float4* __restrict__ bvel, Lists L,
                   const float* __restrict__ dens, float4* __restrict__ acc, float* __restrict__ bforce, uint32_t which, float cf, float cb, float alpha,
                   float beta, float cs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    float4 vi = vel[i];
    if (MULTI && fid_of(vi) != which) return;
    float4 pi = pos[i];
    float rho0 = C.fluids[which].density0;
    float rho_i = dens[i];
    float eta2 = C.h * C.h * 0.01f;
    float fx = 0.f, fy = 0.f, fz = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    if (cf != 0.f)
        for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&vel[j]);
            if (MULTI && fid_of(vj) != which) return;
            float vr = p.dx * (vi.x - vj.x) + p.dy * (vi.y - vj.y) + p.dz * (vi.z - vj.z);
            if (vr < 0.f) {
                float davg = (rho_i + __ldg(&dens[j])) * 0.5f;
                float mu = C.h * vr / (p.d2 + eta2);
                float c = cf * (cs * alpha * mu - beta * mu * mu) * (pj.w / davg) * p.g;
                fx = fmaf(c, p.dx, fx); fy = fmaf(c, p.dy, fy); fz = fmaf(c, p.dz, fz);
            }
        });
    if (cb != 0.f)
        for_boundary_contacts<false, true>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float4 vj = __ldg(&bvel[j]);
            float vr = p.dx * (vi.x - vj.x) + p.dy * (vi.y - vj.y) + p.dz * (vi.z - vj.z);
            if (vr < 0.f) {
                float mu = C.h * vr / (p.d2 + eta2);
                float c = cb * (cs * alpha * mu - beta * mu * mu) * (pj.w * rho0 / rho_i) * p.g;
                bx = fmaf(c, p.dx, bx); by = fmaf(c, p.dy, by); bz = fmaf(c, p.dz, bz);
                if (BFORCE) {  // apply_force(c.j, boundary_acc * -m_i): the RUNNING sum, as the reference (:117)
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

// a14 pass 1: Akinci2013 compute_normals akinci2013_surface_tension.rs:43-68
template <bool MULTI>
__global__ void __launch_bounds__(PASS_T)
k_akinci_normals(const float4* __restrict__ pos, const float4* __restrict__ vel, Lists L, const float* __restrict__ dens, float4* __restrict__ normals,
                 uint32_t which) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    if (MULTI && fid_of(vel[i]) != which) return;
    float4 pi = pos[i];
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for_fluid_contacts<false, true>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
        if (MULTI && fid_of(__ldg(&vel[j])) != which) return;
        float c = p.g * (pj.w / __ldg(&dens[j]));
        nx = fmaf(c, p.dx, nx); ny = fmaf(c, p.dy, ny); nz = fmaf(c, p.dz, nz);
    });
    normals[i] = make_float4(nx * C.h, ny * C.h, nz * C.h, 0.f);
}

__device__ __forceinline__ float powi3(float x) { return x * x * x; }
// akinci2013_surface_tension.rs:71-88
__device__ __forceinline__ float cohesion_kernel(float r, float coh_norm, float h6_64) {
    float hr = powi3(C.h - r) * powi3(r);
    float c = r <= C.h * 0.5f ? 2.0f * hr - h6_64 : (r <= C.h ? hr : 0.f);
    return coh_norm * c;
}
// akinci2013_surface_tension.rs:90-111
__device__ __forceinline__ float adhesion_kernel(float r, float adh_norm) {
    if (r > C.h * 0.5f && r <= C.h) {
        float x = fmaxf(-4.0f * r * r / C.h + 6.0f * r - 2.0f * C.h, 0.f);
        return adh_norm * sqrtf(sqrtf(x));  // powf(0.25)
    }
    return 0.f;
}
// a14 pass 2: Akinci2013SurfaceTension::solve akinci2013_surface_tension.rs:113-192
template <bool MULTI, bool BFORCE>
__global__ void __launch_bounds__(PASS_T)
k_akinci_force(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ bpos, Lists L, const float* __restrict__ dens,
               const float4* __restrict__ normals, float4* __restrict__ acc, float* __restrict__ bforce, uint32_t which, float gamma, float adh,
               float coh_norm, float h6_64, float adh_norm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_fluid) return;
    if (MULTI && fid_of(vel[i]) != which) return;
    float4 pi = pos[i];
    float rho0 = C.fluids[which].density0;
    float rho_i = dens[i];
    float4 ni = normals[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (gamma != 0.f)
        for_fluid_contacts<false, false>(i, pi, L, pos, [&](uint32_t j, const Pair& p, const float4& pj) {
            if (MULTI && fid_of(__ldg(&vel[j])) != which) return;
            float4 nj = __ldg(&normals[j]);
            // cohesion_vec = dir * C(dist) if |dpos|^2 > eps^2 (Unit::try_new_and_get)
            float coh = p.d2 > F32_EPS * F32_EPS ? cohesion_kernel(p.r, coh_norm, h6_64) / p.r : 0.f;
            float cm = coh * (-gamma * pj.w);
            float kij = 2.0f * rho0 / (rho_i + __ldg(&dens[j]));
            ax += (-gamma * (ni.x - nj.x) + cm * p.dx) * kij;
            ay += (-gamma * (ni.y - nj.y) + cm * p.dy) * kij;
            az += (-gamma * (ni.z - nj.z) + cm * p.dz) * kij;
        });
    if (adh != 0.f)
        for_boundary_contacts<false, false>(i, pi, L, bpos, [&](uint32_t j, const Pair& p, const float4& pj) {
            float ad = p.d2 > F32_EPS * F32_EPS ? adhesion_kernel(p.r, adh_norm) / p.r : 0.f;
            float c = ad * adh * (pj.w * rho0);
            ax -= c * p.dx; ay -= c * p.dy; az -= c * p.dz;
            if (BFORCE) {  // apply_force(c.j, adhesion_acc * m_i) :188
                atomicAdd(&bforce[3 * (size_t)j + 0], c * p.dx * pi.w);
                atomicAdd(&bforce[3 * (size_t)j + 1], c * p.dy * pi.w);
                atomicAdd(&bforce[3 * (size_t)j + 2], c * p.dz * pi.w);
            }
        });
    float4 a = acc[i];
    a.x += ax; a.y += ay; a.z += az;
    acc[i] = a;
}

// ------------------------------------------------------------------------------------------------
// Host <-> sorted-order marshalling (Fluid/Boundary host SoA: fluid.rs:12-34, boundary.rs:11-24)
// ------------------------------------------------------------------------------------------------
// staging (original order, packed xyz) -> sorted arrays.  Any pointer may be null.
__global__ void k_import(uint32_t n, const uint32_t* __restrict__ orig, const float* __restrict__ o_pos, const float* __restrict__ o_vel,
                         const float* __restrict__ o_vc, const float* __restrict__ o_mass, const uint32_t* __restrict__ o_fid, float4* __restrict__ pos,
                         float4* __restrict__ vel, float4* __restrict__ vc, uint32_t lo, uint32_t hi) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t g = orig[s];
    if (g < lo || g >= hi) return;
    if (o_pos) {
        float4 p = pos[s];
        p.x = o_pos[3 * (size_t)g]; p.y = o_pos[3 * (size_t)g + 1]; p.z = o_pos[3 * (size_t)g + 2];
        if (o_mass) p.w = o_mass[g];
        pos[s] = p;
    }
    if (o_vel) {
        float4 v = vel[s];
        v.x = o_vel[3 * (size_t)g]; v.y = o_vel[3 * (size_t)g + 1]; v.z = o_vel[3 * (size_t)g + 2];
        if (o_fid) v.w = __uint_as_float(o_fid[g]);
        vel[s] = v;
    }
    if (o_vc) vc[s] = make_float4(o_vc[3 * (size_t)g], o_vc[3 * (size_t)g + 1], o_vc[3 * (size_t)g + 2], 0.f);
}
// sorted float4 array -> staging (original order, packed xyz)
__global__ void k_export3(uint32_t n, const uint32_t* __restrict__ orig, const float4* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t g = orig[s];
    float4 v = src[s];
    dst[3 * (size_t)g] = v.x; dst[3 * (size_t)g + 1] = v.y; dst[3 * (size_t)g + 2] = v.z;
}
// rows of 3 floats indexed by sorted index -> original order
__global__ void k_export_rows3(uint32_t n, const uint32_t* __restrict__ orig, const float* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t g = orig[s];
    dst[3 * (size_t)g] = src[3 * (size_t)s]; dst[3 * (size_t)g + 1] = src[3 * (size_t)s + 1]; dst[3 * (size_t)g + 2] = src[3 * (size_t)s + 2];
}
__global__ void k_export_w(uint32_t n, const uint32_t* __restrict__ orig, const float4* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    dst[orig[s]] = src[s].w;
}
__global__ void k_export1(uint32_t n, const uint32_t* __restrict__ orig, const float* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    dst[orig[s]] = src[s];
}
__global__ void k_export1u(uint32_t n, const uint32_t* __restrict__ orig, const uint32_t* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    dst[orig[s]] = (float)src[s];
}
__global__ void k_import1(uint32_t n, const uint32_t* __restrict__ orig, const float* __restrict__ src, float* __restrict__ dst) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    dst[s] = src[orig[s]];
}
__global__ void k_iota(uint32_t n, uint32_t* __restrict__ a) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) a[s] = s;
}
__global__ void k_sum_u32(uint32_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (unsigned long long)a[i] + b[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

}  // namespace sphk
