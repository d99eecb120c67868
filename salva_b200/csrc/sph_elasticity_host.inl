// Becker2009Elasticity::solve host driver (becker2009_elasticity.rs:268-334).  Included by sph_engine.cu.
namespace {

void elasticity_release(ForceRec& fr) {
    ElasticityState* e = fr.elastic;
    if (!e) return;
    for (void* p : {(void*)e->pos0, (void*)e->nbr0, (void*)e->cnt0, (void*)e->rot, (void*)e->grad_tr, (void*)e->stress, (void*)e->cur, (void*)e->slot_of})
        if (p) cudaFree(p);
    delete e;
    fr.elastic = nullptr;
}

// Allocates every array of a rest pose for n particles into `E` (which must be empty).  On failure everything allocated
// so far is released again, so the caller never sees a half-built state.
sph_status elasticity_alloc(sph_world* w, ElasticityState& E, size_t n, uint32_t cap0, uint32_t stride0) {
    bool ok = cudaMalloc(&E.pos0, (n + 1) * sizeof(float4)) == cudaSuccess && cudaMalloc(&E.cnt0, (n + 1) * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&E.rot, (9 * n + 9) * sizeof(float)) == cudaSuccess && cudaMalloc(&E.grad_tr, (9 * n + 9) * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&E.stress, (6 * n + 6) * sizeof(float)) == cudaSuccess && cudaMalloc(&E.cur, (n + 1) * sizeof(float4)) == cudaSuccess &&
              cudaMalloc(&E.slot_of, (n + 1) * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&E.nbr0, std::max<size_t>((size_t)cap0 * stride0, 1) * sizeof(uint32_t)) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        for (void* p : {(void*)E.pos0, (void*)E.nbr0, (void*)E.cnt0, (void*)E.rot, (void*)E.grad_tr, (void*)E.stress, (void*)E.cur, (void*)E.slot_of})
            if (p) cudaFree(p);
        E = ElasticityState();
        return w->fail(SPH_ERR_OOM, "elasticity rest pose: device allocation failed");
    }
    E.cap0 = cap0;
    E.stride0 = stride0;
    E.n = n;
    return SPH_OK;
}

void elasticity_coefficients(ElasticityState& E, const ForceRec& fr) {
    float young = fr.d.p[0], nu = fr.d.p[1];  // elasticity_coefficients :15-25
    E.d0 = (young * (1.f - nu)) / ((1.f + nu) * (1.f - 2.f * nu));
    E.d1 = (young * nu) / ((1.f + nu) * (1.f - 2.f * nu));
    E.d2 = (young * (1.f - 2.f * nu)) / (2.f * (1.f + nu) * (1.f - 2.f * nu));
}

// init :84-113 — runs when the particle count differs from the captured rest pose.  The new state is built in a local
// and committed to fr.elastic only when every step succeeded (the old state survives an error untouched).
sph_status elasticity_init(sph_world* w, uint32_t fluid, ForceRec& fr) {
    FluidRec& f = w->fluids[fluid];
    const size_t n = f.n;
    ElasticityState* old = fr.elastic;
    const size_t old_n = old ? old->n : 0;
    ElasticityState E;
    elasticity_coefficients(E, fr);
    const uint32_t stride0 = (uint32_t)((n + 31) / 32 * 32);
    const uint32_t cap0 = std::max<uint32_t>(w->cap_f, 16);
    TRY(elasticity_alloc(w, E, n, cap0, stride0));
    ForceRec tmp;  // owns E until it is committed: elasticity_release(tmp) frees it on every error path
    tmp.elastic = new ElasticityState(E);
    float* old_vol = nullptr;
    auto bail = [&](sph_status s) {
        if (old_vol) cudaFree(old_vol);
        elasticity_release(tmp);
        return s;
    };
#define EL_CU(call)                                                                                                        \
    do {                                                                                                                   \
        cudaError_t e_ = (call);                                                                                           \
        if (e_ != cudaSuccess) return bail(w->fail(SPH_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));        \
    } while (0)
    // Vec::resize semantics: leading values of volumes0 / rotations survive a re-initialisation (:90-95)
    if (old && old->pos0 && old_n) {
        EL_CU(cudaMalloc(&old_vol, old_n * sizeof(float)));
        LAUNCH(k_export_w_plain, old_n, 256, (uint32_t)old_n, old->pos0, old_vol);
    }
    const size_t keep = std::min(old_n, n);
    if (old && old->rot && keep) EL_CU(cudaMemcpyAsync(E.rot, old->rot, 9 * keep * sizeof(float), cudaMemcpyDeviceToDevice, w->st));
    LAUNCH(k_el_identity, n - keep, 256, (uint32_t)n, (uint32_t)keep, E.rot);
    EL_CU(cudaMemsetAsync(E.stress, 0, 6 * n * sizeof(float), w->st));
    int c = w->cur;
    Lists L{reinterpret_cast<const uint4*>(w->nbr_f.p), w->nbr_b.p, w->cnt_f.p, w->cnt_b.p, w->g_f.p};
    LAUNCH(k_el_to_orig, w->N, 256, w->pos[c].p, w->orig[c].p, (uint32_t)f.offset, (uint32_t)(f.offset + n), E.cur, E.slot_of);
    EL_CU(cudaMemsetAsync(w->d_scal.p + 11, 0, sizeof(int), w->st));
    LAUNCH(k_el_capture_lists, n, 128, L, w->vel[c].p, w->orig[c].p, E.slot_of, (uint32_t)f.offset, (uint32_t)n, fluid, cap0, stride0, E.nbr0, E.cnt0,
           reinterpret_cast<uint32_t*>(w->d_scal.p + 11));
    int widest = 0;
    EL_CU(cudaMemcpyAsync(&widest, w->d_scal.p + 11, sizeof(int), cudaMemcpyDeviceToHost, w->st));
    EL_CU(cudaStreamSynchronize(w->st));
    if ((uint32_t)widest > cap0) return bail(w->fail(SPH_ERR_INVALID, "elasticity rest list wider than the contact capacity"));
    LAUNCH(k_el_rest_volumes, n, 128, (uint32_t)n, E.cur, E.nbr0, E.cnt0, stride0, old_vol, (uint32_t)old_n, E.pos0);
    EL_CU(cudaStreamSynchronize(w->st));
#undef EL_CU
    if (old_vol) cudaFree(old_vol);
    elasticity_release(fr);      // the previous rest pose
    fr.elastic = tmp.elastic;    // commit
    return SPH_OK;
}

// Rest pose from a snapshot blob: pos0 (n float4) | cnt0 (n u32) | nbr0 (cap0 * stride0 u32) | rot (9 n f32)
sph_status elasticity_restore(sph_world* w, ForceRec& fr, size_t n, uint32_t cap0, uint32_t stride0, const char* blob) {
    ElasticityState E;
    elasticity_coefficients(E, fr);
    TRY(elasticity_alloc(w, E, n, cap0, stride0));
    const size_t nl = (size_t)cap0 * stride0;
    cudaError_t e1 = cudaMemcpyAsync(E.pos0, blob, n * sizeof(float4), cudaMemcpyHostToDevice, w->st);
    blob += n * sizeof(float4);
    cudaError_t e2 = cudaMemcpyAsync(E.cnt0, blob, n * sizeof(uint32_t), cudaMemcpyHostToDevice, w->st);
    blob += n * sizeof(uint32_t);
    cudaError_t e3 = cudaMemcpyAsync(E.nbr0, blob, nl * sizeof(uint32_t), cudaMemcpyHostToDevice, w->st);
    blob += nl * sizeof(uint32_t);
    cudaError_t e4 = cudaMemcpyAsync(E.rot, blob, 9 * n * sizeof(float), cudaMemcpyHostToDevice, w->st);
    cudaError_t e5 = cudaMemsetAsync(E.stress, 0, 6 * n * sizeof(float), w->st);
    cudaError_t e6 = cudaStreamSynchronize(w->st);
    ForceRec tmp;
    tmp.elastic = new ElasticityState(E);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess || e6 != cudaSuccess) {
        elasticity_release(tmp);
        return w->fail(SPH_ERR_CUDA, "elasticity rest pose upload failed");
    }
    elasticity_release(fr);
    fr.elastic = tmp.elastic;
    return SPH_OK;
}

sph_status elasticity_solve(sph_world* w, uint32_t fluid, ForceRec& fr) {
    if (w->tile) return w->fail(SPH_ERR_INVALID, "Becker2009 elasticity needs gather_backend 0");
    FluidRec& f = w->fluids[fluid];
    size_t n = f.n;
    if (n == 0) return SPH_OK;
    if (!fr.elastic || fr.elastic->n != n) TRY(elasticity_init(w, fluid, fr));  // :87
    ElasticityState& E = *fr.elastic;
    int c = w->cur;
    int nonlinear = fr.d.p[2] != 0.f;
    LAUNCH(k_el_to_orig, w->N, 256, w->pos[c].p, w->orig[c].p, (uint32_t)f.offset, (uint32_t)(f.offset + n), E.cur, E.slot_of);
    LAUNCH(k_el_rotations, n, 128, (uint32_t)n, E.cur, E.pos0, E.nbr0, E.cnt0, E.stride0, E.rot);
    LAUNCH(k_el_stresses, n, 128, (uint32_t)n, E.cur, E.pos0, E.nbr0, E.cnt0, E.stride0, E.rot, E.grad_tr, E.stress, E.d0, E.d1, E.d2, nonlinear);
    LAUNCH(k_el_forces, n, 128, (uint32_t)n, E.cur, E.pos0, E.nbr0, E.cnt0, E.stride0, E.rot, E.grad_tr, E.stress, E.slot_of, w->acc.p, nonlinear);
    CU(cudaGetLastError());
    return SPH_OK;
}

}  // namespace
