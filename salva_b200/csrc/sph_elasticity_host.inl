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

// init :84-113 — runs when the particle count differs from the captured rest pose
sph_status elasticity_init(sph_world* w, uint32_t fluid, ForceRec& fr) {
    FluidRec& f = w->fluids[fluid];
    size_t n = f.n;
    if (!fr.elastic) fr.elastic = new ElasticityState();
    ElasticityState& E = *fr.elastic;
    float young = fr.d.p[0], nu = fr.d.p[1];  // elasticity_coefficients :15-25
    E.d0 = (young * (1.f - nu)) / ((1.f + nu) * (1.f - 2.f * nu));
    E.d1 = (young * nu) / ((1.f + nu) * (1.f - 2.f * nu));
    E.d2 = (young * (1.f - 2.f * nu)) / (2.f * (1.f + nu) * (1.f - 2.f * nu));
    // Vec::resize semantics: leading values of volumes0 / rotations survive a re-initialisation (:90-95)
    float4* old_pos0 = E.pos0;
    float* old_rot = E.rot;
    size_t old_n = E.n;
    float* old_vol = nullptr;
    if (old_pos0 && old_n) {
        CU(cudaMalloc(&old_vol, old_n * sizeof(float)));
        LAUNCH(k_export_w_plain, old_n, 256, (uint32_t)old_n, old_pos0, old_vol);
    }
    for (void* p : {(void*)E.nbr0, (void*)E.cnt0, (void*)E.grad_tr, (void*)E.stress, (void*)E.cur, (void*)E.slot_of})
        if (p) cudaFree(p);
    E.nbr0 = nullptr;
    uint32_t stride0 = (uint32_t)((n + 31) / 32 * 32);
    uint32_t cap0 = std::max<uint32_t>(w->cap_f, 16);
    CU(cudaMalloc(&E.pos0, (n + 1) * sizeof(float4)));
    CU(cudaMalloc(&E.cnt0, (n + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&E.rot, (9 * n + 9) * sizeof(float)));
    CU(cudaMalloc(&E.grad_tr, (9 * n + 9) * sizeof(float)));
    CU(cudaMalloc(&E.stress, (6 * n + 6) * sizeof(float)));
    CU(cudaMalloc(&E.cur, (n + 1) * sizeof(float4)));
    CU(cudaMalloc(&E.slot_of, (n + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&E.nbr0, (size_t)cap0 * stride0 * sizeof(uint32_t)));
    E.cap0 = cap0;
    E.stride0 = stride0;
    size_t keep = std::min(old_n, n);
    if (old_rot && keep) CU(cudaMemcpyAsync(E.rot, old_rot, 9 * keep * sizeof(float), cudaMemcpyDeviceToDevice, w->st));
    LAUNCH(k_el_identity, n - keep, 256, (uint32_t)n, (uint32_t)keep, E.rot);
    CU(cudaMemsetAsync(E.stress, 0, 6 * n * sizeof(float), w->st));
    int c = w->cur;
    Lists L{reinterpret_cast<const uint4*>(w->nbr_f.p), w->nbr_b.p, w->cnt_f.p, w->cnt_b.p, w->g_f.p};
    LAUNCH(k_el_to_orig, w->N, 256, w->pos[c].p, w->orig[c].p, (uint32_t)f.offset, (uint32_t)(f.offset + n), E.cur, E.slot_of);
    CU(cudaMemsetAsync(w->d_scal.p + 11, 0, sizeof(int), w->st));
    LAUNCH(k_el_capture_lists, n, 128, L, w->vel[c].p, w->orig[c].p, E.slot_of, (uint32_t)f.offset, (uint32_t)n, fluid, cap0, stride0, E.nbr0, E.cnt0,
           reinterpret_cast<uint32_t*>(w->d_scal.p + 11));
    int widest = 0;
    CU(cudaMemcpyAsync(&widest, w->d_scal.p + 11, sizeof(int), cudaMemcpyDeviceToHost, w->st));
    CU(cudaStreamSynchronize(w->st));
    if ((uint32_t)widest > cap0) return w->fail(SPH_ERR_INVALID, "elasticity rest list wider than the contact capacity");
    LAUNCH(k_el_rest_volumes, n, 128, (uint32_t)n, E.cur, E.nbr0, E.cnt0, stride0, old_vol, (uint32_t)old_n, E.pos0);
    CU(cudaStreamSynchronize(w->st));
    if (old_vol) cudaFree(old_vol);
    if (old_pos0) cudaFree(old_pos0);
    if (old_rot) cudaFree(old_rot);
    E.n = n;
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
