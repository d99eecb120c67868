// sph_viscosity_host.inl — DFSPHViscosity::solve (viscosity/dfsph_viscosity.rs:292-324) on the device; included by
// sph_engine.cu.  force.p = {viscosity_coefficient, min_viscosity_iter, max_viscosity_iter, max_viscosity_error}.
namespace {

void viscosity_release(sph_world* w) {
    ViscosityState& V = w->visc;
    if (V.beta) cudaFree(V.beta);
    if (V.target) cudaFree(V.target);
    if (V.vv) cudaFree(V.vv);
    if (V.u4) cudaFree(V.u4);
    if (V.u2) cudaFree(V.u2);
    V = ViscosityState();
}

sph_status viscosity_ensure(sph_world* w) {
    ViscosityState& V = w->visc;
    size_t need = std::max<size_t>(w->stride, 32);
    if (V.cap >= need) return SPH_OK;
    viscosity_release(w);
    CU(cudaMalloc(&V.beta, 36 * need * sizeof(float)));
    CU(cudaMalloc(&V.target, 6 * need * sizeof(float)));
    CU(cudaMalloc(&V.vv, need * sizeof(float4)));
    CU(cudaMalloc(&V.u4, need * sizeof(float4)));
    CU(cudaMalloc(&V.u2, need * sizeof(float2)));
    V.cap = need;
    return SPH_OK;
}

sph_status viscosity_solve(sph_world* w, uint32_t f, ForceRec& fr) {
    if (w->tile) return w->fail(SPH_ERR_INVALID, "DFSPHViscosity is not implemented by gather_backend 1");
    if (w->slab.active) return w->fail(SPH_ERR_INVALID, "DFSPHViscosity is not supported in slab (multi-GPU) worlds yet");
    const size_t N = w->N;
    if (w->fluids[f].n == 0) return SPH_OK;
    const int c = w->cur;
    const bool multi = w->fluids.size() > 1;
    TRY(viscosity_ensure(w));
    ViscosityState& V = w->visc;
    Lists L{reinterpret_cast<const uint4*>(w->nbr_f.p), w->nbr_b.p, w->cnt_f.p, w->cnt_b.p, w->g_f.p};
    const float visc = fr.d.p[0], max_err = fr.d.p[3];
    const uint32_t min_iter = (uint32_t)fr.d.p[1], max_iter = (uint32_t)fr.d.p[2];
    DISPATCH1(k_visc_betas, multi, N, PASS_T, w->pos[c].p, w->vel[c].p, L, w->dens.p, V.beta, f);                              // :303
    // strain-rate targets (:305); forces see the PREVIOUS step's dt (dfsph_solver.rs:702), like every other plugin
    LAUNCH(k_visc_vv, w->Ntot, 256, w->vel[c].p, w->acc.p, w->dt, V.vv);
    if (multi) LAUNCH((k_visc_rates<true, false>), N, PASS_T, w->pos[c].p, w->vel[c].p, L, w->dens.p, V.vv, V.target, V.beta, V.u4, V.u2, w->partial.p, f, visc);
    else LAUNCH((k_visc_rates<false, false>), N, PASS_T, w->pos[c].p, w->vel[c].p, L, w->dens.p, V.vv, V.target, V.beta, V.u4, V.u2, w->partial.p, f, visc);
    fr.visc_iters = 0;
    for (uint32_t i = 0; i < max_iter; ++i) {  // :307-323
        if (i) LAUNCH(k_visc_vv, w->Ntot, 256, w->vel[c].p, w->acc.p, w->dt, V.vv);
        if (multi) LAUNCH((k_visc_rates<true, true>), N, PASS_T, w->pos[c].p, w->vel[c].p, L, w->dens.p, V.vv, V.target, V.beta, V.u4, V.u2, w->partial.p, f, visc);
        else LAUNCH((k_visc_rates<false, true>), N, PASS_T, w->pos[c].p, w->vel[c].p, L, w->dens.p, V.vv, V.target, V.beta, V.u4, V.u2, w->partial.p, f, visc);
        w->errsum_ready = false;
        if (!(i < min_iter && i + 1 < max_iter)) {  // an evaluation that cannot break the loop needs no read-back
            float avg = 0.f;
            TRY(read_error(w, cdiv(N, PASS_T), &avg));
            fr.visc_err = avg;
            if (avg <= max_err && i >= min_iter) break;
        }
        DISPATCH1(k_visc_accel, multi, N, PASS_T, w->pos[c].p, w->vel[c].p, L, V.u4, V.u2, w->acc.p, f, w->inv_dt);
        fr.visc_iters++;
    }
    return SPH_OK;
}

}  // namespace
