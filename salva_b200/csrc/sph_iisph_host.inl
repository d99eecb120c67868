// IISPHSolver::step host driver (iisph_solver.rs:643-711).  Included by sph_engine.cu.
namespace {

sph_status iisph_ensure(sph_world* w) {
    IisphState& s = w->iisph;
    size_t N = w->N;
    if (N + 8 <= s.cap) return SPH_OK;
    iisph_release(w);
    size_t cap = N + N / 4 + 8;
    CU(cudaMalloc(&s.dii, cap * sizeof(float4)));
    CU(cudaMalloc(&s.dij_pjl, cap * sizeof(float4)));
    CU(cudaMalloc(&s.s, cap * sizeof(float4)));
    CU(cudaMalloc(&s.aii, cap * sizeof(float)));
    CU(cudaMalloc(&s.next_p, cap * sizeof(float)));
    CU(cudaMalloc(&s.prho, cap * sizeof(float)));
    CU(cudaMalloc(&s.next_prho, cap * sizeof(float)));
    s.cap = cap;
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = s.s;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = cap * sizeof(float4);
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.readMode = cudaReadModeElementType;
    CU(cudaCreateTextureObject(&s.tex_s, &rd, &td, nullptr));
    return SPH_OK;
}

void iisph_release(sph_world* w) {
    IisphState& s = w->iisph;
    if (s.tex_s) cudaDestroyTextureObject(s.tex_s);
    for (void* p : {(void*)s.dii, (void*)s.dij_pjl, (void*)s.s, (void*)s.aii, (void*)s.next_p, (void*)s.prho, (void*)s.next_prho})
        if (p) cudaFree(p);
    s = IisphState();
}

const float* iisph_pred(sph_world* w) { return w->pred.p; }

sph_status iisph_step(sph_world* w, float dt_total, const float g[3]) {
    size_t N = w->N;
    int c = w->cur, bc = w->bcur;
    const bool multi = w->fluids.size() > 1, bf = any_bforce(w);
    TRY(iisph_ensure(w));
    IisphState& S = w->iisph;
    Lists L{reinterpret_cast<const uint4*>(w->nbr_f.p), w->nbr_b.p, w->cnt_f.p, w->cnt_b.p, w->g_f.p};
    // predict_advection :653-660 (forces see the PREVIOUS step's dt / inv_dt), then timestep.advance :661
    LAUNCH(k_set_gravity, N, 256, w->vel[c].p, w->vs.p, w->acc.p, g[0], g[1], g[2]);
    TRY(phase_forces(w));
    CU(cudaEventRecord(w->ev[EV_FORCES], w->st));
    timestep_advance(w, dt_total);
    LAUNCH(k_integrate_acc, N, 256, w->vel[c].p, w->vc[c].p, w->vs.p, w->acc.p, w->dt, (float4*)nullptr, (float2*)nullptr, (const float4*)nullptr, (Rec8*)nullptr, (const float*)nullptr);  // :662
    CU(cudaEventRecord(w->ev[EV_INTEG], w->st));
    DISPATCH1(k_iisph_dii, multi, N, PASS_T, w->pos[c].p, w->vel[c].p, w->bpos[bc].p, L, w->dens.p, S.dii, w->dt);            // :665-671
    LAUNCH(k_iisph_warm_start, N, 256, w->press[c].p, w->dens.p, S.prho);                                                        // :673-677
    uint32_t nblk = 0;
    TRY(launch_vel_divergence(w, true, &nblk));                                                                                  // :679-685
    w->errsum_ready = false;  // IISPH ignores this evaluation's error (iisph_solver.rs:679: `let _ =`)
    DISPATCH1(k_iisph_aii, multi, N, PASS_T, w->pos[c].p, w->vel[c].p, w->bpos[bc].p, L, w->dens.p, S.dii, S.aii, w->dt);       // :687-693
    // pressure_solve :422-456
    float *p_cur = w->press[c].p, *p_next = S.next_p, *pr_cur = S.prho, *pr_next = S.next_prho;
    w->stats.n_pressure_iter = w->stats.n_pressure_eval = 0;
    uint32_t maxit = w->force_press >= 0 ? (uint32_t)w->force_press : w->desc.max_pressure_iter;
    for (uint32_t i = 0; i < maxit; ++i) {
        TRY(span_begin(w, SP_PRED));
        DISPATCH1(k_iisph_dij_pjl, multi, N, PASS_T, w->pos[c].p, L, pr_cur, p_cur, S.dii, S.dij_pjl, S.s, w->dt);
        TRY(span_end(w));
        TRY(span_begin(w, SP_PUPD));
        if (multi)
            LAUNCH((k_iisph_next_pressures<true, true>), N, PASS_T, w->pos[c].p, w->vel[c].p, w->bpos[bc].p, L, w->dens.p, w->pred.p, S.aii, p_cur,
                   S.dij_pjl, S.s, S.tex_s, p_next, pr_next, w->partial.p, w->dt, w->desc.omega);
        else
            LAUNCH((k_iisph_next_pressures<false, true>), N, PASS_T, w->pos[c].p, w->vel[c].p, w->bpos[bc].p, L, w->dens.p, w->pred.p, S.aii, p_cur,
                   S.dij_pjl, S.s, S.tex_s, p_next, pr_next, w->partial.p, w->dt, w->desc.omega);
        TRY(span_end(w));
        std::swap(p_cur, p_next);  // :444
        std::swap(pr_cur, pr_next);
        w->stats.n_pressure_iter++;
        w->stats.n_pressure_eval++;
        if (w->force_press < 0 && i < w->desc.min_pressure_iter && i + 1 < maxit) {
            w->errsum_ready = false;  // `i >= min_pressure_iter` is required to break (iisph_solver.rs:446-454): no read-back needed
        } else if (w->force_press < 0) {
            float avg;
            TRY(read_error(w, cdiv(N, PASS_T), &avg));
            w->stats.last_density_error = avg;
            if (avg <= w->desc.max_density_error && i >= w->desc.min_pressure_iter) break;
        }
    }
    if (p_cur != w->press[c].p) CU(cudaMemcpyAsync(w->press[c].p, p_cur, N * sizeof(float), cudaMemcpyDeviceToDevice, w->st));
    DISPATCH2(k_iisph_velocity_changes, multi, bf, N, PASS_T, w->pos[c].p, w->vel[c].p, w->bpos[bc].p, L, pr_cur, w->vc[c].p, w->bforce.p, w->dt);  // :697-703
    CU(cudaEventRecord(w->ev[EV_PRESS], w->st));
    LAUNCH(k_iisph_update, N, 256, w->pos[c].p, w->vel[c].p, w->vc[c].p, w->dt);  // :705-709
    CU(cudaGetLastError());
    return SPH_OK;
}

}  // namespace
