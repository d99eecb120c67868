// IISPHSolver::step host driver (iisph_solver.rs:643-711) — filled in by a later milestone.
namespace {
sph_status iisph_step(sph_world* w, float, const float*) { return w->fail(SPH_ERR_INVALID, "IISPH solver is not built yet"); }
void iisph_release(sph_world*) {}
const float* iisph_pred(sph_world* w) { return w->pred.p; }
}  // namespace
