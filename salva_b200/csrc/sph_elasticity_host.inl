// Becker2009Elasticity::solve host driver (becker2009_elasticity.rs:268-334) — filled in by a later milestone.
namespace {
sph_status elasticity_solve(sph_world* w, uint32_t, ForceRec&) { return w->fail(SPH_ERR_INVALID, "Becker2009 elasticity is not built yet"); }
void elasticity_release(ForceRec&) {}
}  // namespace
