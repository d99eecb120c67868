// sph_elasticity.cuh — Becker2009 corotated elasticity (becker2009_elasticity.rs).
#pragma once
#include "sph_kernels.cuh"

struct ElasticityState {
    size_t n = 0;
};
