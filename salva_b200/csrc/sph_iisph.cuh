// sph_iisph.cuh — IISPH pressure solver kernels (iisph_solver.rs).
#pragma once
#include "sph_kernels.cuh"

struct IisphState {
    float4* dii = nullptr;
    float4* dij_pjl = nullptr;
    float* aii = nullptr;
    float* next_p = nullptr;
    float* pred = nullptr;
    size_t cap = 0;
};
