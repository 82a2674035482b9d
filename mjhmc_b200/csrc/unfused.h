// Host-side declarations shared by api.cu and unfused.cu.
#pragma once
#include "common.cuh"

namespace mjhmc {

struct DistParams {
    int kind, d, nbasis;
    double p[4];
    double coef[12];
    const void *a0, *a1, *a2;
};

struct FullPtrs { void *X, *V, *G, *EX, *EV; };

cudaError_t launch_energy(int dtype, const DistParams& dp, const void* X, long long n, long long ld, void* E, cudaStream_t s);
cudaError_t launch_gradient(int dtype, const DistParams& dp, const void* X, long long n, long long ld, void* G, cudaStream_t s);
cudaError_t launch_kinetic(int dtype, int d, const void* V, long long n, long long ld, void* EV, cudaStream_t s);
cudaError_t launch_kick(int dtype, int d, void* X, void* V, const void* G, long long n, long long ld, double eps,
                        bool drift, cudaStream_t s);
cudaError_t launch_transition(int dtype, const LaunchParams& p, const FullPtrs& cur, const FullPtrs& prop,
                              const void* H_flf, cudaStream_t s);

}  // namespace mjhmc
