// Instantiation unit for the register-resident fused kernels.  Compiled several
// times by mjhmc_b200/build.py with -DMJ_T=<type> -DMJ_TAG=<tag> -DMJ_DA=<dim> -DMJ_DB=<dim>
// so the template instantiations build in parallel.
#include "fused_elementwise.cuh"

#define MJ_CAT2(a, b) a##b
#define MJ_CAT(a, b) MJ_CAT2(a, b)

namespace mjhmc {

fused_launch_fn MJ_CAT(find_fused_, MJ_TAG)(int dist_kind, int D) {
    if (D == MJ_DA) return pick_dist<MJ_T, MJ_DA>(dist_kind);
    if (D == MJ_DB) return pick_dist<MJ_T, MJ_DB>(dist_kind);
    return nullptr;
}

}  // namespace mjhmc
