// Instantiation unit for the streaming separable-energy kernels (stream_separable.cuh).  Compiled several
// times by mjhmc_b200/build.py with -DMJ_T=<type> -DMJ_TAG=<tag> -DMJ_DA=<dims per thread>.
#include "stream_separable.cuh"

#define MJ_CAT2(a, b) a##b
#define MJ_CAT(a, b) MJ_CAT2(a, b)

namespace mjhmc {

stream_launch_fn MJ_CAT(find_stream_, MJ_TAG)(int dist_kind, int DT) {
    if (DT == MJ_DA) return pick_stream_dist<MJ_T, MJ_DA>(dist_kind);
    return nullptr;
}

}  // namespace mjhmc
