// tcgen05 / TMEM / TMA back end of the row-shifted GEMM (precision 1 and 2).
#include "gemm.cuh"

namespace fsb {

size_t tc_packed_weight_bytes(const ConvGeom& c) { return simt_packed_weight_bytes(c); }
int tc_pack_weights(const float*, const float*, const ConvGeom&, void*, cudaStream_t) {
    set_error("tcgen05 back end not built yet");
    return FSB_E_INVALID;
}
int tc_fwd(int, const void*, const void*, float*, const ConvGeom&, cudaStream_t) {
    set_error("tcgen05 back end not built yet");
    return FSB_E_INVALID;
}
int tc_dgrad(int, const void*, const void*, float*, const ConvGeom&, cudaStream_t) {
    set_error("tcgen05 back end not built yet");
    return FSB_E_INVALID;
}
size_t tc_wgrad_scratch_bytes(const ConvGeom& c) { return simt_wgrad_scratch_bytes(c); }
int tc_wgrad(int, const void*, const void*, float*, void*, const ConvGeom&, cudaStream_t) {
    set_error("tcgen05 back end not built yet");
    return FSB_E_INVALID;
}

}  // namespace fsb
