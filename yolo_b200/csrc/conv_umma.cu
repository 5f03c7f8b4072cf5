// placeholder until the tcgen05 kernel lands
#include "conv_umma.cuh"
namespace yb {
int umma_prepare_weights(UmmaConv& u, int, const float*, int, int, int, int, int, int, int, bool, int, cudaStream_t) { u.eligible = false; return YOLO_OK; }
int umma_build_maps(UmmaConv&, void*, int, int, int, int, int, int) { return YOLO_OK; }
int launch_conv_umma(const UmmaConv&, const ConvDesc&, cudaStream_t) { return fail(YOLO_E_UNSUPPORTED, "umma path not built"); }
void umma_release(UmmaConv&) {}
}
