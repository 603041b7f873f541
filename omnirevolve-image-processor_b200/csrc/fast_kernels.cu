// fast_kernels.cu -- placeholder: the fast path is not implemented yet, everything routes to the
// generic kernels.
#include "fast_kernels.cuh"

void fast_ctx_release(omni_ctx *) {}
bool fast_resize_2x_ok(const u8 *, int, size_t, const u8 *, int, size_t) { return false; }
cudaError_t fast_resize_2x(const u8 *, size_t, u8 *, int, int, size_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t fast_assign(omni_ctx *, const u8 *px, int h, int w, size_t pitch, const AssignParams &P, int mode_lab,
                        u8 *labels, size_t lpitch, cudaStream_t st)
{
    return g_assign(px, h, w, pitch, P, mode_lab, labels, lpitch, st);
}
bool fast_masks_supported(int, int) { return false; }
int fast_layer_masks(omni_ctx *, const u8 *, int, int, size_t, int, int, int, u8 *, size_t, size_t, cudaStream_t) { return OMNI_ERR_UNSUPPORTED; }
bool fast_edges_supported(const omni_edge_params *) { return false; }
int fast_edges(omni_ctx *, const u8 *, int, int, int, size_t, size_t, const omni_edge_params *, const BlurParams &, int, int,
               u8 *, size_t, size_t, cudaStream_t) { return OMNI_ERR_UNSUPPORTED; }
int fast_color_edge(omni_ctx *, const u8 *, int, int, size_t, const AssignParams &, const omni_edge_params *, const BlurParams &,
                    int, int, u8 *, size_t, u8 *, size_t, size_t, u8 *, size_t, size_t, cudaStream_t) { return OMNI_ERR_UNSUPPORTED; }
