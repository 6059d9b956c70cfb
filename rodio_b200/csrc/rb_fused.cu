// rb_fused.cu — fused resample -> channel-map -> effects -> mix kernels (fast path).
#include "rb_fused.h"

struct rb_fused_plan {
    int unused;
};

cudaError_t rb_fused_try_create(const rb_fused_stream*, size_t, uint16_t, float*, uint64_t, uint32_t, int, cudaStream_t,
                                rb_fused_plan** out) {
    *out = nullptr;   // first milestone: every batch goes through the general path
    return cudaSuccess;
}
cudaError_t rb_fused_run(rb_fused_plan*, cudaStream_t) { return cudaSuccess; }
void rb_fused_destroy(rb_fused_plan* p) { delete p; }
uint32_t rb_fused_launch_count(const rb_fused_plan*) { return 0; }
