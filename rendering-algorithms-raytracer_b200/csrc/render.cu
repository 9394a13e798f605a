// render.cu — Scene::raytraceImage on the GPU (wavefront path tracer).  (stub: filled in next)
#include "context.cuh"
namespace miro { void render_state_free(miro_gpu_ctx*) {} }
extern "C" int miro_gpu_render(miro_gpu_ctx* ctx, const miro_gpu_camera*, const miro_gpu_render_params*, float*) {
    return miro::set_error(ctx, MIRO_GPU_EUNSUPPORTED, "miro_gpu_render: not built yet");
}
