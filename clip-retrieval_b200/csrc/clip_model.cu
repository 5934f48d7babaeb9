// Embed path handle (placeholder until the forward lands).
#include "common.cuh"
using namespace b200;
extern "C" {
int b200_clip_create(const b200_clip_config*, int, b200_clip**) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_destroy(b200_clip*) { return B200_OK; }
int b200_clip_load_weights(b200_clip*, const b200_tensor_view*, int) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_encode_image_device(b200_clip*, const float*, int, void*, int, int, void*) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_encode_text_device(b200_clip*, const int64_t*, int, void*, int, int, void*) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_encode_image(b200_clip*, const float*, int, void*, int, int) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_encode_text(b200_clip*, const int64_t*, int, void*, int, int) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_last_timing(const b200_clip*, float*, int*) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
int b200_clip_set_profiling(b200_clip*, int) { set_error("embed path not built yet"); return B200_ERR_UNSUPPORTED; }
}
