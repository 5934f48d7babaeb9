// IVF-Flat search path (placeholder until the list-scan kernels land).
#include "index.cuh"
namespace b200 {
int ivf_create(b200_index*, int, const float*) { set_error("IVF-Flat not built yet"); return B200_ERR_UNSUPPORTED; }
int ivf_finalize(b200_index*) { return B200_OK; }
int ivf_add_synthetic(b200_index*, int64_t, int64_t, const b200_synth_spec*) { set_error("IVF-Flat not built yet"); return B200_ERR_UNSUPPORTED; }
int ivf_search_keys(b200_index*, const float*, int, int, unsigned long long*, cudaStream_t) { set_error("IVF-Flat not built yet"); return B200_ERR_UNSUPPORTED; }
int ivf_lists(b200_index*, int64_t*, int64_t*) { set_error("IVF-Flat not built yet"); return B200_ERR_UNSUPPORTED; }
void ivf_free(b200_index*) {}
}
