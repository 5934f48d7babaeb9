// JPEG decode on the GPU in front of the image transform: the reference decodes with PIL inside
// its DataLoader workers (clip_retrieval/clip_inference/reader.py:98-106: `Image.open(...)` then `preprocess`); here
// the compressed bytes cross PCIe and nvJPEG (CUDA toolkit) decodes them into the packed RGB uint8 HWC batch layout
// b200_preproc_run consumes — decode -> resize/crop/normalise -> patch embedding without the pixels visiting the host.
// nvJPEG is resolved at run time (dlopen libnvjpeg.so.12), so the library links and loads without it.
// Not bit-identical to PIL/libjpeg-turbo (IDCT and chroma upsampling differ by a few grey levels); the tests bound it.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <vector>

namespace b200 {

// the handful of nvjpeg.h declarations this file needs (opaque handles, plain C ABI)
typedef struct nvjpegHandle* nvjpegHandle_t;
typedef struct nvjpegJpegState* nvjpegJpegState_t;
struct nvjpegImage_t { unsigned char* channel[4]; size_t pitch[4]; };
constexpr int NVJPEG_OUTPUT_RGBI = 5;

struct NvjpegApi {
  void* lib = nullptr;
  int (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  int (*Destroy)(nvjpegHandle_t) = nullptr;
  int (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  int (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  int (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, int*, int*, int*) = nullptr;
  int (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, int, nvjpegImage_t*, cudaStream_t) = nullptr;
};

static NvjpegApi* nvjpeg_api() {
  static NvjpegApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnvjpeg.so.12", RTLD_NOW);
    if (!h) h = dlopen("libnvjpeg.so", RTLD_NOW);
    if (!h) return;
    api.lib = h;
    api.CreateSimple = (decltype(api.CreateSimple))dlsym(h, "nvjpegCreateSimple");
    api.Destroy = (decltype(api.Destroy))dlsym(h, "nvjpegDestroy");
    api.JpegStateCreate = (decltype(api.JpegStateCreate))dlsym(h, "nvjpegJpegStateCreate");
    api.JpegStateDestroy = (decltype(api.JpegStateDestroy))dlsym(h, "nvjpegJpegStateDestroy");
    api.GetImageInfo = (decltype(api.GetImageInfo))dlsym(h, "nvjpegGetImageInfo");
    api.Decode = (decltype(api.Decode))dlsym(h, "nvjpegDecode");
  });
  return (api.lib && api.CreateSimple && api.JpegStateCreate && api.GetImageInfo && api.Decode) ? &api : nullptr;
}

}  // namespace b200

struct b200_jpeg {
  int device = 0;
  b200::nvjpegHandle_t handle = nullptr;
  b200::nvjpegJpegState_t state = nullptr;
  std::mutex mu;
};

using namespace b200;

extern "C" {

int b200_jpeg_create(int device, b200_jpeg** out) {
  B200_CHECK(out != nullptr, B200_ERR_INVALID, "jpeg_create: null out");
  NvjpegApi* api = nvjpeg_api();
  B200_CHECK(api != nullptr, B200_ERR_UNSUPPORTED, "jpeg_create: libnvjpeg.so.12 could not be loaded");
  int ndev = 0;
  B200_CUDA(cudaGetDeviceCount(&ndev));
  B200_CHECK(device >= 0 && device < ndev, B200_ERR_INVALID, "jpeg_create: device %d of %d", device, ndev);
  DeviceGuard g(device);
  b200_jpeg* j = new (std::nothrow) b200_jpeg();
  B200_CHECK(j != nullptr, B200_ERR_OOM, "jpeg_create: host allocation failed");
  j->device = device;
  int e = api->CreateSimple(&j->handle);
  if (e == 0) e = api->JpegStateCreate(j->handle, &j->state);
  if (e != 0) {
    set_error("jpeg_create: nvjpeg initialisation failed with status %d", e);
    b200_jpeg_destroy(j);
    return B200_ERR_CUDA;
  }
  *out = j;
  return B200_OK;
}

int b200_jpeg_destroy(b200_jpeg* j) {
  if (!j) return B200_OK;
  NvjpegApi* api = nvjpeg_api();
  DeviceGuard g(j->device);
  cudaDeviceSynchronize();
  if (api) {
    if (j->state && api->JpegStateDestroy) api->JpegStateDestroy(j->state);
    if (j->handle && api->Destroy) api->Destroy(j->handle);
  }
  delete j;
  return B200_OK;
}

// Sizes of n JPEG streams without decoding them: h_heights[i], h_widths[i]; returns B200_ERR_INVALID naming the
// first stream nvJPEG cannot parse (the caller decodes that image on the host, as the reference does).
int b200_jpeg_info(b200_jpeg* j, const uint8_t* const* h_streams, const size_t* h_sizes, int n, int32_t* h_heights,
                   int32_t* h_widths) {
  B200_CHECK(j && h_streams && h_sizes && h_heights && h_widths && n >= 0, B200_ERR_INVALID, "jpeg_info: bad argument");
  NvjpegApi* api = nvjpeg_api();
  B200_CHECK(api != nullptr, B200_ERR_UNSUPPORTED, "jpeg_info: libnvjpeg.so.12 could not be loaded");
  std::lock_guard<std::mutex> lock(j->mu);
  for (int i = 0; i < n; i++) {
    int comps = 0, subs = 0, w[4] = {0, 0, 0, 0}, h[4] = {0, 0, 0, 0};
    const int e = api->GetImageInfo(j->handle, h_streams[i], h_sizes[i], &comps, &subs, w, h);
    B200_CHECK(e == 0 && w[0] > 0 && h[0] > 0, B200_ERR_INVALID, "jpeg_info: stream %d is not a JPEG nvJPEG can parse (status %d)", i, e);
    h_heights[i] = h[0];
    h_widths[i] = w[0];
  }
  return B200_OK;
}

// Decode n JPEG streams (host memory) into d_pixels: image i as RGB uint8 HWC at byte offset h_offsets[i] (sizes from
// b200_jpeg_info) — exactly the `pixels` / `h_offsets` / `h_heights` / `h_widths` arguments of b200_preproc_run with
// pixels_on_device = 1.  Asynchronous on `stream` after the bitstreams have been consumed.
int b200_jpeg_decode(b200_jpeg* j, const uint8_t* const* h_streams, const size_t* h_sizes, int n, uint8_t* d_pixels,
                     const int64_t* h_offsets, const int32_t* h_heights, const int32_t* h_widths, void* stream) {
  B200_CHECK(j && h_streams && h_sizes && d_pixels && h_offsets && h_heights && h_widths && n >= 0, B200_ERR_INVALID,
             "jpeg_decode: bad argument");
  NvjpegApi* api = nvjpeg_api();
  B200_CHECK(api != nullptr, B200_ERR_UNSUPPORTED, "jpeg_decode: libnvjpeg.so.12 could not be loaded");
  std::lock_guard<std::mutex> lock(j->mu);
  DeviceGuard g(j->device);
  for (int i = 0; i < n; i++) {
    nvjpegImage_t img;
    for (int c = 0; c < 4; c++) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = d_pixels + h_offsets[i];
    img.pitch[0] = (size_t)h_widths[i] * 3;
    const int e = api->Decode(j->handle, j->state, h_streams[i], h_sizes[i], NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream);
    B200_CHECK(e == 0, B200_ERR_CUDA, "jpeg_decode: nvjpegDecode failed on stream %d with status %d", i, e);
    count_launch();
  }
  return B200_OK;
}

}  // extern "C"
