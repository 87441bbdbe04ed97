// JPEG decode on the GPU for the real-data input path: the reference decodes its `<seq>/<id>.jpg`
// triples on the CPU inside tf.data (data_loader.py:241-249, tf.image.decode_jpeg under /cpu:0) and feeds uint8
// [B,H,3W,3] to the graph.  davo_decode_jpeg_batch decodes a batch of JPEG byte strings with nvJPEG straight into the
// caller's device tensor of that layout, on the caller's stream.
//
// nvJPEG is a CUDA toolkit library; like NCCL it is bound at run time (dlopen libnvjpeg.so.12), so the library has no
// link-time dependency on it and hosts without it only lose this entry point.  Decoders differ in the last bit of
// their IDCT / chroma upsampling: nvJPEG's pixels are within a few levels of libjpeg's (what TensorFlow and PIL use),
// not identical -- the host decode (PIL in davo_b200/data_loader.py) stays the default for parity.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>
#include <string>

namespace davo_jpeg {

struct Api {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
  std::string why;
};

inline const Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[3] = {std::getenv("DAVO_B200_NVJPEG_LIB"), "libnvjpeg.so.12", "libnvjpeg.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      const char* e = dlerror();
      a.why = std::string("libnvjpeg.so.12 not found (") + (e ? e : "?") + "); set DAVO_B200_NVJPEG_LIB";
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(a.lib, s);
      if (!p && a.why.empty()) a.why = std::string("nvJPEG symbol missing: ") + s;
      return p;
    };
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(sym("nvjpegCreateSimple"));
    a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("nvjpegDestroy"));
    a.StateCreate = reinterpret_cast<decltype(a.StateCreate)>(sym("nvjpegJpegStateCreate"));
    a.StateDestroy = reinterpret_cast<decltype(a.StateDestroy)>(sym("nvjpegJpegStateDestroy"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(sym("nvjpegDecode"));
  });
  return a;
}

}  // namespace davo_jpeg
