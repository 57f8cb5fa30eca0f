// Internal C++ interface of the hot-path kernels (everything here is reached only through
// the extern "C" functions of api.cu / include/tapir_b200.h).
#pragma once
#include "../../include/tapir_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace tapir {

// Bump allocator over caller-provided workspace; with base == nullptr it only measures.
struct Arena {
  char* base;
  size_t off;
  size_t cap;
  bool ok;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), off(0), cap(c), ok(true) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    if (base && off > cap) ok = false;
    return p;
  }
};

// ---- elementwise / normalisation (backbone.cu) ------------------------------------------
int split_planes(const float* src, long long ld_src, __nv_bfloat16* dst, long long ld_dst,
                 long long plane_stride, long long rows, int cols, int cols_padded, int planes,
                 cudaStream_t s);
int stem_conv(const void* video, int video_u8, const float* w_packed, int frames, int H, int W,
              float* out, cudaStream_t s);
int instnorm_stats(const float* x, int frames, long long hw, int C, double* sums, cudaStream_t s);
int instnorm_relu_split(const float* x, const double* sums, const float* w, const float* b,
                        int frames, long long hw, int C, __nv_bfloat16* out, long long plane_stride,
                        int planes, double* zero_next, int zero_channels, cudaStream_t s);
int im2col_s2(const __nv_bfloat16* in, long long in_plane_stride, int frames, int H, int W, int C,
              int taps, __nv_bfloat16* out, long long out_plane_stride, int planes, cudaStream_t s);
int layernorm_split(const float* x, long long rows, int C, const float* w, const float* b,
                    float* y, __nv_bfloat16* planes_out, long long plane_stride, int planes,
                    cudaStream_t s);
int l2_normalize(const float* x, long long rows, int C, float* out, cudaStream_t s);
int bilinear_resize(const float* src, int frames, int H, int W, int C, float* dst, int oH, int oW,
                    cudaStream_t s);
size_t backbone_workspace_bytes(int frames, int H, int W, int extra_convs, int planes);
int backbone_stem(const tapir_backbone_weights* w, const void* video_chunk, int video_u8,
                  int pass_frames, int H, int W, int frame0, int nframes, void* ws, size_t ws_bytes,
                  cudaStream_t s);
int backbone_forward(const tapir_backbone_weights* w, const void* video, int video_u8, int frames,
                     int H, int W, float* lowres, float* hires, void* ws, size_t ws_bytes,
                     cudaStream_t s, cudaEvent_t hires_ready = nullptr);

// ---- callers either side of the path (frames_io.cu) ----------------
int ingest_frames(const uint8_t* src, int frames, int H, int W, int crop_y, int crop_x, int crop_h,
                  int crop_w, float* dst, int oH, int oW, cudaStream_t s);
int postprocess_occlusions(const float* occ, const float* expd, long long n, uint8_t* visible,
                           cudaStream_t s);
int tapvid_counts(const tapir_tapvid_args* a, cudaStream_t s);

// ---- stage A (stage_a.cu) ---------------------------------------------------------------
int sample_query_features(const float* grid, int T, int gh, int gw, int C, const float* query_tyx,
                          int N, int vT, int vH, int vW, float* out, cudaStream_t s);
int cost_volume_head(const tapir_head_weights* w, const float* cost_volume, int N, int T,
                     const float* query_tyx, float temperature, int init_h, int init_w,
                     float* points, float* occ, float* expd, int* argmax, cudaStream_t s);
size_t cost_volume_workspace_bytes(int N, int T, int gh, int gw, int C);
int cost_volume_tracks(const tapir_head_weights* w, const float* qfeat, const float* grid, int N,
                       int T, int gh, int gw, int C, const float* query_tyx, float temperature,
                       int init_h, int init_w, float* points, float* occ, float* expd, int* argmax,
                       void* ws, size_t ws_bytes, cudaStream_t s);

// ---- refinement (refine.cu) -------------------------------------------------------------
int pool_pyramid(const float* grid, int T, int h, int w, int C, float* out, cudaStream_t s);
int local_corr(const tapir_corr_args* a, cudaStream_t s);
size_t mixer_workspace_bytes(long long rows, int planes);
int mixer_forward(const tapir_mixer_weights* w, const tapir_mixer_io* io, void* ws, size_t ws_bytes,
                  cudaStream_t s);
int refine_update(const tapir_update_args* a, cudaStream_t s);

}  // namespace tapir
