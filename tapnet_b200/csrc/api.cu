// extern "C" surface of libtapir_b200.so (include/tapir_b200.h).
#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"

namespace tapir {
const char* last_error();
void profile_enable(int on);
int profile_report(char* buf, size_t cap);
}

using namespace tapir;

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// NVTX range per stage: shows up in nsys / ncu --nvtx timelines, costs one
// predictable branch when no tool is attached (header-only NVTX v3, injection library loaded lazily).
struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};
}  // namespace

extern "C" {

const char* tapir_last_error(void) { return tapir::last_error(); }
int tapir_abi_version(void) { return TAPIR_B200_ABI_VERSION; }
unsigned long long tapir_launch_count(void) { return tapir::g_launch_count; }

void tapir_profile_enable(int32_t on) { tapir::profile_enable(on); }
int tapir_profile_report(char* buf, size_t capacity) { return tapir::profile_report(buf, capacity); }

int tapir_split_planes(const float* src, int64_t ld_src, void* dst, int64_t ld_dst,
                       int64_t plane_stride, int64_t rows, int32_t cols, int32_t cols_padded,
                       int32_t planes, void* stream) {
  Range nvtx_range("tapir_split_planes");
  TAPIR_CHECK_ARG(src != nullptr && dst != nullptr, "tapir_split_planes: null pointer");
  return split_planes(src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, plane_stride, rows, cols,
                      cols_padded, planes, S(stream));
}

int tapir_gemm(const void* a_planes, int32_t lda, int64_t a_plane_stride, const tapir_linear* b,
               int64_t M, int32_t conv3x3, int32_t frames, int32_t H, int32_t W, int32_t C,
               const float* residual, int32_t ldr, int32_t act_gelu, float* out_f32, int32_t ldo,
               void* out_planes, int32_t ldp, int64_t out_plane_stride, int32_t out_P,
               double* stats, int32_t rows_per_frame, int32_t impl, void* stream) {
  Range nvtx_range("tapir_gemm");
  TAPIR_CHECK_ARG(a_planes != nullptr && b != nullptr && b->w != nullptr, "tapir_gemm: null pointer");
  TAPIR_CHECK_ARG(M > 0 && M < (1ll << 31), "tapir_gemm: M out of range");
  GemmArgs g;
  g.mode = conv3x3 ? kGemmConv3x3 : kGemmPlain;
  g.planes = b->planes;
  g.M = (int)M;
  g.N = b->N;
  g.K = b->K;
  g.a = static_cast<const __nv_bfloat16*>(a_planes);
  g.lda = lda;
  g.a_plane_stride = a_plane_stride;
  g.frames = frames; g.H = H; g.W = W; g.C = C;
  g.b = static_cast<const __nv_bfloat16*>(b->w);
  g.ldb = b->K;
  g.b_plane_stride = (long long)b->N * b->K;
  g.bias = b->bias;
  g.k_logical = b->k_logical;
  g.residual = residual; g.ldr = ldr;
  g.act = act_gelu ? 1 : 0;
  g.out_f32 = out_f32; g.ldo = ldo;
  g.out_planes = static_cast<__nv_bfloat16*>(out_planes);
  g.ldp = ldp; g.out_plane_stride = out_plane_stride; g.out_P = out_P;
  g.stats = stats; g.rows_per_frame = rows_per_frame;
  if (impl == 1) return gemm_simt(g, S(stream));
  if (impl == 0) return gemm_tc(g, S(stream));
  set_error("tapir_gemm: impl must be 0 (tcgen05) or 1 (simt)");
  return kBadArgument;
}

int tapir_bilinear_resize(const float* src, int32_t frames, int32_t H, int32_t W, int32_t C,
                          float* dst, int32_t oH, int32_t oW, void* stream) {
  Range nvtx_range("tapir_bilinear_resize");
  TAPIR_CHECK_ARG(src && dst && frames > 0 && H > 0 && W > 0 && C > 0 && oH > 0 && oW > 0,
                  "tapir_bilinear_resize: bad arguments");
  return bilinear_resize(src, frames, H, W, C, dst, oH, oW, S(stream));
}

size_t tapir_backbone_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t extra_convs,
                                      int32_t planes) {
  return backbone_workspace_bytes(frames, H, W, extra_convs, planes);
}

int tapir_backbone_forward(const tapir_backbone_weights* w, const float* video, int32_t frames,
                           int32_t H, int32_t W, float* lowres, float* hires, void* workspace,
                           size_t workspace_bytes, void* stream) {
  Range nvtx_range("tapir_backbone_forward");
  return backbone_forward(w, video, 0, frames, H, W, lowres, hires, workspace, workspace_bytes, S(stream));
}

int tapir_backbone_forward_u8(const tapir_backbone_weights* w, const uint8_t* video,
                              int32_t frames, int32_t H, int32_t W, float* lowres, float* hires,
                              void* workspace, size_t workspace_bytes, void* stream) {
  Range nvtx_range("tapir_backbone_forward_u8");
  return backbone_forward(w, video, 1, frames, H, W, lowres, hires, workspace, workspace_bytes, S(stream));
}

int tapir_backbone_forward_ex(const tapir_backbone_weights* w, const void* video, int32_t video_u8,
                              int32_t frames, int32_t H, int32_t W, float* lowres, float* hires,
                              void* workspace, size_t workspace_bytes, void* hires_ready_event,
                              void* stream) {
  Range nvtx_range("tapir_backbone_forward_ex");
  return backbone_forward(w, video, video_u8 ? 1 : 0, frames, H, W, lowres, hires, workspace,
                          workspace_bytes, S(stream), static_cast<cudaEvent_t>(hires_ready_event));
}

int tapir_backbone_stem(const tapir_backbone_weights* w, const void* video_chunk, int32_t video_u8,
                        int32_t pass_frames, int32_t H, int32_t W, int32_t frame0, int32_t nframes,
                        void* workspace, size_t workspace_bytes, void* stream) {
  Range nvtx_range("tapir_backbone_stem");
  return backbone_stem(w, video_chunk, video_u8 ? 1 : 0, pass_frames, H, W, frame0, nframes, workspace,
                       workspace_bytes, S(stream));
}

int tapir_ingest_frames(const uint8_t* src, int32_t frames, int32_t H, int32_t W, int32_t crop_y,
                        int32_t crop_x, int32_t crop_h, int32_t crop_w, float* dst, int32_t oH,
                        int32_t oW, void* stream) {
  Range nvtx_range("tapir_ingest_frames");
  return ingest_frames(src, frames, H, W, crop_y, crop_x, crop_h, crop_w, dst, oH, oW, S(stream));
}

int tapir_postprocess_occlusions(const float* occ, const float* expd, int64_t n, uint8_t* visible,
                                 void* stream) {
  Range nvtx_range("tapir_postprocess_occlusions");
  return postprocess_occlusions(occ, expd, n, visible, S(stream));
}

int tapir_tapvid_counts(const tapir_tapvid_args* args, void* stream) {
  Range nvtx_range("tapir_tapvid_counts");
  return tapvid_counts(args, S(stream));
}

int tapir_sample_query_features(const float* grid, int32_t T, int32_t gh, int32_t gw, int32_t C,
                                const float* query_tyx, int32_t N, int32_t vT, int32_t vH,
                                int32_t vW, float* out, void* stream) {
  Range nvtx_range("tapir_sample_query_features");
  return sample_query_features(grid, T, gh, gw, C, query_tyx, N, vT, vH, vW, out, S(stream));
}

size_t tapir_cost_volume_workspace_bytes(int32_t N, int32_t T, int32_t gh, int32_t gw, int32_t C) {
  return cost_volume_workspace_bytes(N, T, gh, gw, C);
}

int tapir_cost_volume_tracks(const tapir_head_weights* w, const float* qfeat, const float* grid,
                             int32_t N, int32_t T, int32_t gh, int32_t gw, int32_t C,
                             const float* query_tyx, float softmax_temperature, int32_t init_h,
                             int32_t init_w, float* points, float* occ, float* expd,
                             int32_t* argmax, void* workspace, size_t workspace_bytes,
                             void* stream) {
  Range nvtx_range("tapir_cost_volume_tracks");
  return cost_volume_tracks(w, qfeat, grid, N, T, gh, gw, C, query_tyx, softmax_temperature, init_h,
                            init_w, points, occ, expd, argmax, workspace, workspace_bytes, S(stream));
}

int tapir_pool_pyramid(const float* grid, int32_t T, int32_t h, int32_t w, int32_t C, float* out,
                       void* stream) {
  Range nvtx_range("tapir_pool_pyramid");
  return pool_pyramid(grid, T, h, w, C, out, S(stream));
}

int tapir_local_corr(const tapir_corr_args* args, void* stream) {
  Range nvtx_range("tapir_local_corr");
  return local_corr(args, S(stream));
}

size_t tapir_mixer_workspace_bytes(int64_t rows, int32_t planes) {
  return mixer_workspace_bytes(rows, planes);
}

int tapir_mixer_forward(const tapir_mixer_weights* w, const tapir_mixer_io* io, void* workspace,
                        size_t workspace_bytes, void* stream) {
  Range nvtx_range("tapir_mixer_forward");
  return mixer_forward(w, io, workspace, workspace_bytes, S(stream));
}

int tapir_refine_update(const tapir_update_args* args, void* stream) {
  Range nvtx_range("tapir_refine_update");
  return refine_update(args, S(stream));
}

}  // extern "C"
