/* tapir_b200 - C ABI of the B200-native TAPIR / BootsTAPIR inference hot path.
 *
 * The reference (google-deepmind/tapnet) has no FFI: its boundary is the Python class
 * tapnet/torch/tapir_model.py:TAPIR.  These entry points are what a C-ABI replacement of that
 * class's stages binds; each cites the reference code it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless marked "host"; tensors are dense, channel-last,
 *    fp32 unless stated; "planes" are bf16 [P][rows][ld] split representations (x = p0+p1(+p2)).
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host and
 *    never allocate, free or retain pointers; scratch comes from the caller (`workspace`,
 *    sized by the matching *_workspace_bytes).
 *  - return 0 on success; non-zero = error (1 bad argument, 2 unsupported, 3 CUDA error,
 *    4 workspace too small) with a message available from tapir_last_error() (thread local).
 *  - no CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef TAPIR_B200_H_
#define TAPIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAPIR_B200_ABI_VERSION 2 /* 2: TAPIR_MAX_CORR_LEVELS 3 -> 5 (tapir_corr_args grew) */
#define TAPIR_MAX_MIXER_BLOCKS 12
#define TAPIR_NUM_RESNET_BLOCKS 8
#define TAPIR_MAX_EXTRA_BLOCKS 5
#define TAPIR_MAX_CORR_LEVELS 5 /* hires + lowres + up to 3 pooled levels (pyramid_level <= 3) */

/* A dense layer / convolution prepared for the split-bf16 tensor-core GEMM:
 * w = bf16 planes [planes][N][K] (K padded with zeros to a multiple of 64; for 3x3
 * convolutions K is ordered (ky, kx, cin)), bias fp32 [N] or NULL. */
typedef struct tapir_linear {
  const void* w;
  const float* bias;
  int32_t N;
  int32_t K;
  int32_t planes;
  int32_t k_logical; /* un-padded contraction length (profiling only; 0 = K) */
} tapir_linear;

/* nets.py:247-327 BlockV2 (InstanceNorm affine eps 1e-5). */
typedef struct tapir_resnet_block {
  tapir_linear proj;  /* 1x1, only when has_proj */
  tapir_linear conv0; /* 3x3, stride `stride` */
  tapir_linear conv1; /* 3x3, stride 1 */
  const float* bn0_w;
  const float* bn0_b;
  const float* bn1_w;
  const float* bn1_b;
  int32_t cin, cout, stride, has_proj;
} tapir_resnet_block;

/* nets.py:25-62 ExtraConvBlock. */
typedef struct tapir_extra_block {
  const float* ln_w;
  const float* ln_b;
  tapir_linear conv;   /* 3x3 256 -> 1024 (+bias, tanh-GELU) */
  tapir_linear conv1;  /* 3x3 1024 -> 256 (+bias) */
} tapir_extra_block;

typedef struct tapir_backbone_weights {
  const float* stem_w; /* [7][7][3][64] fp32 (ky,kx,cin,cout) repack of initial_conv.weight */
  tapir_resnet_block blocks[TAPIR_NUM_RESNET_BLOCKS];
  tapir_extra_block extra[TAPIR_MAX_EXTRA_BLOCKS];
  int32_t num_extra; /* 0 when extra_convs=False */
  int32_t planes;    /* split-bf16 planes used for activations (1..3) */
} tapir_backbone_weights;

/* tapir_model.py:121-127 torch_cost_volume_track_mods; all fp32, torch layouts. */
typedef struct tapir_head_weights {
  const float* hid1_w; /* [16][1][3][3] */
  const float* hid1_b;
  const float* hid2_w; /* [1][16][3][3] */
  const float* hid2_b;
  const float* hid3_w; /* [32][16][3][3] */
  const float* hid3_b;
  const float* hid4_w; /* [16][32] */
  const float* hid4_b;
  const float* occ_w;  /* [2][16] */
  const float* occ_b;
} tapir_head_weights;

/* nets.py:107-186 PIPsConvBlock. dw*_w are the torch Conv1d weights [2048][1][3]. */
typedef struct tapir_mixer_block {
  const float* ln_w;
  const float* dw1_w;
  const float* dw1_b;
  const float* dw2_w;
  const float* dw2_b;
  const float* ln1_w;
  tapir_linear up;   /* 512 -> 2048 (+bias, tanh-GELU) */
  tapir_linear down; /* 2048 -> 512 (+bias) */
} tapir_mixer_block;

/* nets.py:189-244 PIPSMLPMixer. */
typedef struct tapir_mixer_weights {
  tapir_linear linear;   /* in_dim (padded to 64) -> 512 */
  tapir_linear linear_1; /* 512 -> 388 */
  const float* ln_w;
  tapir_mixer_block blocks[TAPIR_MAX_MIXER_BLOCKS];
  int32_t num_blocks;
  int32_t planes;
} tapir_mixer_weights;

typedef struct tapir_mixer_io {
  const void* x_planes;    /* bf16 [planes][rows][ldx]: mixer input rows (n-major, t-minor) */
  int64_t x_plane_stride;  /* elements */
  int32_t ldx;
  int32_t num_points;      /* n */
  int32_t num_frames;      /* T ; rows = n*T */
  int32_t causal;          /* use_casual_conv */
  /* causal context (nets.py:149-176), host arrays of num_blocks device pointers, each fp32
   * [n][2][512] (ctx1) / [n][2][2048] (ctx2); all NULL = zero context / not requested */
  const float* const* ctx1_in;
  const float* const* ctx2_in;
  float* const* ctx1_out;
  float* const* ctx2_out;
  float* out;              /* fp32 [rows][ldo] ; first 388 columns valid */
  int32_t ldo;
  int32_t reserved;
} tapir_mixer_io;

typedef struct tapir_corr_level {
  const float* grid; /* [T][h][w][C] fp32, L2-normalised features */
  int32_t h, w, C, reserved;
} tapir_corr_level;

/* tapir_model.py:599-658: local correlation + assembly of the mixer input row. */
typedef struct tapir_corr_args {
  tapir_corr_level levels[TAPIR_MAX_CORR_LEVELS];
  int32_t num_levels;    /* 2 + pyramid_level */
  int32_t num_points;    /* n */
  int32_t num_frames;    /* T */
  int32_t init_h, init_w;/* coordinate frame of `pos` (initial_resolution) */
  int32_t planes;
  const float* pos;      /* [n][T][2] (x,y) */
  const float* occ;      /* [n][T] */
  const float* expd;     /* [n][T] */
  /* 128-ch (hires) and 256-ch (lowres) halves of the per-row feature: element (i,t,c) at
   * ptr[i*stride_n + t*stride_t + c]; stride_t = 0 broadcasts a per-query feature. */
  const float* feat_hi;
  int64_t feat_hi_stride_n, feat_hi_stride_t;
  const float* feat_lo;
  int64_t feat_lo_stride_n, feat_lo_stride_t;
  void* out_planes;      /* bf16 [planes][n*T][ld] */
  int64_t out_plane_stride;
  int32_t ld;            /* >= 388 + 49*num_levels, multiple of 64; pad columns are zeroed */
  int32_t reserved;
} tapir_corr_args;

/* tapir_model.py:674-685 + train2orig :435-441. */
typedef struct tapir_update_args {
  const float* res;      /* mixer output [n*T][ld_res] */
  int32_t ld_res;
  int32_t num_points, num_frames;
  int32_t init_h, init_w;       /* initial_resolution */
  int32_t resize_h, resize_w;   /* refinement resolution of this level */
  int32_t video_h, video_w;     /* original video size (for tracks_out) */
  int32_t reserved;
  const float* feat_hi;
  int64_t feat_hi_stride_n, feat_hi_stride_t;
  const float* feat_lo;
  int64_t feat_lo_stride_n, feat_lo_stride_t;
  float* pos;            /* [n][T][2] in/out (initial_resolution pixels) */
  const float* occ_in;   /* [n][T] */
  const float* expd_in;
  float* occ_out;        /* [n][T] (may alias occ_in) */
  float* expd_out;
  float* feat_out;       /* [n][T][384] = res[4:] + feat */
  float* tracks_out;     /* [n][T][2] pos scaled to the video size, or NULL */
} tapir_update_args;

/* ---- library ---------------------------------------------------------------------- */
const char* tapir_last_error(void);
int tapir_abi_version(void);
/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
unsigned long long tapir_launch_count(void);

/* Per-launch device timing (CUDA events on the launch stream).  tapir_profile_report
 * synchronises the device, writes a JSON object {"kernel class": {"launches", "ms", "flops",
 * "bytes"}} (algorithmic flops / bytes per SURVEY.md 8(d)) into buf and clears the records. */
void tapir_profile_enable(int32_t on);
int tapir_profile_report(char* buf, size_t capacity);

/* fp32 [rows][ld_src] -> bf16 planes [planes][rows][ld_dst]; columns cols..cols_padded-1 are
 * zero filled.  Used to prepare weights and features for the GEMMs. */
int tapir_split_planes(const float* src, int64_t ld_src, void* dst, int64_t ld_dst,
                       int64_t plane_stride, int64_t rows, int32_t cols, int32_t cols_padded,
                       int32_t planes, void* stream);

/* Generic split-bf16 GEMM  out = act(A . B^T + bias) + residual  (see csrc/gemm.cuh).
 * impl: 0 = tcgen05 (product path), 1 = SIMT cross-check. conv3x3 != 0 treats A as NHWC
 * planes [planes][frames][H][W][C] (stride 1, zero padding 1).  stats (nullable): fp64
 * [frames][N][2] accumulators that receive the per-(frame, column) sum and sum of squares of
 * the output (the InstanceNorm statistics of the produced tensor); the caller zeroes them. */
int tapir_gemm(const void* a_planes, int32_t lda, int64_t a_plane_stride, const tapir_linear* b,
               int64_t M, int32_t conv3x3, int32_t frames, int32_t H, int32_t W, int32_t C,
               const float* residual, int32_t ldr, int32_t act_gelu, float* out_f32, int32_t ldo,
               void* out_planes, int32_t ldp, int64_t out_plane_stride, int32_t out_P,
               double* stats, int32_t rows_per_frame, int32_t impl, void* stream);

/* ---- a3: tapir_model.py:293-392 get_feature_grids (one resolution, one frame chunk) ---- */
/* utils.py:26-42 bilinear resize, align_corners=False, [frames][H][W][C] -> [frames][oH][oW][C] */
int tapir_bilinear_resize(const float* src, int32_t frames, int32_t H, int32_t W, int32_t C,
                          float* dst, int32_t oH, int32_t oW, void* stream);
size_t tapir_backbone_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t extra_convs,
                                      int32_t planes);
/* video [frames][H][W][3] in [-1,1] -> lowres [frames][H/8][W/8][256], hires
 * [frames][H/4][W/4][128], both L2-normalised over channels (nets.py ResNet + ExtraConvs). */
int tapir_backbone_forward(const tapir_backbone_weights* w, const float* video, int32_t frames,
                           int32_t H, int32_t W, float* lowres, float* hires, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Same, reading raw uint8 [0,255] frames: preprocess_frames (pytorch_live_demo.py:30-41,
 * x / 255 * 2 - 1) is applied inside the stem conv's loads, the float video never exists. */
int tapir_backbone_forward_u8(const tapir_backbone_weights* w, const uint8_t* video,
                              int32_t frames, int32_t H, int32_t W, float* lowres, float* hires,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Same two calls with one more argument: `hires_ready_event` (a cudaEvent_t as void*, or NULL)
 * is recorded on `stream` as soon as the hires output is final, i.e. before ResNet groups 2-3
 * and the ExtraConvs run.  The multi-GPU path starts the all-gather of the hires
 * grid on another stream at that point, hidden behind the rest of the backbone.  video_u8 != 0
 * selects the uint8 reader. */
int tapir_backbone_forward_ex(const tapir_backbone_weights* w, const void* video, int32_t video_u8,
                              int32_t frames, int32_t H, int32_t W, float* lowres, float* hires,
                              void* workspace, size_t workspace_bytes, void* hires_ready_event,
                              void* stream);

/* Host-resident clips: the stem convolution (the only reader of the video) of frames
 * [frame0, frame0 + nframes) of a `pass_frames`-frame backbone pass, run as soon as that chunk
 * has arrived over PCIe while later chunks are still in flight.  `video_chunk` points at frame
 * `frame0`.  After every frame of the pass has been through this call (same workspace, same
 * pass_frames / H / W), tapir_backbone_forward_ex with video == NULL runs the rest. */
int tapir_backbone_stem(const tapir_backbone_weights* w, const void* video_chunk, int32_t video_u8,
                        int32_t pass_frames, int32_t H, int32_t W, int32_t frame0, int32_t nframes,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame ingest (the step before the path) ------------------------ */
/* uint8 frames [frames][H][W][3] -> crop window (get_frame's centre square crop,
 * pytorch_live_demo.py:88-95, or any window) -> preprocess_frames (:30-41) -> utils.bilinear
 * (utils.py:26-42) to [frames][oH][oW][3] float in [-1,1]; one pass. */
int tapir_ingest_frames(const uint8_t* src, int32_t frames, int32_t H, int32_t W, int32_t crop_y,
                        int32_t crop_x, int32_t crop_h, int32_t crop_w, float* dst, int32_t oH,
                        int32_t oW, void* stream);

/* ---- output post-processing (the step after the path) ---------------- */
/* pytorch_live_demo.py:57-59 / utils/model_utils.py:376-389:
 * visible[i] = (1 - sigmoid(occ[i])) * (1 - sigmoid(expd[i])) > 0.5 */
int tapir_postprocess_occlusions(const float* occ, const float* expd, int64_t n, uint8_t* visible,
                                 void* stream);

/* tapvid/evaluation_datasets.py:48-192 compute_tapvid_metrics, as exact integer counters per
 * track; the host forms the ratios from their sums.  counts [B][N][TAPIR_TAPVID_COUNTERS]:
 * [0] evaluated frames, [1] occlusion prediction correct, [2] ground-truth visible,
 * [3+i] within 2^i px and visible, [8+i] true positives, [13+i] false positives (i = 0..4). */
#define TAPIR_TAPVID_COUNTERS 18
typedef struct {
  const float* query_points;    /* [B][N][3]  (t, y, x) */
  const uint8_t* gt_occluded;   /* [B][N][T]  bool */
  const float* gt_tracks;       /* [B][N][T][2] (x, y) */
  const uint8_t* pred_occluded; /* [B][N][T] bool, or NULL to threshold the two logit arrays */
  const float* pred_occ_logits; /* [B][N][T] (used when pred_occluded is NULL) */
  const float* pred_expd_logits;
  const float* pred_tracks;     /* [B][N][T][2] */
  int32_t B, N, T;
  int32_t query_mode;           /* 0 = 'first', 1 = 'strided' */
  int32_t* counts;
} tapir_tapvid_args;
int tapir_tapvid_counts(const tapir_tapvid_args* args, void* stream);

/* ---- a4: tapir_model.py:217-291 + utils.py:45-73 (trilinear, border clamp) ---------- */
/* query_tyx [N][3] in video coordinates (vT,vH,vW) -> out [N][C] */
int tapir_sample_query_features(const float* grid, int32_t T, int32_t gh, int32_t gw, int32_t C,
                                const float* query_tyx, int32_t N, int32_t vT, int32_t vH,
                                int32_t vW, float* out, void* stream);

/* ---- a6: tapir_model.py:687-761 + utils.py:116-193 --------------------------------- */
size_t tapir_cost_volume_workspace_bytes(int32_t N, int32_t T, int32_t gh, int32_t gw, int32_t C);
/* qfeat [N][C], grid [T][gh][gw][C] (gh=gw=32 supported), query_tyx [N][3] in
 * initial_resolution pixel coordinates or NULL.  Outputs: points [N][T][2] (x,y in
 * initial_resolution pixels), occ/expd logits [N][T], argmax [N][T] flat cell index
 * (nullable). */
int tapir_cost_volume_tracks(const tapir_head_weights* w, const float* qfeat, const float* grid,
                             int32_t N, int32_t T, int32_t gh, int32_t gw, int32_t C,
                             const float* query_tyx, float softmax_temperature, int32_t init_h,
                             int32_t init_w, float* points, float* occ, float* expd,
                             int32_t* argmax, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- a9: tapir_model.py:519-527 avg_pool3d (2,2,1) ---------------------------------- */
int tapir_pool_pyramid(const float* grid, int32_t T, int32_t h, int32_t w, int32_t C, float* out,
                       void* stream);

/* ---- a7 / a8 ------------------------------------------------------------------------ */
int tapir_local_corr(const tapir_corr_args* args, void* stream);
size_t tapir_mixer_workspace_bytes(int64_t rows, int32_t planes);
int tapir_mixer_forward(const tapir_mixer_weights* w, const tapir_mixer_io* io, void* workspace,
                        size_t workspace_bytes, void* stream);
int tapir_refine_update(const tapir_update_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAPIR_B200_H_ */
