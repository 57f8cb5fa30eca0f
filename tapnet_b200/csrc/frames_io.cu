// The callers either side of the hot path, as HBM-bound one-pass kernels:
//   * frame ingest: uint8 camera / dataset frames -> centre crop -> bilinear resize -> [-1,1]
//     (pytorch_live_demo.py:30-41 preprocess_frames, :88-95 get_frame, utils.py:26-42 bilinear)
//   * output post-processing: visibility flag (pytorch_live_demo.py:57-59,
//     utils/model_utils.py:376-389) and the TAP-Vid metric counters
//     (tapvid/evaluation_datasets.py:48-192) without leaving the device.
#include "kernels.cuh"

namespace tapir {
namespace {

int grid_for(long long total, int block = 256) {
  long long g = ceil_div_ll(total, block);
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// preprocess_frames: x / 255 * 2 - 1 in fp32, every operation rounded (no FMA contraction)
__device__ __forceinline__ float normalise_u8(uint8_t v) {
  return __fsub_rn(__fmul_rn(__fdiv_rn((float)v, 255.f), 2.f), 1.f);
}

// The 256 possible inputs of normalise_u8, tabulated once per CTA (the IEEE division is ~10
// instructions; the table makes the kernels pure load/store streams).
__device__ __forceinline__ void fill_lut(float* lut) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = normalise_u8((uint8_t)i);
  __syncthreads();
}

// preprocess_frames alone (no crop, no resize): 4 bytes in, one float4 out per thread and trip,
// both fully coalesced.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ src,
                                                            float* __restrict__ dst, long long n) {
  __shared__ float lut[256];
  fill_lut(lut);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
#pragma unroll 4
  for (; i < n4; i += stride) {
    const uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = make_float4(lut[v.x], lut[v.y], lut[v.z], lut[v.w]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = lut[src[(n4 << 2) + threadIdx.x]];
}

// crop + normalise + resize.  A CTA walks over runs of 256 consecutive output pixels: one thread
// computes one pixel (3 channels) into shared memory, then the run's 768 floats leave as float4
// rows (the direct 12-byte-strided stores would touch every sector three times).  The resize is
// F.interpolate(mode='bilinear', align_corners=False) of the normalised crop, i.e. the same
// arithmetic as bilinear_resize_kernel (backbone.cu) applied after normalise_u8.
__global__ void __launch_bounds__(256) ingest_frames_kernel(const uint8_t* __restrict__ src, int H,
                                                            int W, int cy, int cx, int ch, int cw,
                                                            float* __restrict__ dst, int oH, int oW,
                                                            long long total, bool dst_aligned16) {
  __shared__ float lut[256];
  __shared__ __align__(16) float stage[256 * 3];
  fill_lut(lut);
  const float sy = (float)ch / (float)oH, sx = (float)cw / (float)oW;
  const bool same = (ch == oH && cw == oW);
  for (long long i0 = blockIdx.x * 256LL; i0 < total; i0 += (long long)gridDim.x * 256) {
    const long long i = i0 + threadIdx.x;
    if (i < total) {
      int ox, oy;
      long long f;
      if (total <= 0x7fffffffLL) {  // 32-bit divisions (the 64-bit ones cost more than the loads)
        const unsigned u = (unsigned)i, r = u / (unsigned)oW;
        ox = (int)(u - r * (unsigned)oW);
        const unsigned q = r / (unsigned)oH;
        oy = (int)(r - q * (unsigned)oH);
        f = q;
      } else {
        ox = (int)(i % oW);
        const long long r = i / oW;
        oy = (int)(r % oH);
        f = r / oH;
      }
      const uint8_t* b = src + (f * H + cy) * (long long)W * 3 + (long long)cx * 3;
      float* o = stage + threadIdx.x * 3;
      if (same) {
        const uint8_t* p = b + ((long long)oy * W + ox) * 3;
        o[0] = lut[p[0]];
        o[1] = lut[p[1]];
        o[2] = lut[p[2]];
      } else {
        float fy = sy * (oy + 0.5f) - 0.5f;
        float fx = sx * (ox + 0.5f) - 0.5f;
        if (fy < 0.f) fy = 0.f;
        if (fx < 0.f) fx = 0.f;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + ((y0 < ch - 1) ? 1 : 0), x1 = x0 + ((x0 < cw - 1) ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const uint8_t* p00 = b + ((long long)y0 * W + x0) * 3;
        const uint8_t* p01 = b + ((long long)y0 * W + x1) * 3;
        const uint8_t* p10 = b + ((long long)y1 * W + x0) * 3;
        const uint8_t* p11 = b + ((long long)y1 * W + x1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v00 = lut[p00[c]], v01 = lut[p01[c]];
          const float v10 = lut[p10[c]], v11 = lut[p11[c]];
          o[c] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
        }
      }
    }
    __syncthreads();
    const long long left = total - i0;  // pixels in this run
    if (left >= 256 && dst_aligned16) {
      if (threadIdx.x < 192)  // i0 * 12 bytes is a multiple of 16: aligned float4 rows
        reinterpret_cast<float4*>(dst + i0 * 3)[threadIdx.x] = reinterpret_cast<const float4*>(stage)[threadIdx.x];
    } else {
      for (int e = threadIdx.x; e < (int)left * 3; e += 256) dst[i0 * 3 + e] = stage[e];
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// pytorch_live_demo.py:57-59: visible = (1 - sigmoid(occ)) * (1 - sigmoid(expd)) > 0.5
__device__ __forceinline__ bool visible_from_logits(float occ, float expd) {
  return __fmul_rn(1.f - sigmoidf(occ), 1.f - sigmoidf(expd)) > 0.5f;
}

__global__ void __launch_bounds__(256) postprocess_occlusions_kernel(const float* __restrict__ occ,
                                                                     const float* __restrict__ expd,
                                                                     long long n,
                                                                     uint8_t* __restrict__ visible) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    visible[i] = visible_from_logits(occ[i], expd[i]) ? 1 : 0;
}

// tapvid/evaluation_datasets.py:48-192.  One warp per track (b, n); the 18 integer counters of a
// track are exact, the ratios are formed by the host from their sums (per video or per track).
//   [0] evaluated frames            [1] occlusion prediction == ground truth
//   [2] ground-truth visible        [3+i] within threshold 2^i and visible (i = 0..4)
//   [8+i] true positives            [13+i] false positives
__global__ void __launch_bounds__(256) tapvid_counts_kernel(const tapir_tapvid_args a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int tracks = a.B * a.N;
  if (warp >= tracks) return;
  const int T = a.T;
  // evaluation_datasets.py:125-127: np.round (half to even) then int32
  const int qf = __float2int_rn(a.query_points[(long long)warp * 3]);
  const long long base = (long long)warp * T;
  int cnt[TAPIR_TAPVID_COUNTERS];
#pragma unroll
  for (int k = 0; k < TAPIR_TAPVID_COUNTERS; ++k) cnt[k] = 0;
  for (int t = lane; t < T; t += 32) {
    // :116-123: 'first' evaluates frames after the query frame, 'strided' all but the query frame
    const bool eval = (a.query_mode == 0) ? (t > qf) : (t != qf);
    if (!eval) continue;
    const bool gt_occ = a.gt_occluded[base + t] != 0;
    bool pred_occ;
    if (a.pred_occluded != nullptr)
      pred_occ = a.pred_occluded[base + t] != 0;
    else
      pred_occ = !visible_from_logits(a.pred_occ_logits[base + t], a.pred_expd_logits[base + t]);
    const float dx = __fsub_rn(a.pred_tracks[(base + t) * 2 + 0], a.gt_tracks[(base + t) * 2 + 0]);
    const float dy = __fsub_rn(a.pred_tracks[(base + t) * 2 + 1], a.gt_tracks[(base + t) * 2 + 1]);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // :148-151
    const bool vis = !gt_occ, pvis = !pred_occ;
    cnt[0] += 1;
    cnt[1] += (pred_occ == gt_occ);
    cnt[2] += vis;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float th = (float)(1 << i);
      const bool within = d2 < th * th;
      const bool correct = within && vis;
      cnt[3 + i] += correct;
      cnt[8 + i] += correct && pvis;
      cnt[13 + i] += ((!vis) && pvis) || ((!within) && pvis);  // :177-179
    }
  }
#pragma unroll
  for (int k = 0; k < TAPIR_TAPVID_COUNTERS; ++k) {
    int v = cnt[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.counts[(long long)warp * TAPIR_TAPVID_COUNTERS + k] = v;
  }
}

}  // namespace

int ingest_frames(const uint8_t* src, int frames, int H, int W, int crop_y, int crop_x, int crop_h,
                  int crop_w, float* dst, int oH, int oW, cudaStream_t s) {
  TAPIR_CHECK_ARG(src && dst && frames > 0 && H > 0 && W > 0 && oH > 0 && oW > 0,
                  "ingest_frames: bad arguments");
  TAPIR_CHECK_ARG(crop_y >= 0 && crop_x >= 0 && crop_h > 0 && crop_w > 0 && crop_y + crop_h <= H &&
                      crop_x + crop_w <= W,
                  "ingest_frames: crop (%d,%d,%d,%d) outside the %dx%d frame", crop_y, crop_x, crop_h,
                  crop_w, H, W);
  const long long total = (long long)frames * oH * oW;
  ProfileScope ps("ingest_frames", s, 0.0,
                  (double)frames * crop_h * crop_w * 3 + (double)total * 12);
  if (crop_y == 0 && crop_x == 0 && crop_h == H && crop_w == W && oH == H && oW == W &&
      (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long n = total * 3;
    preprocess_u8_kernel<<<grid_for(n / 4 / 4 + 1), 256, 0, s>>>(src, dst, n);
  } else {
    ingest_frames_kernel<<<grid_for(total), 256, 0, s>>>(
        src, H, W, crop_y, crop_x, crop_h, crop_w, dst, oH, oW, total,
        (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  }
  count_launch();
  TAPIR_LAUNCH_CHECK("ingest_frames_kernel");
  return kOk;
}

int postprocess_occlusions(const float* occ, const float* expd, long long n, uint8_t* visible,
                           cudaStream_t s) {
  TAPIR_CHECK_ARG(occ && expd && visible && n > 0, "postprocess_occlusions: bad arguments");
  ProfileScope ps("postprocess_occlusions", s, 0.0, (double)n * 9);
  postprocess_occlusions_kernel<<<grid_for(n), 256, 0, s>>>(occ, expd, n, visible);
  count_launch();
  TAPIR_LAUNCH_CHECK("postprocess_occlusions_kernel");
  return kOk;
}

int tapvid_counts(const tapir_tapvid_args* a, cudaStream_t s) {
  TAPIR_CHECK_ARG(a != nullptr && a->B > 0 && a->N > 0 && a->T > 0, "tapvid_counts: bad shape");
  TAPIR_CHECK_ARG(a->query_points && a->gt_occluded && a->gt_tracks && a->pred_tracks && a->counts,
                  "tapvid_counts: null pointer");
  TAPIR_CHECK_ARG(a->pred_occluded != nullptr || (a->pred_occ_logits && a->pred_expd_logits),
                  "tapvid_counts: need pred_occluded or both logit arrays");
  TAPIR_CHECK_ARG(a->query_mode == 0 || a->query_mode == 1,
                  "tapvid_counts: query_mode must be 0 ('first') or 1 ('strided')");
  const long long tracks = (long long)a->B * a->N;
  ProfileScope ps("tapvid_counts", s, 0.0, (double)tracks * a->T * 18);
  const int warps_per_block = 8;
  tapvid_counts_kernel<<<(unsigned)ceil_div_ll(tracks, warps_per_block), 32 * warps_per_block, 0, s>>>(*a);
  count_launch();
  TAPIR_LAUNCH_CHECK("tapvid_counts_kernel");
  return kOk;
}

}  // namespace tapir
