// Stage A (SURVEY.md 8(a4),(a6)): query-feature sampling, global cost volume and the
// track / occlusion head.
//
// The cost volume itself is the split-bf16 tcgen05 GEMM (planes=3, i.e. fp32-equivalent,
// because its arg-max must match the reference bit-exactly): cv[n][t*h*w + cell] with
// M = queries, N = T*h*w cells, K = 256.  `cost_volume_head_kernel` then consumes one
// 32x32 map per CTA entirely from shared memory: conv3x3 1->16 + ReLU, conv3x3 16->1,
// temperature soft-max, arg-max, radius-5 soft arg-max, and the occlusion branch
// (pad(0,2,0,2), conv3x3 stride 2 16->32 + ReLU, spatial mean, 32->16->2 MLP).
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace tapir {

namespace {

// ------------------------------------------------------------------------ a4
// utils.py:45-73 map_coordinates_3d: grid_sample(5-D, bilinear, align_corners=False,
// padding_mode='border') at (t+0.5, y, x) / (T, h, w).  We mirror ATen's arithmetic:
// g = 2*(c/size) - 1 ; pix = ((g + 1) * size - 1) / 2 ; clamp to [0, size-1].
__device__ __forceinline__ float unnormalize_border(float c, int size) {
  const float fs = (float)size;
  const float g = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(c, fs)), 1.f);
  float pix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), fs), 1.f), 2.f);
  pix = fminf(fmaxf(pix, 0.f), fs - 1.f);
  return pix;
}

__global__ void __launch_bounds__(128) sample_query_kernel(const float* __restrict__ grid, int T,
                                                           int gh, int gw, int C,
                                                           const float* __restrict__ q, int vT,
                                                           int vH, int vW, float* __restrict__ out) {
  const int n = blockIdx.x;
  // utils.convert_grid_coordinates: coords * out / in (tapir_model.py:266-277)
  float t = __fdiv_rn(__fmul_rn(q[n * 3 + 0], (float)T), (float)vT);
  const float y = __fdiv_rn(__fmul_rn(q[n * 3 + 1], (float)gh), (float)vH);
  const float x = __fdiv_rn(__fmul_rn(q[n * 3 + 2], (float)gw), (float)vW);
  t = __fadd_rn(t, 0.5f);
  const float pt = unnormalize_border(t, T), py = unnormalize_border(y, gh), px = unnormalize_border(x, gw);
  const int t0 = (int)floorf(pt), y0 = (int)floorf(py), x0 = (int)floorf(px);
  const float ft = pt - t0, fy = py - y0, fx = px - x0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int tt = t0 + dt, yy = y0 + dy, xx = x0 + dx;
          const float wgt = (dt ? ft : 1.f - ft) * (dy ? fy : 1.f - fy) * (dx ? fx : 1.f - fx);
          if (tt < T && yy < gh && xx < gw && tt >= 0 && yy >= 0 && xx >= 0)
            acc += wgt * grid[(((long long)tt * gh + yy) * gw + xx) * C + c];
        }
    out[(long long)n * C + c] = acc;
  }
}

// ------------------------------------------------------------------------ a6 head
constexpr int kG = 32;            // cost-volume map side (initial_resolution 256 / stride 8)
constexpr int kOccW = kG + 3;     // 35: index -1 .. 33 (1 left halo for hid2, 2 right for hid3)
constexpr int kOccPlane = kOccW * kOccW;

struct HeadSmem {
  float cv[(kG + 2) * (kG + 2)];     // zero-padded cost map
  float occ[16 * kOccPlane];         // ReLU(hid1) with halo, channel-major
  float w3[16 * 9 * 32];             // hid3 weights as [ci][tap][co]
  // hid1 / hid2 weights, 9 taps padded to 12 per channel: three broadcast LDS.128 per channel
  __align__(16) float w1[16 * 12];
  __align__(16) float w2[16 * 12];
  float b1[16], b3[32];
  float w4[16 * 32], b4[16], w5[2 * 16], b5[2];
  float red_f[8 * 4];
  int red_i[8];
  float mean32[8][32];
  float bcast[4];
  int bcast_i;
};

__device__ __forceinline__ void block_reduce_max_first(float v, int idx, HeadSmem& sm, float* out_v,
                                                       int* out_i) {
  // maximum value; among equal values the LOWEST index (torch.argmax on CPU)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { sm.red_f[warp] = v; sm.red_i[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = sm.red_f[0];
    int bi = sm.red_i[0];
    for (int k = 1; k < 8; ++k)
      if (sm.red_f[k] > bv || (sm.red_f[k] == bv && sm.red_i[k] < bi)) { bv = sm.red_f[k]; bi = sm.red_i[k]; }
    sm.bcast[0] = bv;
    sm.bcast_i = bi;
  }
  __syncthreads();
  *out_v = sm.bcast[0];
  *out_i = sm.bcast_i;
}

__device__ __forceinline__ void block_reduce_sum3(float a, float b, float c, HeadSmem& sm, float* oa,
                                                  float* ob, float* oc) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { sm.red_f[warp * 4] = a; sm.red_f[warp * 4 + 1] = b; sm.red_f[warp * 4 + 2] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0, y = 0, z = 0;
    for (int k = 0; k < 8; ++k) { x += sm.red_f[k * 4]; y += sm.red_f[k * 4 + 1]; z += sm.red_f[k * 4 + 2]; }
    sm.bcast[0] = x; sm.bcast[1] = y; sm.bcast[2] = z;
  }
  __syncthreads();
  *oa = sm.bcast[0]; *ob = sm.bcast[1]; *oc = sm.bcast[2];
}

__global__ void __launch_bounds__(256) cost_volume_head_kernel(
    const tapir_head_weights w, const float* __restrict__ cost_volume, int T,
    const float* __restrict__ query_tyx, float temperature, int init_h, int init_w,
    float* __restrict__ points, float* __restrict__ occ_out, float* __restrict__ expd_out,
    int* __restrict__ argmax_out) {
  extern __shared__ __align__(16) uint8_t head_smem_raw[];
  HeadSmem& sm = *reinterpret_cast<HeadSmem*>(head_smem_raw);
  const int t = blockIdx.x, n = blockIdx.y;
  const int tid = threadIdx.x;
  const float* cv = cost_volume + ((long long)n * T + t) * (kG * kG);

  // ---- stage weights and the zero-padded map
  for (int i = tid; i < (kG + 2) * (kG + 2); i += 256) sm.cv[i] = 0.f;
  for (int i = tid; i < 16 * kOccPlane; i += 256) sm.occ[i] = 0.f;
  for (int i = tid; i < 16 * 9 * 32; i += 256) {
    const int co = i & 31, tap = (i >> 5) % 9, ci = i / (32 * 9);
    sm.w3[i] = w.hid3_w[(co * 16 + ci) * 9 + tap];
  }
  if (tid < 192) {
    const int co = tid / 12, k = tid - co * 12;
    sm.w1[tid] = (k < 9) ? w.hid1_w[co * 9 + k] : 0.f;
    sm.w2[tid] = (k < 9) ? w.hid2_w[co * 9 + k] : 0.f;
  }
  if (tid < 16) { sm.b1[tid] = w.hid1_b[tid]; sm.b4[tid] = w.hid4_b[tid]; }
  if (tid < 32) { sm.b3[tid] = w.hid3_b[tid]; sm.w5[tid] = w.occ_w[tid]; }
  if (tid < 2) sm.b5[tid] = w.occ_b[tid];
  for (int i = tid; i < 512; i += 256) sm.w4[i] = w.hid4_w[i];
  __syncthreads();
  for (int i = tid; i < kG * kG; i += 256) sm.cv[((i >> 5) + 1) * (kG + 2) + (i & 31) + 1] = cv[i];
  __syncthreads();

  // ---- hid1: conv3x3 1->16 (padding 1) + ReLU.  Thread = row y, 4 consecutive x.
  const int y = tid >> 3, x0 = (tid & 7) * 4;
  {
    float win[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) win[r][c] = sm.cv[(y + r) * (kG + 2) + x0 + c];
    for (int co = 0; co < 16; ++co) {
      float a[4] = {sm.b1[co], sm.b1[co], sm.b1[co], sm.b1[co]};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float wv = sm.w1[co * 12 + ky * 3 + kx];
#pragma unroll
          for (int p = 0; p < 4; ++p) a[p] = fmaf(win[ky][kx + p], wv, a[p]);
        }
#pragma unroll
      for (int p = 0; p < 4; ++p) sm.occ[co * kOccPlane + (y + 1) * kOccW + (x0 + p + 1)] = fmaxf(a[p], 0.f);
    }
  }
  __syncthreads();

  // ---- hid2: conv3x3 16->1 (padding 1), x temperature
  float heat[4];
  {
    const float b2 = w.hid2_b[0];
    float a[4] = {b2, b2, b2, b2};
    for (int ci = 0; ci < 16; ++ci) {
      const float* op = sm.occ + ci * kOccPlane;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float row[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) row[c] = op[(y + ky) * kOccW + x0 + c];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float wv = sm.w2[ci * 12 + ky * 3 + kx];
#pragma unroll
          for (int p = 0; p < 4; ++p) a[p] = fmaf(row[kx + p], wv, a[p]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) heat[p] = a[p] * temperature;
  }

  // ---- softmax over the 1024 cells, argmax of the PROBABILITIES (first index on ties),
  //      soft arg-max within radius 5 (utils.py:116-150)
  float lmax = fmaxf(fmaxf(heat[0], heat[1]), fmaxf(heat[2], heat[3]));
  int dummy;
  float gmax;
  block_reduce_max_first(lmax, 0, sm, &gmax, &dummy);
  float e[4], lsum = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) { e[p] = expf(heat[p] - gmax); lsum += e[p]; }
  float gsum, u1, u2;
  block_reduce_sum3(lsum, 0.f, 0.f, sm, &gsum, &u1, &u2);
  float prob[4];
  float pbest = -1.f;
  int ibest = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    prob[p] = e[p] / gsum;
    if (prob[p] > pbest) { pbest = prob[p]; ibest = y * kG + x0 + p; }
  }
  float pm;
  int am;
  block_reduce_max_first(pbest, ibest, sm, &pm, &am);
  const float cx = (float)(am & 31) + 0.5f, cy = (float)(am >> 5) + 0.5f;
  float sx = 0.f, sy = 0.f, sw = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float px = (float)(x0 + p) + 0.5f, py = (float)y + 0.5f;
    const float dx = px - cx, dy = py - cy;
    if (dx * dx + dy * dy < 25.f) { sx += px * prob[p]; sy += py * prob[p]; sw += prob[p]; }
  }
  float tx, ty, tw;
  block_reduce_sum3(sx, sy, sw, sm, &tx, &ty, &tw);

  // ---- hid3: pad(0,2,0,2), conv3x3 stride 2 16->32 + ReLU; thread = one of 16x16 outputs
  {
    const int oy = tid >> 4, ox = tid & 15;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = sm.b3[c];
    for (int ci = 0; ci < 16; ++ci) {
      const float* op = sm.occ + ci * kOccPlane + (2 * oy + 1) * kOccW + (2 * ox + 1);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = op[ky * kOccW + kx];
          const float4* wr = reinterpret_cast<const float4*>(sm.w3 + (ci * 9 + ky * 3 + kx) * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 w4 = wr[g];
            acc[4 * g + 0] = fmaf(v, w4.x, acc[4 * g + 0]);
            acc[4 * g + 1] = fmaf(v, w4.y, acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(v, w4.z, acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(v, w4.w, acc[4 * g + 3]);
          }
        }
    }
    // spatial mean over the 256 outputs: warp-level transpose-reduce, then across 8 warps
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = fmaxf(acc[c], 0.f);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
      for (int j = 0; j < s; ++j) {
        const bool up = (lane & s) != 0;
        const float send = up ? acc[j] : acc[j + s];
        const float keep = up ? acc[j + s] : acc[j];
        acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    sm.mean32[warp][lane] = acc[0];  // lane l holds channel l summed over the warp
  }
  __syncthreads();
  if (tid < 32) {
    float m = 0.f;
    for (int k = 0; k < 8; ++k) m += sm.mean32[k][tid];
    sm.mean32[0][tid] = m * (1.0f / 256.0f);
  }
  __syncthreads();
  if (tid < 16) {
    float a = sm.b4[tid];
    for (int k = 0; k < 32; ++k) a = fmaf(sm.mean32[0][k], sm.w4[tid * 32 + k], a);
    sm.mean32[1][tid] = fmaxf(a, 0.f);
  }
  __syncthreads();
  if (tid < 2) {
    float a = sm.b5[tid];
    for (int k = 0; k < 16; ++k) a = fmaf(sm.mean32[1][k], sm.w5[tid * 16 + k], a);
    const long long o = (long long)n * T + t;
    if (tid == 0) occ_out[o] = a; else expd_out[o] = a;
  }
  if (tid == 0) {
    const long long o = (long long)n * T + t;
    const float den = fmaxf(tw, 1e-12f);
    // utils.py:165-169: cell units -> initial_resolution pixels (coords * out / in)
    float px = __fdiv_rn(__fmul_rn(tx / den, (float)init_w), (float)kG);
    float py = __fdiv_rn(__fmul_rn(ty / den, (float)init_h), (float)kG);
    if (query_tyx != nullptr) {
      // utils.py:171-191: on the query frame the track passes through the query point
      const float qf = rintf(query_tyx[n * 3 + 0]);
      if (qf == (float)t) { px = query_tyx[n * 3 + 2]; py = query_tyx[n * 3 + 1]; }
    }
    points[o * 2 + 0] = px;
    points[o * 2 + 1] = py;
    if (argmax_out != nullptr) argmax_out[o] = am;
  }
}


// ------------------------------------------------------------------------ a6 head, tensor-core hid3
// Same arithmetic as cost_volume_head_kernel, re-organised so that the 16->32 stride-2 conv
// (80 % of the head FLOPs) runs on the tensor cores as a split-bf16 implicit GEMM per map:
// M = 256 output pixels, N = 32 channels, K = 9 taps x 16 input channels (one m16n8k16 k-step per
// tap), three MMAs per product (hi*hi + hi*lo + lo*hi), fp32 accumulation.  ReLU(hid1) is kept as
// bf16 hi/lo planes in channel-last order (A fragments are plain 32-bit shared loads), and the
// 16->1 conv is evaluated in exact fp32 from registers as nine per-tap channel dot products
// (`stap`) that are then gathered - so the soft-argmax path never sees bf16.
constexpr int kHT = 512;                       // threads per map
constexpr int kPixW = 9;                       // 32-bit words per pixel: 16 bf16 + 1 pad (bank spread)
constexpr int kPlaneWords = kOccPlane * kPixW; // 35 x 35 pixels
constexpr int kTapW = (kG + 2) * (kG + 2);

// W3S = 32-bit words per hid3 weight row.  With the pixel stride (9) every weight-fragment load
// of phase D was 2-way bank conflicted (lane group g = 7 lands on the banks of g = 0: 49 M of the
// 145 M shared wavefronts per launch, profiles/r01_ncu_summary.md); 12 makes {12 g + tq} distinct
// mod 32.  Validated against the whole parity suite in round 2 (gpurun_out/head_w3s.log).
constexpr int kW3Stride = 12;
template <int W3S>
struct HeadTcSmem {
  float cv[kTapW];
  float stap[9 * kTapW];
  uint32_t occ_hi[kPlaneWords], occ_lo[kPlaneWords];
  uint32_t w3_hi[9 * 32 * W3S], w3_lo[9 * 32 * W3S];  // [tap][co][ci pairs]
  // hid1 / hid2 weights, 9 taps padded to 12 per channel: three broadcast LDS.128 per channel
  __align__(16) float w1[16 * 12];
  __align__(16) float w2[16 * 12];
  float b1[16], b3[32];
  float w4[16 * 32], b4[16], w5[2 * 16], b5[2];
  float red_f[16 * 4];
  int red_i[16];
  float mean32[16][32];
  float bcast[4];
  int bcast_i;
  // one slot per block reduction of a map (4 per map): a slot is rewritten only in the next map,
  // at least one __syncthreads later, so every reduction needs a single barrier
  float slot_f[4][16 * 3];
  int slot_i[4][16];
};

// Single-barrier block reductions (16 warps): warp results go to a private slot, then EVERY thread
// folds the 16 partials itself in the same fixed order (16 broadcast LDS instead of two more
// barriers; the head kernel spent 21 % of its stall samples on barriers).
template <class SM>
__device__ __forceinline__ void block_max_first_1s(float v, int idx, SM& sm, int slot, float* out_v,
                                                   int* out_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm.slot_f[slot][warp] = v; sm.slot_i[slot][warp] = idx; }
  __syncthreads();
  float bv = sm.slot_f[slot][0];
  int bi = sm.slot_i[slot][0];
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    const float kv = sm.slot_f[slot][k];
    const int ki = sm.slot_i[slot][k];
    if (kv > bv || (kv == bv && ki < bi)) { bv = kv; bi = ki; }
  }
  *out_v = bv;
  *out_i = bi;
}

template <class SM>
__device__ __forceinline__ void block_sum3_1s(float a, float b, float c, SM& sm, int slot, float* oa,
                                              float* ob, float* oc) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm.slot_f[slot][warp * 3] = a; sm.slot_f[slot][warp * 3 + 1] = b; sm.slot_f[slot][warp * 3 + 2] = c; }
  __syncthreads();
  float x = 0, y = 0, z = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) { x += sm.slot_f[slot][k * 3]; y += sm.slot_f[slot][k * 3 + 1]; z += sm.slot_f[slot][k * 3 + 2]; }
  *oa = x; *ob = y; *oc = z;
}

template <class SM>
__device__ __forceinline__ void block_max_first(float v, int idx, SM& sm, int nwarps, float* out_v,
                                                int* out_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { sm.red_f[warp] = v; sm.red_i[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = sm.red_f[0];
    int bi = sm.red_i[0];
    for (int k = 1; k < nwarps; ++k)
      if (sm.red_f[k] > bv || (sm.red_f[k] == bv && sm.red_i[k] < bi)) { bv = sm.red_f[k]; bi = sm.red_i[k]; }
    sm.bcast[0] = bv;
    sm.bcast_i = bi;
  }
  __syncthreads();
  *out_v = sm.bcast[0];
  *out_i = sm.bcast_i;
}

template <class SM>
__device__ __forceinline__ void block_sum3(float a, float b, float c, SM& sm, int nwarps, float* oa,
                                           float* ob, float* oc) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { sm.red_f[warp * 4] = a; sm.red_f[warp * 4 + 1] = b; sm.red_f[warp * 4 + 2] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0, y = 0, z = 0;
    for (int k = 0; k < nwarps; ++k) { x += sm.red_f[k * 4]; y += sm.red_f[k * 4 + 1]; z += sm.red_f[k * 4 + 2]; }
    sm.bcast[0] = x; sm.bcast[1] = y; sm.bcast[2] = z;
  }
  __syncthreads();
  *oa = sm.bcast[0]; *ob = sm.bcast[1]; *oc = sm.bcast[2];
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

static_assert(sizeof(HeadTcSmem<kW3Stride>) <= 227 * 1024,
              "head kernel shared memory exceeds the 227 KB opt-in limit");

template <int W3S>
__global__ void __launch_bounds__(kHT, 1) cost_volume_head_tc_kernel(
    const tapir_head_weights w, const float* __restrict__ cost_volume, int T, int num_maps,
    const float* __restrict__ query_tyx, float temperature, int init_h, int init_w,
    float* __restrict__ points, float* __restrict__ occ_out, float* __restrict__ expd_out,
    int* __restrict__ argmax_out) {
  extern __shared__ __align__(16) uint8_t head_tc_smem_raw[];
  HeadTcSmem<W3S>& sm = *reinterpret_cast<HeadTcSmem<W3S>*>(head_tc_smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- phase A (once per persistent CTA): zero halos, stage weights (hid3 weights as bf16
  // hi/lo planes [tap][co][ci]); the interiors are fully rewritten for every map
  for (int i = tid; i < kTapW; i += kHT) sm.cv[i] = 0.f;
  for (int i = tid; i < 9 * kTapW; i += kHT) sm.stap[i] = 0.f;
  for (int i = tid; i < kPlaneWords; i += kHT) { sm.occ_hi[i] = 0u; sm.occ_lo[i] = 0u; }
  for (int i = tid; i < 9 * 32 * 8; i += kHT) {
    const int cp = i & 7, co = (i >> 3) & 31, tap = i >> 8;
    float v0 = w.hid3_w[(co * 16 + 2 * cp) * 9 + tap], v1 = w.hid3_w[(co * 16 + 2 * cp + 1) * 9 + tap];
    const uint32_t hi = bf16x2_split(v0, v1);
    const uint32_t lo = bf16x2_split(v0, v1);
    sm.w3_hi[(tap * 32 + co) * W3S + cp] = hi;
    sm.w3_lo[(tap * 32 + co) * W3S + cp] = lo;
  }
  if (tid < 192) {
    const int co = tid / 12, k = tid - co * 12;
    sm.w1[tid] = (k < 9) ? w.hid1_w[co * 9 + k] : 0.f;
    sm.w2[tid] = (k < 9) ? w.hid2_w[co * 9 + k] : 0.f;
  }
  if (tid < 16) { sm.b1[tid] = w.hid1_b[tid]; sm.b4[tid] = w.hid4_b[tid]; }
  if (tid < 32) { sm.b3[tid] = w.hid3_b[tid]; sm.w5[tid] = w.occ_w[tid]; }
  if (tid < 2) sm.b5[tid] = w.occ_b[tid];
  sm.w4[tid] = w.hid4_w[tid];

  // the next map's 4 KB are fetched into registers (2 floats per thread) while the current map
  // is processed, so the L2/HBM latency is not exposed between maps
  float2 pre = make_float2(0.f, 0.f);
  if ((int)blockIdx.x < num_maps)
    pre = reinterpret_cast<const float2*>(cost_volume + (long long)blockIdx.x * (kG * kG))[tid];
  __syncthreads();  // phase A's zero fill of `cv` is ordered before the first map's interior writes
  for (int map = blockIdx.x; map < num_maps; map += gridDim.x) {
  const int n = map / T, t = map - n * T;
  // (no barrier here: `cv` was last read in the previous map's phase B, several barriers ago;
  // the barrier below publishes the new map and, on the first trip, the staged weights)
  {
    const int i = 2 * tid;  // pixels i, i+1 of the 32x32 map (same row)
    float* d = sm.cv + ((i >> 5) + 1) * (kG + 2) + (i & 31) + 1;
    d[0] = pre.x;
    d[1] = pre.y;
    const int nxt = map + gridDim.x;
    if (nxt < num_maps) pre = reinterpret_cast<const float2*>(cost_volume + (long long)nxt * (kG * kG))[tid];
  }
  __syncthreads();

  // ---- phase B: hid1 (conv3x3 1->16, padding 1, ReLU) for 2 adjacent pixels of one row; the
  // 16 channel values go to the bf16 planes and into the nine per-tap dot products of hid2
  const int y = tid >> 4, x0 = (tid & 15) * 2;
  {
    float win[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) win[r][c] = sm.cv[(y + r) * (kG + 2) + x0 + c];
    float st[2][9];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int k = 0; k < 9; ++k) st[p][k] = 0.f;
    const int pix0 = (y + 1) * kOccW + (x0 + 1);
#pragma unroll 2
    for (int cp = 0; cp < 8; ++cp) {
      float o[2][2];  // [pixel][channel of the pair]
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int co = 2 * cp + e;
        float a0 = sm.b1[co], a1 = a0;
        float wk[12];
#pragma unroll
        for (int v = 0; v < 3; ++v)
          *reinterpret_cast<float4*>(wk + 4 * v) = reinterpret_cast<const float4*>(sm.w1 + co * 12)[v];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = wk[ky * 3 + kx];
            a0 = fmaf(win[ky][kx], wv, a0);
            a1 = fmaf(win[ky][kx + 1], wv, a1);
          }
        o[0][e] = fmaxf(a0, 0.f);
        o[1][e] = fmaxf(a1, 0.f);
#pragma unroll
        for (int v = 0; v < 3; ++v)
          *reinterpret_cast<float4*>(wk + 4 * v) = reinterpret_cast<const float4*>(sm.w2 + co * 12)[v];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float w2v = wk[k];
          st[0][k] = fmaf(o[0][e], w2v, st[0][k]);
          st[1][k] = fmaf(o[1][e], w2v, st[1][k]);
        }
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        float v0 = o[p][0], v1 = o[p][1];
        const uint32_t hi = bf16x2_split(v0, v1);
        const uint32_t lo = bf16x2_split(v0, v1);
        sm.occ_hi[(pix0 + p) * kPixW + cp] = hi;
        sm.occ_lo[(pix0 + p) * kPixW + cp] = lo;
      }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int k = 0; k < 9; ++k) sm.stap[k * kTapW + (y + 1) * (kG + 2) + (x0 + p + 1)] = st[p][k];
  }
  __syncthreads();

  // ---- phase C: hid2 = b2 + sum over taps of the neighbours' per-tap dot products; softmax,
  // arg-max of the probabilities (lowest index on ties), radius-5 soft arg-max (utils.py:116-150)
  float heat[2];
  {
    const float b2 = w.hid2_b[0];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      float a = b2;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          a += sm.stap[(ky * 3 + kx) * kTapW + (y + ky) * (kG + 2) + (x0 + p + kx)];
      heat[p] = a * temperature;
    }
  }
  int dummy;
  float gmax;
  block_max_first_1s(fmaxf(heat[0], heat[1]), 0, sm, 0, &gmax, &dummy);
  float e[2];
  e[0] = expf(heat[0] - gmax);
  e[1] = expf(heat[1] - gmax);
  float gsum, u1, u2;
  block_sum3_1s(e[0] + e[1], 0.f, 0.f, sm, 1, &gsum, &u1, &u2);
  float prob[2];
  float pbest = -1.f;
  int ibest = 0;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    prob[p] = e[p] / gsum;
    if (prob[p] > pbest) { pbest = prob[p]; ibest = y * kG + x0 + p; }
  }
  float pm;
  int am;
  block_max_first_1s(pbest, ibest, sm, 2, &pm, &am);
  const float cx = (float)(am & 31) + 0.5f, cy = (float)(am >> 5) + 0.5f;
  float sx = 0.f, sy = 0.f, sw = 0.f;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float px = (float)(x0 + p) + 0.5f, py = (float)y + 0.5f;
    const float dx = px - cx, dy = py - cy;
    if (dx * dx + dy * dy < 25.f) { sx += px * prob[p]; sy += py * prob[p]; sw += prob[p]; }
  }
  float tx, ty, tw;
  block_sum3_1s(sx, sy, sw, sm, 3, &tx, &ty, &tw);

  // ---- phase D: hid3 (pad(0,2,0,2), conv3x3 stride 2, 16->32) on the tensor cores.
  // warp = output row oy (16 output pixels = one m16 tile), 4 n8 tiles, one k16 step per tap.
  {
    const int oy = warp, g = lane >> 2, tq = lane & 3;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - ky * 3;
      const int pa = ((2 * oy + ky + 1) * kOccW + (2 * g + kx + 1)) * kPixW;  // output pixel ox = g
      const int pb = pa + 16 * kPixW;                                         // ox = g + 8
      uint32_t ah[4], al[4];
      ah[0] = sm.occ_hi[pa + tq]; ah[1] = sm.occ_hi[pb + tq];
      ah[2] = sm.occ_hi[pa + tq + 4]; ah[3] = sm.occ_hi[pb + tq + 4];
      al[0] = sm.occ_lo[pa + tq]; al[1] = sm.occ_lo[pb + tq];
      al[2] = sm.occ_lo[pa + tq + 4]; al[3] = sm.occ_lo[pb + tq + 4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int wb = (tap * 32 + nt * 8 + g) * W3S + tq;
        uint32_t bh[2] = {sm.w3_hi[wb], sm.w3_hi[wb + 4]};
        uint32_t bl[2] = {sm.w3_lo[wb], sm.w3_lo[wb + 4]};
        mma_bf16_16816(acc[nt], al, bh);
        mma_bf16_16816(acc[nt], ah, bl);
        mma_bf16_16816(acc[nt], ah, bh);
      }
    }
    // bias + ReLU, then the spatial sum of this warp's 16 pixels per channel: a thread owns
    // rows g, g+8 and channels nt*8 + 2*tq + {0,1}
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const float b = sm.b3[nt * 8 + 2 * tq + e2];
        float v = fmaxf(acc[nt][e2] + b, 0.f) + fmaxf(acc[nt][2 + e2] + b, 0.f);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) sm.mean32[warp][nt * 8 + 2 * tq + e2] = v;
      }
    }
  }
  __syncthreads();
  // The rest of the map (channel means, 32 -> 16 -> 2 MLP, outputs) is a few hundred FLOPs: warp 0
  // does it alone with shuffles while the other 15 warps already stage the next map (the mean32
  // rows are rewritten only in the next map's phase D, four barriers from here).
  if (warp == 0) {
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) m += sm.mean32[k][lane];
    m *= (1.0f / 256.0f);                       // lane = channel
    float a = (lane < 16) ? sm.b4[lane] : 0.f;  // hid4: 32 -> 16, ReLU
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float mk = __shfl_sync(0xffffffffu, m, k);
      if (lane < 16) a = fmaf(mk, sm.w4[lane * 32 + k], a);
    }
    a = fmaxf(a, 0.f);
    float o2 = (lane < 2) ? sm.b5[lane] : 0.f;  // occ_out: 16 -> 2
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float hk = __shfl_sync(0xffffffffu, a, k);
      if (lane < 2) o2 = fmaf(hk, sm.w5[lane * 16 + k], o2);
    }
    const long long o = (long long)n * T + t;
    if (lane == 0) occ_out[o] = o2;
    if (lane == 1) expd_out[o] = o2;
    if (lane == 0) {
      const float den = fmaxf(tw, 1e-12f);
      float px = __fdiv_rn(__fmul_rn(tx / den, (float)init_w), (float)kG);
      float py = __fdiv_rn(__fmul_rn(ty / den, (float)init_h), (float)kG);
      if (query_tyx != nullptr) {
        const float qf = rintf(query_tyx[n * 3 + 0]);
        if (qf == (float)t) { px = query_tyx[n * 3 + 2]; py = query_tyx[n * 3 + 1]; }
      }
      points[o * 2 + 0] = px;
      points[o * 2 + 1] = py;
      if (argmax_out != nullptr) argmax_out[o] = am;
    }
  }
  }  // map loop
}


// ------------------------------------------------------------------------ a6 head, any map size
// initial_resolution != (256, 256) (constructor argument, tapir_model.py:86) gives a cost map that
// is not 32 x 32.  Same arithmetic as the kernels above for a gh x gw map: one CTA per map,
// ReLU(hid1) and the heat map staged in a per-CTA slice of the caller's workspace (L2 resident).
// A correctness path (every published checkpoint / demo uses 256 x 256), not a tuned one.
struct GenericRed {
  float f[8 * 4];
  int i[8];
  float mean32[8][32];
  float out[4];
  int out_i;
};

__device__ __forceinline__ void generic_max_first(float v, int idx, GenericRed& r, float* ov, int* oi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float xv = __shfl_xor_sync(0xffffffffu, v, o);
    const int xi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (xv > v || (xv == v && xi < idx)) { v = xv; idx = xi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { r.f[warp] = v; r.i[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = r.f[0];
    int bi = r.i[0];
    for (int k = 1; k < 8; ++k)
      if (r.f[k] > bv || (r.f[k] == bv && r.i[k] < bi)) { bv = r.f[k]; bi = r.i[k]; }
    r.out[0] = bv;
    r.out_i = bi;
  }
  __syncthreads();
  *ov = r.out[0];
  *oi = r.out_i;
}

__device__ __forceinline__ void generic_sum3(float a, float b, float c, GenericRed& r, float* oa,
                                             float* ob, float* oc) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { r.f[warp * 4] = a; r.f[warp * 4 + 1] = b; r.f[warp * 4 + 2] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0, y = 0, z = 0;
    for (int k = 0; k < 8; ++k) { x += r.f[k * 4]; y += r.f[k * 4 + 1]; z += r.f[k * 4 + 2]; }
    r.out[0] = x; r.out[1] = y; r.out[2] = z;
  }
  __syncthreads();
  *oa = r.out[0]; *ob = r.out[1]; *oc = r.out[2];
}

__global__ void __launch_bounds__(256) cost_volume_head_generic_kernel(
    const tapir_head_weights w, const float* __restrict__ cost_volume, int T, int num_maps, int gh,
    int gw, const float* __restrict__ query_tyx, float temperature, int init_h, int init_w,
    float* __restrict__ scratch, float* __restrict__ points, float* __restrict__ occ_out,
    float* __restrict__ expd_out, int* __restrict__ argmax_out) {
  __shared__ GenericRed red;
  __shared__ float w1s[16 * 9], b1s[16], w2s[16 * 9], w3s[32 * 16 * 9], b3s[32], w4s[16 * 32], b4s[16],
      w5s[2 * 16], b5s[2];
  const int tid = threadIdx.x;
  const int pw = gw + 3, ph = gh + 3;  // index -1 .. size+1 (1 halo before, 2 after)
  const int plane = pw * ph;
  const int cells = gh * gw;
  float* occ = scratch + (size_t)blockIdx.x * (16 * plane + cells);  // [16][ph][pw]
  float* heat = occ + 16 * plane;                                    // [gh][gw]
  for (int i = tid; i < 16 * 9; i += 256) { w1s[i] = w.hid1_w[i]; w2s[i] = w.hid2_w[i]; }
  for (int i = tid; i < 32 * 16 * 9; i += 256) w3s[i] = w.hid3_w[i];
  for (int i = tid; i < 512; i += 256) w4s[i] = w.hid4_w[i];
  if (tid < 16) { b1s[tid] = w.hid1_b[tid]; b4s[tid] = w.hid4_b[tid]; }
  if (tid < 32) { b3s[tid] = w.hid3_b[tid]; w5s[tid] = w.occ_w[tid]; }
  if (tid < 2) b5s[tid] = w.occ_b[tid];
  for (int i = tid; i < 16 * plane; i += 256) occ[i] = 0.f;  // halos stay zero for every map
  const float b2 = w.hid2_b[0];
  const int oh = (gh - 1) / 2 + 1, ow = (gw - 1) / 2 + 1;  // pad (0,2,0,2), 3x3, stride 2, no padding
  for (int map = blockIdx.x; map < num_maps; map += gridDim.x) {
    const int n = map / T, t = map - n * T;
    const float* cv = cost_volume + (size_t)map * cells;
    __syncthreads();
    // hid1: conv3x3 1 -> 16, padding 1, ReLU
    for (int i = tid; i < cells; i += 256) {
      const int y = i / gw, x = i - y * gw;
      float win[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = y + ky - 1, xx = x + kx - 1;
          win[ky * 3 + kx] = (yy >= 0 && yy < gh && xx >= 0 && xx < gw) ? cv[yy * gw + xx] : 0.f;
        }
      for (int co = 0; co < 16; ++co) {
        float a = b1s[co];
#pragma unroll
        for (int k = 0; k < 9; ++k) a = fmaf(win[k], w1s[co * 9 + k], a);
        occ[co * plane + (y + 1) * pw + (x + 1)] = fmaxf(a, 0.f);
      }
    }
    __syncthreads();
    // hid2: conv3x3 16 -> 1, padding 1, x temperature; running maximum
    float lmax = -FLT_MAX;
    for (int i = tid; i < cells; i += 256) {
      const int y = i / gw, x = i - y * gw;
      float a = b2;
      for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            a = fmaf(occ[ci * plane + (y + ky) * pw + (x + kx)], w2s[ci * 9 + ky * 3 + kx], a);
      const float h = a * temperature;
      heat[i] = h;
      lmax = fmaxf(lmax, h);
    }
    int dummy;
    float gmax;
    generic_max_first(lmax, 0, red, &gmax, &dummy);
    float lsum = 0.f;
    for (int i = tid; i < cells; i += 256) lsum += expf(heat[i] - gmax);
    float gsum, u1, u2;
    generic_sum3(lsum, 0.f, 0.f, red, &gsum, &u1, &u2);
    float pbest = -1.f;
    int ibest = 0;
    for (int i = tid; i < cells; i += 256) {  // ascending i: the first maximum is kept
      const float pr = expf(heat[i] - gmax) / gsum;
      if (pr > pbest) { pbest = pr; ibest = i; }
    }
    float pm;
    int am;
    generic_max_first(pbest, ibest, red, &pm, &am);
    const float cx = (float)(am % gw) + 0.5f, cy = (float)(am / gw) + 0.5f;
    float sx = 0.f, sy = 0.f, sw = 0.f;
    for (int i = tid; i < cells; i += 256) {
      const int y = i / gw, x = i - y * gw;
      const float px = (float)x + 0.5f, py = (float)y + 0.5f;
      const float dx = px - cx, dy = py - cy;
      if (dx * dx + dy * dy < 25.f) {
        const float pr = expf(heat[i] - gmax) / gsum;
        sx += px * pr; sy += py * pr; sw += pr;
      }
    }
    float tx, ty, tw;
    generic_sum3(sx, sy, sw, red, &tx, &ty, &tw);
    // hid3: conv3x3 stride 2 16 -> 32 on the (0,2,0,2)-padded activation, ReLU, spatial mean
    float msum[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) msum[c] = 0.f;
    for (int o = tid; o < oh * ow; o += 256) {
      const int oy = o / ow, ox = o - oy * ow;
      float acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = b3s[c];
      for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float v = occ[ci * plane + (2 * oy + ky + 1) * pw + (2 * ox + kx + 1)];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = fmaf(v, w3s[(c * 16 + ci) * 9 + ky * 3 + kx], acc[c]);
          }
#pragma unroll
      for (int c = 0; c < 32; ++c) msum[c] += fmaxf(acc[c], 0.f);
    }
    {
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float v = warp_sum(msum[c]);
        if (lane == 0) red.mean32[warp][c] = v;
      }
    }
    __syncthreads();
    if (tid < 32) {
      float m = 0.f;
      for (int k = 0; k < 8; ++k) m += red.mean32[k][tid];
      red.mean32[0][tid] = m / (float)(oh * ow);
    }
    __syncthreads();
    if (tid < 16) {
      float a = b4s[tid];
      for (int k = 0; k < 32; ++k) a = fmaf(red.mean32[0][k], w4s[tid * 32 + k], a);
      red.mean32[1][tid] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (tid < 2) {
      float a = b5s[tid];
      for (int k = 0; k < 16; ++k) a = fmaf(red.mean32[1][k], w5s[tid * 16 + k], a);
      const long long o = (long long)n * T + t;
      if (tid == 0) occ_out[o] = a; else expd_out[o] = a;
    }
    if (tid == 0) {
      const long long o = (long long)n * T + t;
      const float den = fmaxf(tw, 1e-12f);
      float px = __fdiv_rn(__fmul_rn(tx / den, (float)init_w), (float)gw);
      float py = __fdiv_rn(__fmul_rn(ty / den, (float)init_h), (float)gh);
      if (query_tyx != nullptr) {
        const float qf = rintf(query_tyx[n * 3 + 0]);
        if (qf == (float)t) { px = query_tyx[n * 3 + 2]; py = query_tyx[n * 3 + 1]; }
      }
      points[o * 2 + 0] = px;
      points[o * 2 + 1] = py;
      if (argmax_out != nullptr) argmax_out[o] = am;
    }
  }
}

}  // namespace

int sample_query_features(const float* grid, int T, int gh, int gw, int C, const float* query_tyx,
                          int N, int vT, int vH, int vW, float* out, cudaStream_t s) {
  TAPIR_CHECK_ARG(grid && query_tyx && out && N > 0 && T > 0 && gh > 0 && gw > 0 && C > 0,
                  "sample_query_features: bad arguments");
  ProfileScope ps("sample_query", s, 0.0, (double)N * C * 4 * 9);
  sample_query_kernel<<<N, 128, 0, s>>>(grid, T, gh, gw, C, query_tyx, vT, vH, vW, out);
  count_launch();
  TAPIR_LAUNCH_CHECK("sample_query_kernel");
  return kOk;
}

int cost_volume_head(const tapir_head_weights* w, const float* cost_volume, int N, int T,
                     const float* query_tyx, float temperature, int init_h, int init_w,
                     float* points, float* occ, float* expd, int* argmax, cudaStream_t s) {
  static int use_simt = -1;
  static PerDeviceOnce configured;
  if (use_simt < 0) {
    const char* e = getenv("TAPIR_B200_HEAD");
    use_simt = (e != nullptr && strcmp(e, "simt") == 0) ? 1 : 0;
  }
  if (configured.pending()) {
    TAPIR_CUDA(cudaFuncSetAttribute(cost_volume_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(HeadSmem)));
    TAPIR_CUDA(cudaFuncSetAttribute(cost_volume_head_tc_kernel<kW3Stride>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(HeadTcSmem<kW3Stride>)));
    configured.mark();
  }
  TAPIR_CHECK_ARG(N <= 65535, "cost_volume_head: at most 65535 queries per call (got %d)", N);
  // head = 2,950,208 FLOP per (n,t); 4 KB map in, 16 B out
  ProfileScope ps("cost_volume.head", s, 2950208.0 * N * T, (double)N * T * (4096 + 16));
  dim3 grid(T, N);
  if (use_simt) {
    cost_volume_head_kernel<<<grid, 256, sizeof(HeadSmem), s>>>(*w, cost_volume, T, query_tyx, temperature,
                                                               init_h, init_w, points, occ, expd, argmax);
  } else {
    const int maps = N * T;
    const int ctas = maps < num_sms() ? maps : num_sms();
    cost_volume_head_tc_kernel<kW3Stride><<<ctas, kHT, sizeof(HeadTcSmem<kW3Stride>), s>>>(
        *w, cost_volume, T, maps, query_tyx, temperature, init_h, init_w, points, occ, expd, argmax);
  }
  count_launch();
  TAPIR_LAUNCH_CHECK("cost_volume_head_kernel");
  return kOk;
}

namespace {
constexpr int kCvPlanes = 3;
struct CvPlan {
  __nv_bfloat16* q;
  __nv_bfloat16* g;
  float* cv;
  float* generic_scratch;  // maps other than 32 x 32 only
  int generic_ctas;
};
constexpr int kGenericCtasPerSm = 4;
size_t plan_cv(Arena& a, int N, int T, int gh, int gw, int C, CvPlan* p) {
  const long long cells = (long long)T * gh * gw;
  p->q = a.take<__nv_bfloat16>((size_t)N * C * kCvPlanes);
  p->g = a.take<__nv_bfloat16>((size_t)cells * C * kCvPlanes);
  p->cv = a.take<float>((size_t)N * cells);
  p->generic_scratch = nullptr;
  p->generic_ctas = 0;
  if (gh != kG || gw != kG) {
    const long long maps = (long long)N * T;
    const long long cap = 148ll * kGenericCtasPerSm;  // independent of the device: sizes must agree
    p->generic_ctas = (int)(maps < cap ? maps : cap);
    p->generic_scratch =
        a.take<float>((size_t)p->generic_ctas * (16 * (size_t)(gh + 3) * (gw + 3) + (size_t)gh * gw));
  }
  return a.off;
}
}  // namespace

size_t cost_volume_workspace_bytes(int N, int T, int gh, int gw, int C) {
  Arena a(nullptr, 0);
  CvPlan p;
  return plan_cv(a, N, T, gh, gw, C, &p) + 256;
}

int cost_volume_tracks(const tapir_head_weights* w, const float* qfeat, const float* grid, int N,
                       int T, int gh, int gw, int C, const float* query_tyx, float temperature,
                       int init_h, int init_w, float* points, float* occ, float* expd, int* argmax,
                       void* ws, size_t ws_bytes, cudaStream_t s) {
  TAPIR_CHECK_ARG(w && qfeat && grid && points && occ && expd, "cost_volume_tracks: null pointer");
  TAPIR_CHECK_ARG(N > 0 && T > 0 && C % 64 == 0, "cost_volume_tracks: bad shape N=%d T=%d C=%d", N, T, C);
  TAPIR_CHECK_ARG(gh >= 2 && gw >= 2, "cost_volume_tracks: map %dx%d too small", gh, gw);
  Arena arena(ws, ws_bytes);
  CvPlan p;
  plan_cv(arena, N, T, gh, gw, C, &p);
  if (!arena.ok) {
    set_error("cost_volume_tracks: workspace too small (%zu < %zu)", ws_bytes, arena.off);
    return kWorkspaceTooSmall;
  }
  const long long cells = (long long)T * gh * gw;
  TAPIR_RETURN_IF(split_planes(qfeat, C, p.q, C, (long long)N * C, N, C, C, kCvPlanes, s));
  TAPIR_RETURN_IF(split_planes(grid, C, p.g, C, cells * C, cells, C, C, kCvPlanes, s));
  GemmArgs g;
  g.planes = kCvPlanes;
  g.M = N;
  g.N = (int)cells;
  g.K = C;
  g.a = p.q; g.lda = C; g.a_plane_stride = (long long)N * C;
  g.b = p.g; g.ldb = C; g.b_plane_stride = cells * C;
  g.out_f32 = p.cv; g.ldo = (int)cells;
  g.tag = "cost_volume.gemm";
  TAPIR_RETURN_IF(gemm(g, s));
  if (gh != kG || gw != kG) {
    // any other initial_resolution (tapir_model.py:86): generic-size head
    ProfileScope ps("cost_volume.head", s, 2950208.0 / 1024.0 * gh * gw * N * T,
                    (double)N * T * (4.0 * gh * gw + 16));
    cost_volume_head_generic_kernel<<<p.generic_ctas, 256, 0, s>>>(
        *w, p.cv, T, N * T, gh, gw, query_tyx, temperature, init_h, init_w, p.generic_scratch, points,
        occ, expd, argmax);
    count_launch();
    TAPIR_LAUNCH_CHECK("cost_volume_head_generic_kernel");
    return kOk;
  }
  return cost_volume_head(w, p.cv, N, T, query_tyx, temperature, init_h, init_w, points, occ, expd,
                          argmax, s);
}

}  // namespace tapir
