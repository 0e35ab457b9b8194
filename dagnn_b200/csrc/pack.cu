// Parameter packing for one (direction, layer) — layout documented in include/dagnn_b200.h.
// Runs once per parameter version (not per forward): transposes the GRU weights into the K-major,
// zero-padded, unit-sliced stream the level kernel bulk-copies, and folds the attention linear layer into
// a key vector + two edge-type coefficients.
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                      DagnnPackLayout L, float* __restrict__ packed) {
  const int K = L.Kin + L.Kh;
  const int64_t total = (int64_t)L.NS * K * 3 * DAGNN_UNIT_SLICE;
  float* w = packed + L.w_off;
  // [slice][k-block of 16][kq 4][gate 3][unit 32][k4 4]: a thread (= unit) reads float4 over k, a warp 512 contiguous bytes
  constexpr int kBlk = DAGNN_K_BLOCK * 3 * DAGNN_UNIT_SLICE;
  const int nkb = K / DAGNN_K_BLOCK;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k4 = (int)(idx % 4);
    const int u = (int)((idx / 4) % DAGNN_UNIT_SLICE);
    const int g = (int)((idx / (4 * DAGNN_UNIT_SLICE)) % 3);
    const int kq = (int)((idx / (12 * DAGNN_UNIT_SLICE)) % (DAGNN_K_BLOCK / 4));
    const int kb = (int)((idx / kBlk) % nkb);
    const int sl = (int)(idx / ((int64_t)kBlk * nkb));
    const int k = kb * DAGNN_K_BLOCK + kq * 4 + k4;
    const int unit = sl * DAGNN_UNIT_SLICE + u;
    float v = 0.f;
    if (unit < L.H) {
      if (k < L.Kin) {
        if (k < L.Din) v = w_ih[((size_t)g * L.H + unit) * L.Din + k];
      } else {
        const int kh = k - L.Kin;
        if (kh < L.H) v = w_hh[((size_t)g * L.H + unit) * L.H + kh];
      }
    }
    w[idx] = v;
  }
}

// tensor-core image of the GRU weights: per 64-unit tile `ut`, per 32-wide k chunk c (input chunks, then hidden chunks),
// a hi tile then a lo tile, each [192 gate-unit rows][128 B] in the K-major SWIZZLE_128B layout of tc.cuh — exactly the
// bytes the level kernel bulk-copies into shared memory and hands to tcgen05.mma as the B operand.
// Row order: input chunks [n | r | z] (D columns 0..191), hidden chunks [r | z | n] (D columns 64..255).
__global__ void __launch_bounds__(256) k_pack_tc(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                 DagnnPackLayout L, float* __restrict__ packed) {
  const int nci = L.Kin32 / 32, nch = L.Kh32 / 32, nc = nci + nch;
  const int64_t total = (int64_t)L.UT * nc * 192 * 32;          // one thread-iteration per (ut, c, row j, kk)
  float* img = packed + L.tc_off;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % 32);
    const int j = (int)((idx / 32) % 192);
    const int c = (int)((idx / (32 * 192)) % nc);
    const int ut = (int)(idx / ((int64_t)32 * 192 * nc));
    const int unit = ut * 64 + (j & 63);
    const int jb = j >> 6;
    float w = 0.f;
    if (unit < L.H) {
      if (c < nci) {
        const int g = jb == 0 ? 2 : jb - 1;
        const int k = c * 32 + kk;
        if (k < L.Din) w = w_ih[((size_t)g * L.H + unit) * L.Din + k];
      } else {
        const int g = jb;
        const int k = (c - nci) * 32 + kk;
        if (k < L.H) w = w_hh[((size_t)g * L.H + unit) * L.H + k];
      }
    }
    const float hi = tc::tf32_rn(w);
    const float lo = w - hi;
    const size_t tile = ((size_t)ut * nc + c) * 2 * (192 * 32);
    const uint32_t off = ((uint32_t)(j >> 3) * 1024u + (uint32_t)(j & 7) * 128u + ((uint32_t)((kk >> 2) ^ (j & 7)) << 4) + (uint32_t)(kk & 3) * 4u) >> 2;
    img[tile + off] = hi;
    img[tile + 192 * 32 + off] = lo;
  }
}

__global__ void __launch_bounds__(256) k_pack_small(const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                    const float* __restrict__ attn_w, int Dq, const float* __restrict__ edge_w,
                                                    DagnnPackLayout L, float* __restrict__ packed) {
  const int HP = L.NS * DAGNN_UNIT_SLICE;
  float* bias = packed + L.bias_off;
  float* wk = packed + L.wk_off;
  float* attnc = packed + L.attnc_off;
  float* vidk = packed + L.vidk_off;
  for (int u = threadIdx.x; u < HP; u += blockDim.x) {
    const bool in = u < L.H;
    bias[0 * HP + u] = in ? b_ih[u] + b_hh[u] : 0.f;
    bias[1 * HP + u] = in ? b_ih[L.H + u] + b_hh[L.H + u] : 0.f;
    bias[2 * HP + u] = in ? b_ih[2 * L.H + u] : 0.f;
    bias[3 * HP + u] = in ? b_hh[2 * L.H + u] : 0.f;
    wk[u] = in ? attn_w[Dq + u] : 0.f;
  }
  for (int j = threadIdx.x; j < L.nvid; j += blockDim.x) vidk[j] = attn_w[Dq + L.H + j];
  // attnc[c] = sum_u wk[u] * W_e[u, c]
  __shared__ float red[2][8];
  float s0 = 0.f, s1 = 0.f;
  if (edge_w) {
    for (int u = threadIdx.x; u < L.H; u += blockDim.x) {
      const float k = attn_w[Dq + u];
      s0 += k * edge_w[2 * u];
      s1 += k * edge_w[2 * u + 1];
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    attnc[0] = a; attnc[1] = b; attnc[2] = 0.f; attnc[3] = 0.f;
  }
}

}  // namespace dagnn

using namespace dagnn;

extern "C" int dagnn_pack_layout(int32_t Din, int32_t H, int32_t nvid, DagnnPackLayout* out) {
  DAGNN_REQUIRE(out && Din > 0 && H > 0 && nvid >= 0, "pack_layout args");
  DagnnPackLayout L;
  L.Din = Din; L.H = H; L.nvid = nvid;
  L.Kin = round_up(Din, DAGNN_K_BLOCK);
  L.Kh = round_up(H, DAGNN_K_BLOCK);
  L.NS = ceil_div(H, DAGNN_UNIT_SLICE);
  const int64_t HP = (int64_t)L.NS * DAGNN_UNIT_SLICE;
  int64_t off = 0;
  L.w_off = off;     off += (int64_t)L.NS * (L.Kin + L.Kh) * 3 * DAGNN_UNIT_SLICE;
  L.bias_off = off;  off += 4 * HP;
  L.wk_off = off;    off += HP;
  L.attnc_off = off; off += 4;
  L.vidk_off = off;  off += round_up64(nvid, 4);
  L.Kin32 = round_up(Din, 32);
  L.Kh32 = round_up(H, 32);
  L.UT = ceil_div(H, 64);
  off = round_up64(off, 256);                                  // 1024-byte aligned images (bulk copies need 16)
  L.tc_off = off;    off += (int64_t)L.UT * ((L.Kin32 + L.Kh32) / 32) * 2 * (192 * 32);
  L.total_floats = off;
  *out = L;
  return DAGNN_OK;
}

extern "C" int dagnn_pack_params_f32(const float* weight_ih, const float* weight_hh, const float* bias_ih, const float* bias_hh,
                                     const float* attn_w, int32_t Dq, const float* edge_w, const DagnnPackLayout* layout,
                                     float* packed, void* stream_) {
  DAGNN_REQUIRE(weight_ih && weight_hh && bias_ih && bias_hh && attn_w && layout && packed, "pack_params: null pointer");
  DAGNN_REQUIRE(Dq >= 0, "pack_params: Dq");
  DAGNN_REQUIRE((((uintptr_t)packed) & 15) == 0, "pack_params: packed must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int64_t total = (int64_t)layout->NS * (layout->Kin + layout->Kh) * 3 * DAGNN_UNIT_SLICE;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  k_pack_weights<<<blocks, 256, 0, st>>>(weight_ih, weight_hh, *layout, packed);
  if (int rc = check_launch("k_pack_weights")) return rc;
  const int64_t ttc = (int64_t)layout->UT * ((layout->Kin32 + layout->Kh32) / 32) * 192 * 32;
  k_pack_tc<<<(int)((ttc + 255) / 256 < 148 * 8 ? (ttc + 255) / 256 : 148 * 8), 256, 0, st>>>(weight_ih, weight_hh, *layout, packed);
  if (int rc = check_launch("k_pack_tc")) return rc;
  k_pack_small<<<1, 256, 0, st>>>(bias_ih, bias_hh, attn_w, Dq, edge_w, *layout, packed);
  return check_launch("k_pack_small");
}
