// Parameter packing for one (direction, layer) — layout documented in include/dagnn_b200.h.
// Runs once per parameter version (not per forward): splits the GRU weights into fp16 hi / lo parts and lays them out
// as the exact shared-memory images (K-major SWIZZLE_128B tiles) the level kernel bulk-copies and hands to
// tcgen05.mma as the B operand, and folds the attention linear layer into a key vector + two edge-type coefficients.
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

// One projection image = the concatenation [W_a ; W_b] of up to two GRU weight matrices (each [3H, K], gate order r, z, n)
// that share the same K-wide operand, as the B operand of the projection GEMM  out[row, col] = sum_k a[row, k] * W[col, k]:
// columns of one matrix are laid out [r (Hq) | z (Hq) | n (Hq)] and zero-padded to Mc = roundup(3 Hq, 64); the image holds,
// for every block of 64 columns and every 64-wide k chunk, a fp16 hi tile then a lo tile, each [64 rows][128 B] in the
// K-major SWIZZLE_128B layout of tc.cuh (hi = rn_f16(w), lo = rn_f16(w - hi)).
__global__ void __launch_bounds__(256) k_pack_proj(const float* __restrict__ w_a, const float* __restrict__ w_b, int H, int K, int Hq,
                                                   int Mc, int nck, __half* __restrict__ img) {
  const int nblk = (w_b ? 2 : 1) * (Mc / 64);
  const int64_t total = (int64_t)nblk * nck * 64 * 64;          // one thread-iteration per (col block, chunk, row j, kk)
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % 64);
    const int j = (int)((idx / 64) % 64);
    const int c = (int)((idx / 4096) % nck);
    const int cb = (int)(idx / ((int64_t)4096 * nck));
    const int col = cb * 64 + j;
    const float* w = col < Mc ? w_a : w_b;
    const int cm = col < Mc ? col : col - Mc;                   // column inside its matrix: gate * Hq + unit
    const int g = cm / Hq, u = cm - g * Hq;
    const int k = c * 64 + kk;
    float v = 0.f;
    if (g < 3 && u < H && k < K) v = w[((size_t)g * H + u) * K + k];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t tile = ((size_t)cb * nck + c) * 2 * 4096;
    const uint32_t off = (uint32_t)(j >> 3) * 512u + (uint32_t)(j & 7) * 64u + ((uint32_t)((kk >> 3) ^ (j & 7)) << 3) + (uint32_t)(kk & 7);
    img[tile + off] = hi;
    img[tile + 4096 + off] = lo;
  }
}

__global__ void __launch_bounds__(256) k_pack_small(const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                    const float* __restrict__ attn_w, int Dq, const float* __restrict__ edge_w,
                                                    DagnnPackLayout L, float* __restrict__ packed) {
  const int HP = L.HP;
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

extern "C" int dagnn_pack_layout(int32_t Din, int32_t H, int32_t nvid, int32_t first_layer, int32_t last_layer, DagnnPackLayout* out) {
  DAGNN_REQUIRE(out && Din > 0 && H > 0 && nvid >= 0, "pack_layout args");
  DagnnPackLayout L;
  memset(&L, 0, sizeof(L));
  L.Din = Din; L.H = H; L.nvid = nvid; L.first_layer = first_layer ? 1 : 0; L.last_layer = last_layer ? 1 : 0;
  L.Hq = round_up(H, 4);
  L.Mc = round_up(3 * L.Hq, 64);
  L.Kin64 = round_up(Din, 64);
  L.Kh64 = round_up(H, 64);
  L.HP = round_up(H, 64);
  int64_t off = 0;
  L.bias_off = off;  off += 4 * (int64_t)L.HP;
  L.wk_off = off;    off += L.HP;
  L.attnc_off = off; off += 4;
  L.vidk_off = off;  off += round_up64(nvid, 4);
  off = round_up64(off, 256);                                  // images 1024-byte aligned inside the blob
  // imgx: W_ih of this layer over its input (only layer 0 needs it: deeper layers' W_ih rides in the previous layer's imgh)
  L.imgx_off = off;  if (L.first_layer) off += (int64_t)(L.Mc / 64) * (L.Kin64 / 64) * 2 * 4096 / 2;     // halfs -> 4-byte units
  // imgh: [W_hh of this layer ; W_ih of the next layer] over this layer's state
  L.imgh_off = off;  off += (int64_t)((L.last_layer ? 1 : 2) * (L.Mc / 64)) * (L.Kh64 / 64) * 2 * 4096 / 2;
  L.total_floats = off;
  *out = L;
  return DAGNN_OK;
}

extern "C" int dagnn_pack_params_f32(const float* weight_ih, const float* weight_hh, const float* bias_ih, const float* bias_hh,
                                     const float* attn_w, int32_t Dq, const float* edge_w, const float* weight_ih_next,
                                     const DagnnPackLayout* layout, float* packed, void* stream_) {
  DAGNN_REQUIRE(weight_ih && weight_hh && bias_ih && bias_hh && attn_w && layout && packed, "pack_params: null pointer");
  DAGNN_REQUIRE(Dq >= 0, "pack_params: Dq");
  DAGNN_REQUIRE((((uintptr_t)packed) & 15) == 0, "pack_params: packed must be 16-byte aligned");
  const DagnnPackLayout& L = *layout;
  DAGNN_REQUIRE(L.last_layer || weight_ih_next, "pack_params: weight_ih_next is required unless this is the last layer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  auto launch = [&](const float* wa, const float* wb, int K, int nck, int64_t off) -> int {
    const int64_t tot = (int64_t)(wb ? 2 : 1) * (L.Mc / 64) * nck * 4096;
    const int blocks = (int)((tot + 255) / 256 < 148 * 8 ? (tot + 255) / 256 : 148 * 8);
    k_pack_proj<<<blocks, 256, 0, st>>>(wa, wb, L.H, K, L.Hq, L.Mc, nck, reinterpret_cast<__half*>(packed + off));
    return check_launch("k_pack_proj");
  };
  if (L.first_layer)
    if (int rc = launch(weight_ih, nullptr, L.Din, L.Kin64 / 64, L.imgx_off)) return rc;
  if (int rc = launch(weight_hh, L.last_layer ? nullptr : weight_ih_next, L.H, L.Kh64 / 64, L.imgh_off)) return rc;
  k_pack_small<<<1, 256, 0, st>>>(bias_ih, bias_hh, attn_w, Dq, edge_w, L, packed);
  return check_launch("k_pack_small");
}
