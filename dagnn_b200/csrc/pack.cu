// Parameter packing for one (direction, layer) — layout documented in include/dagnn_b200.h.
// Runs once per parameter version (not per forward): splits the GRU weights into fp16 hi / lo parts and lays them out
// as the exact shared-memory images (K-major SWIZZLE_128B tiles) the level kernel bulk-copies and hands to
// tcgen05.mma as the B operand, and folds the attention linear layer into a key vector + two edge-type coefficients.
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

// One image = for every block of U units (U = 16 or 64), for every 64-wide k chunk c (input chunks first, then hidden
// chunks): a hi tile then a lo tile, each [3U gate-unit rows][128 B = 64 halfs] in the layout of tc.cuh.
// Row order: input chunks [n | r | z] (TMEM columns 0..3U-1), hidden chunks [r | z | n] (TMEM columns U..4U-1).
__global__ void __launch_bounds__(256) k_pack_img(const float* __restrict__ w_ih, const float* __restrict__ w_hh, DagnnPackLayout L,
                                                  int U, int nblocks, __half* __restrict__ img) {
  const int nci = L.Kin64 / 64, nch = L.Kh64 / 64, nc = nci + nch;
  const int rows = 3 * U;
  const int64_t total = (int64_t)nblocks * nc * rows * 64;          // one thread-iteration per (block, c, row j, kk)
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % 64);
    const int j = (int)((idx / 64) % rows);
    const int c = (int)((idx / (64 * (int64_t)rows)) % nc);
    const int ub = (int)(idx / (64 * (int64_t)rows * nc));
    const int unit = ub * U + (j % U);
    const int jb = j / U;
    float w = 0.f;
    if (unit < L.H) {
      if (c < nci) {
        const int g = jb == 0 ? 2 : jb - 1;
        const int k = c * 64 + kk;
        if (k < L.Din) w = w_ih[((size_t)g * L.H + unit) * L.Din + k];
      } else {
        const int g = jb;
        const int k = (c - nci) * 64 + kk;
        if (k < L.H) w = w_hh[((size_t)g * L.H + unit) * L.H + k];
      }
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const size_t tile = ((size_t)ub * nc + c) * 2 * ((size_t)rows * 64);
    const uint32_t off = (uint32_t)(j >> 3) * 512u + (uint32_t)(j & 7) * 64u + ((uint32_t)((kk >> 3) ^ (j & 7)) << 3) + (uint32_t)(kk & 7);
    img[tile + off] = hi;
    img[tile + (size_t)rows * 64 + off] = lo;
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

extern "C" int dagnn_pack_layout(int32_t Din, int32_t H, int32_t nvid, DagnnPackLayout* out) {
  DAGNN_REQUIRE(out && Din > 0 && H > 0 && nvid >= 0, "pack_layout args");
  DagnnPackLayout L;
  memset(&L, 0, sizeof(L));
  L.Din = Din; L.H = H; L.nvid = nvid;
  L.Kin64 = round_up(Din, 64);
  L.Kh64 = round_up(H, 64);
  L.NG = ceil_div(H, 16);
  L.NT = ceil_div(H, 64);
  L.HP = L.NT * 64;
  const int64_t nc = (L.Kin64 + L.Kh64) / 64;
  int64_t off = 0;
  L.bias_off = off;  off += 4 * (int64_t)L.HP;
  L.wk_off = off;    off += L.HP;
  L.attnc_off = off; off += 4;
  L.vidk_off = off;  off += round_up64(nvid, 4);
  off = round_up64(off, 256);                                  // images 1024-byte aligned
  L.img16_off = off; off += (int64_t)L.NG * nc * 2 * (48 * 64) / 2;     // halfs -> 4-byte units
  L.img64_off = off; off += (int64_t)L.NT * nc * 2 * (192 * 64) / 2;
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
  const DagnnPackLayout& L = *layout;
  const int64_t nc = (L.Kin64 + L.Kh64) / 64;
  for (int pass = 0; pass < 2; ++pass) {
    const int U = pass == 0 ? 16 : 64, nb = pass == 0 ? L.NG : L.NT;
    const int64_t tot = (int64_t)nb * nc * 3 * U * 64;
    const int blocks = (int)((tot + 255) / 256 < 148 * 8 ? (tot + 255) / 256 : 148 * 8);
    __half* img = reinterpret_cast<__half*>(packed + (pass == 0 ? L.img16_off : L.img64_off));
    k_pack_img<<<blocks, 256, 0, st>>>(weight_ih, weight_hh, L, U, nb, img);
    if (int rc = check_launch("k_pack_img")) return rc;
  }
  k_pack_small<<<1, 256, 0, st>>>(bias_ih, bias_hh, attn_w, Dq, edge_w, L, packed);
  return check_launch("k_pack_small");
}
