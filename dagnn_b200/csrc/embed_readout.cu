// Node encoder (embedding gather-sum) and graph readout (masked segment pool) kernels — both pure HBM/L2
// streaming work: one coalesced pass, float4 where the layout allows.
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

// X[v,:] = T[x[v,0],:] + A[x[v,1],:] + P[min(depth[v],max_depth),:]      (ogbg-code/utils.py:26-28)
// one warp per node, lanes stride the feature dimension (float4 when D % 4 == 0)
template <bool VEC4>
__global__ void __launch_bounds__(256) k_embed(const int64_t* __restrict__ x, const int64_t* __restrict__ depth,
                                               const float* __restrict__ T, const float* __restrict__ A,
                                               const float* __restrict__ P, int max_depth, long long n_types, long long n_attrs,
                                               int N, int D, float* __restrict__ X, int64_t ldx,
                                               unsigned char* __restrict__ ximg) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int v = blockIdx.x * wpb + (threadIdx.x >> 5); v < N; v += gridDim.x * wpb) {
    const long long t = x[2 * (size_t)v], a = x[2 * (size_t)v + 1];
    long long dp = depth[v];
    dp = dp > max_depth ? max_depth : dp;
    float* o = X + (size_t)v * ldx;
    // an index outside its table (nn.Embedding raises IndexError there, ogbg-code/utils.py:27): nothing is read out of
    // bounds, the row becomes NaN and every output that depends on it is NaN — loud, not silently wrong
    if (t < 0 || t >= n_types || a < 0 || a >= n_attrs || dp < 0) {
      for (int c = lane; c < D; c += 32) o[c] = __int_as_float(0x7fc00000);
      if (VEC4 && ximg) {
        const int nck = (D + 63) >> 6;
        unsigned char* itile = ximg + (size_t)(v >> 7) * nck * (2 * 128 * tc::ROW_BYTES);
        for (int c = lane * 4; c < nck * 64; c += 128) {
          unsigned char* ihi = itile + (size_t)(c >> 6) * (2 * 128 * tc::ROW_BYTES) + tc::tile_off(v & 127, (c & 63) >> 3) + (c & 4) * 2;
          *reinterpret_cast<uint2*>(ihi) = make_uint2(0x7e007e00u, 0x7e007e00u);       // fp16 NaN
          *reinterpret_cast<uint2*>(ihi + 128 * tc::ROW_BYTES) = make_uint2(0u, 0u);
        }
      }
      continue;
    }
    const float* tr = T + (size_t)t * D;
    const float* ar = A + (size_t)a * D;
    const float* pr = P + (size_t)dp * D;
    if (VEC4) {
      // optional second copy of the row: fp16 hi / lo halves in the tcgen05 operand-image layout the sweep's first
      // projection bulk-copies (128-node tiles x 64-wide k chunks, 32 KB each; k padding zero-filled)
      const int nck = (D + 63) >> 6;
      unsigned char* itile = ximg ? ximg + (size_t)(v >> 7) * nck * (2 * 128 * tc::ROW_BYTES) : nullptr;
      if (itile)
        for (int c = D + lane * 4; c < nck * 64; c += 128) {
          unsigned char* ihi = itile + (size_t)(c >> 6) * (2 * 128 * tc::ROW_BYTES) + tc::tile_off(v & 127, (c & 63) >> 3) + (c & 4) * 2;
          *reinterpret_cast<uint2*>(ihi) = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(ihi + 128 * tc::ROW_BYTES) = make_uint2(0u, 0u);
        }
      for (int c = lane * 4; c < D; c += 128) {
        const float4 tv = __ldg(reinterpret_cast<const float4*>(tr + c));
        const float4 av = __ldg(reinterpret_cast<const float4*>(ar + c));
        const float4 pv = __ldg(reinterpret_cast<const float4*>(pr + c));
        float4 r;
        r.x = tv.x + av.x + pv.x; r.y = tv.y + av.y + pv.y; r.z = tv.z + av.z + pv.z; r.w = tv.w + av.w + pv.w;
        *reinterpret_cast<float4*>(o + c) = r;
        if (itile) {
          unsigned char* ihi = itile + (size_t)(c >> 6) * (2 * 128 * tc::ROW_BYTES) + tc::tile_off(v & 127, (c & 63) >> 3) + (c & 4) * 2;
          const __half2 h0 = __floats2half2_rn(r.x, r.y), h1 = __floats2half2_rn(r.z, r.w);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          const __half2 l0 = __floats2half2_rn(r.x - f0.x, r.y - f0.y), l1 = __floats2half2_rn(r.z - f1.x, r.w - f1.y);
          *reinterpret_cast<uint2*>(ihi) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          *reinterpret_cast<uint2*>(ihi + 128 * tc::ROW_BYTES) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
      }
    } else {
      for (int c = lane; c < D; c += 32) o[c] = __ldg(tr + c) + __ldg(ar + c) + __ldg(pr + c);
    }
  }
}

struct ReadoutArgs {
  int nblocks, pool, B;
  DagnnReadoutBlock blk[DAGNN_MAX_READOUT_BLOCKS];
  const int* pos[DAGNN_MAX_DIRS];
  const int* gptr;
  const int* summary;        // [2] != 0: the schedule is unusable (level table overflow / bad indices), nothing is read
};

// grid (B, nblocks, column chunks of 128), 256 threads: warp w scans nodes [v0 + 32w, v0 + 32w + 32) (+256 ...) of the
// graph — lane j tests node j of the chunk (filter, position lookup), the selected ones are then read row by row with
// every lane owning 4 columns (coalesced 128-byte segments, independent loads in flight); warps combine through smem.
__global__ void __launch_bounds__(256) k_readout(const __grid_constant__ ReadoutArgs a, float* __restrict__ out, int64_t ldo) {
  __shared__ float part[8][128];
  __shared__ int pcnt[8];
  const int g = blockIdx.x;
  const DagnnReadoutBlock& b = a.blk[blockIdx.y];
  const int c0 = blockIdx.z * 128;
  if (c0 >= b.width || a.summary[2] != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = a.gptr[g], v1 = a.gptr[g + 1];
  const int* pos = b.index_mode ? a.pos[b.dir] : nullptr;
  int vb = v0, ve = v1;
  if (b.filter == 2) vb = max(v0, v1 - 1);
  if (b.filter == 3) ve = min(v1, v0 + 1);
  const bool is_max = a.pool == 0;
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = is_max ? -INFINITY : 0.f;
  int cnt = 0;
  for (int base = vb + 32 * warp; base < ve; base += 256) {
    const int v = base + lane;
    bool sel = v < ve;
    if (sel && b.filter == 1) sel = b.filter_lvl[v] == 0;
    const long long row = sel ? (pos ? (long long)pos[v] : (long long)v) : 0;
    unsigned m = __ballot_sync(0xffffffffu, sel);
    cnt += __popc(m);
    while (m) {
      // four selected rows per round, their loads in flight together (the pooled value depends on all of them, a row at a
      // time is a chain of L2 round trips)
      float h[4][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const bool on = m != 0u;
        const int j0 = on ? __ffs(m) - 1 : 0;
        m &= m - 1;                                     // (0 stays 0)
        const long long r = __shfl_sync(0xffffffffu, row, j0);
        const float* srow = b.src + (size_t)r * b.ld + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          h[t][j] = (on && c0 + c < b.width) ? __ldcg(srow + c) : (is_max ? -INFINITY : 0.f);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = is_max ? ((h[t][j] > acc[j] || h[t][j] != h[t][j]) ? h[t][j] : acc[j]) : acc[j] + h[t][j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) part[warp][lane + 32 * j] = acc[j];
  if (lane == 0) pcnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    float r = part[0][c];
    int n = pcnt[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      r = is_max ? ((part[w][c] > r || part[w][c] != part[w][c]) ? part[w][c] : r) : r + part[w][c];
      n += pcnt[w];
    }
    if (n == 0) r = 0.f;
    else if (a.pool == 1) r = r / (float)n;
    if (c0 + c < b.width) out[(size_t)g * ldo + b.out_col + c0 + c] = r;
  }
}

}  // namespace dagnn

using namespace dagnn;

extern "C" size_t dagnn_operand_image_bytes(int64_t N, int32_t D) {
  if (N < 0 || D < 1) return 0;
  return (size_t)((N + 127) >> 7) * (size_t)((D + 63) >> 6) * (2 * 128 * tc::ROW_BYTES);
}

extern "C" int dagnn_embed_f32(const int64_t* x, const int64_t* depth, const float* type_tab, const float* attr_tab,
                               const float* depth_tab, int max_depth, int64_t n_types, int64_t n_attrs, int64_t N, int D, float* X,
                               int64_t ldx, void* x_image, void* stream_) {
  DAGNN_REQUIRE(x && depth && type_tab && attr_tab && depth_tab && X, "embed: null pointer");
  DAGNN_REQUIRE(N > 0 && N < (1ll << 31) && D > 0 && ldx >= D && max_depth >= 0 && n_types > 0 && n_attrs > 0, "embed: sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && ((((uintptr_t)type_tab | (uintptr_t)attr_tab | (uintptr_t)depth_tab | (uintptr_t)X) & 15) == 0);
  if (x_image && (!vec || ((uintptr_t)x_image & 1023) != 0))
    return set_err(DAGNN_E_INVALID, "embed: the operand image needs D %% 4 == 0, 16-byte aligned tables / X and a 1024-byte aligned image");
  const int blocks = (int)((N + 7) / 8 < 148 * 16 ? (N + 7) / 8 : 148 * 16);
  if (vec) k_embed<true><<<blocks, 256, 0, st>>>(x, depth, type_tab, attr_tab, depth_tab, max_depth, n_types, n_attrs, (int)N, D, X, ldx,
                                                  static_cast<unsigned char*>(x_image));
  else k_embed<false><<<blocks, 256, 0, st>>>(x, depth, type_tab, attr_tab, depth_tab, max_depth, n_types, n_attrs, (int)N, D, X, ldx,
                                         nullptr);
  return check_launch("k_embed");
}

extern "C" int dagnn_readout_f32(const DagnnSchedule* s, const DagnnReadoutBlock* blocks, int32_t nblocks, int32_t pool, float* out,
                                 int64_t ldo, void* stream_) {
  DAGNN_REQUIRE(s && blocks && out, "readout: null pointer");
  DAGNN_REQUIRE(nblocks > 0 && nblocks <= DAGNN_MAX_READOUT_BLOCKS, "readout: nblocks");
  DAGNN_REQUIRE(pool >= 0 && pool <= 2, "readout: pool");
  DAGNN_REQUIRE(s->B > 0 && s->gptr && s->summary, "readout: schedule has no graph pointers");
  ReadoutArgs a;
  a.nblocks = nblocks; a.pool = pool; a.B = (int)s->B;
  a.gptr = s->gptr;
  a.summary = s->summary;
  for (int d = 0; d < DAGNN_MAX_DIRS; ++d) a.pos[d] = d < s->dirs ? s->pos[d] : nullptr;
  int maxw = 0;
  for (int i = 0; i < nblocks; ++i) {
    a.blk[i] = blocks[i];
    DAGNN_REQUIRE(blocks[i].src && blocks[i].width > 0 && blocks[i].ld >= blocks[i].width, "readout: block");
    DAGNN_REQUIRE(blocks[i].filter >= 0 && blocks[i].filter <= 3, "readout: filter");
    DAGNN_REQUIRE(blocks[i].filter != 1 || blocks[i].filter_lvl, "readout: filter_lvl");
    DAGNN_REQUIRE(!blocks[i].index_mode || (blocks[i].dir >= 0 && blocks[i].dir < s->dirs), "readout: dir");
    maxw = blocks[i].width > maxw ? blocks[i].width : maxw;
  }
  dim3 grid((unsigned)s->B, (unsigned)nblocks, (unsigned)ceil_div(maxw, 128));
  k_readout<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(a, out, ldo);
  return check_launch("k_readout");
}
