// Backward of the level sweep, the readout and the node encoder: what torch.autograd derives from
// ogbg-code/model/dagnn.py:141-202 (AttnConv :362-373 + PyG softmax / scatter-add, nn.GRUCell :181, index_put :182, the pooled
// readout :184-202) and ogbg-code/utils.py:26-28, so that `loss.backward()` (main_pyg.py:55-65, dvae/train.py:255-264) runs
// on this package's modules.
//
// Per (direction d, layer i), layers from the last to the first (the input gradient of layer i + 1 lands in dH[d][i]):
//   recompute, all nodes at once (no level dependence: every state of the forward is known)
//     k_bwd_gather   m_v = sum_e alpha_e h_e and the softmax weights alpha_e of every in-edge (CSR order)
//     GEMM           Gi = inp W_ih^T, Gh = m W_hh^T                (fp16 x 3 split on tcgen05, gemm.cu)
//     k_bwd_gates    r, z, n, hn = W_hn m + b_hn
//   reverse level loop l = L-1 .. 0 (the only sequential part)
//     k_bwd_cell     dGi, dGh, dm (direct z * dh part) of the level's rows from dh
//     GEMM           dm += dGh W_hh                                (rows of the level)
//     k_bwd_attn     d alpha, d score; dh of the predecessors (atomicAdd into earlier levels), d wk, d edge coefficients
//   all nodes at once
//     GEMM           d inp = dGi W_ih  -> dH[d][i-1] (+=) or dX (scattered through perm)
//     GEMM           dW_ih = dGi^T inp, dW_hh = dGh^T m;  k_colsum: db_ih, db_hh
// The query part of attn_lin, attn_lin.bias and edge_encoder.bias shift every in-edge score of a node by the same constant:
// their gradient is exactly zero (autograd produces rounding noise there).
#include <vector>

#include "common.cuh"
#include "sync.cuh"

namespace dagnn {

int gemm_f16x3(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor, float* C, long long ldc,
               const float* bias, int M, int N, int K, int beta, cudaStream_t st);

__device__ __forceinline__ int level_of(const int* __restrict__ lvl_off, int L, int p) {     // largest l with lvl_off[l] <= p
  int lo = 0, hi = L;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (lvl_off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

struct AttnP {
  const int* rowptr; const int* col; const float* eattr; const int* perm; const int* lvl_off;
  const float* Hs; long long ldh;
  const float* attn_w; int Dq; const float* edge_w;      // raw parameters
  int H, nvid, L;
};
__device__ __forceinline__ float edge_score_terms(const AttnP& A, int e, int sp, float ca0, float ca1) {
  float sc = 0.f;
  if (A.eattr) sc = ca0 * A.eattr[2 * (size_t)e] + ca1 * A.eattr[2 * (size_t)e + 1];
  if (A.nvid > 0) sc += __ldg(A.attn_w + A.Dq + A.H + (A.perm[sp] % A.nvid));
  return sc;
}
// ca_c = sum_u wk[u] W_e[u, c]   (one warp)
__device__ __forceinline__ void edge_coeffs(const AttnP& A, int lane, float& ca0, float& ca1) {
  ca0 = 0.f; ca1 = 0.f;
  if (A.edge_w && A.eattr) {
    for (int u = lane; u < A.H; u += 32) {
      const float k = __ldg(A.attn_w + A.Dq + u);
      ca0 = fmaf(k, __ldg(A.edge_w + 2 * u), ca0);
      ca1 = fmaf(k, __ldg(A.edge_w + 2 * u + 1), ca1);
    }
    ca0 = warp_sum(ca0); ca1 = warp_sum(ca1);
  }
}

// one warp per position p >= first position of level 1: alpha[e] for its in-edges, M[p] = sum_e alpha_e h_e
__global__ void __launch_bounds__(256) k_bwd_gather(const AttnP A, int p_begin, int N, float* __restrict__ alpha, float* __restrict__ M) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float ca0, ca1;
  edge_coeffs(A, lane, ca0, ca1);
  const float* wk = A.attn_w + A.Dq;
  for (int p = p_begin + blockIdx.x * wpb + (threadIdx.x >> 5); p < N; p += gridDim.x * wpb) {
    const int lstart = A.lvl_off[level_of(A.lvl_off, A.L, p)];
    const int e0 = A.rowptr[p], e1 = A.rowptr[p + 1];
    float mx = -INFINITY;
    for (int e = e0; e < e1; ++e) {
      const int sp = A.col[e];
      float sc = edge_score_terms(A, e, sp, ca0, ca1);
      if (sp < lstart) {
        float dt = 0.f;
        const float* h = A.Hs + (size_t)sp * A.ldh;
        for (int u = lane; u < A.H; u += 32) dt = fmaf(__ldg(wk + u), h[u], dt);
        sc += warp_sum(dt);
      }
      if (lane == 0) alpha[e] = sc;
      mx = (sc > mx || sc != sc) ? sc : mx;
    }
    __syncwarp();
    float den = 0.f;
    for (int e = e0 + lane; e < e1; e += 32) den += expf(alpha[e] - mx);
    den = warp_sum(den);
    const float inv = 1.f / (den + 1e-16f);
    __syncwarp();
    for (int e = e0 + lane; e < e1; e += 32) alpha[e] = expf(alpha[e] - mx) * inv;
    __syncwarp();
    float* mrow = M + (size_t)p * A.ldh;
    for (int u = lane; u < A.H; u += 32) {
      float acc = 0.f;
      for (int e = e0; e < e1; ++e) {
        const int sp = A.col[e];
        if (sp < lstart) acc = fmaf(alpha[e], A.Hs[(size_t)sp * A.ldh + u], acc);
      }
      mrow[u] = acc;
    }
  }
}

// G1 = Gi (-> r | z | n in place), G2 = Gh (-> . | . | hn in place); biases added here
__global__ void __launch_bounds__(256) k_bwd_gates(float* __restrict__ G1, float* __restrict__ G2, long long ldg, const float* __restrict__ b_ih,
                                                   const float* __restrict__ b_hh, int N, int H) {
  const long long total = (long long)N * H;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / H), u = (int)(idx - (long long)p * H);
    float* g1 = G1 + (size_t)p * ldg;
    float* g2 = G2 + (size_t)p * ldg;
    const float r = 1.f / (1.f + expf(-(g1[u] + b_ih[u] + g2[u] + b_hh[u])));
    const float z = 1.f / (1.f + expf(-(g1[H + u] + b_ih[H + u] + g2[H + u] + b_hh[H + u])));
    const float hn = g2[2 * H + u] + b_hh[2 * H + u];
    const float n = tanhf(g1[2 * H + u] + b_ih[2 * H + u] + r * hn);
    g1[u] = r; g1[H + u] = z; g1[2 * H + u] = n; g2[2 * H + u] = hn;
  }
}

// rows [a, b) of a level: gate gradients from dh
__global__ void __launch_bounds__(256) k_bwd_cell(const float* __restrict__ G1, const float* __restrict__ G2, long long ldg,
                                                  const float* __restrict__ M, const float* __restrict__ dH, long long ldh,
                                                  float* __restrict__ dGi, float* __restrict__ dGh, float* __restrict__ dM, int a, int b, int H) {
  const long long total = (long long)(b - a) * H;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = a + (int)(idx / H), u = (int)(idx % H);
    const float* g1 = G1 + (size_t)p * ldg;
    const float r = g1[u], z = g1[H + u], n = g1[2 * H + u], hn = G2[(size_t)p * ldg + 2 * H + u];
    const float m = M[(size_t)p * ldh + u], dh = dH[(size_t)p * ldh + u];
    const float dn = dh * (1.f - z), dz = dh * (m - n);
    const float dpn = dn * (1.f - n * n);
    const float dar = dpn * hn * r * (1.f - r), daz = dz * z * (1.f - z);
    float* gi = dGi + (size_t)p * ldg;
    float* gh = dGh + (size_t)p * ldg;
    gi[u] = dar; gi[H + u] = daz; gi[2 * H + u] = dpn;
    gh[u] = dar; gh[H + u] = daz; gh[2 * H + u] = dpn * r;
    dM[(size_t)p * ldh + u] = dh * z;
  }
}

// one warp per row of the level (level > 0): softmax / score backward, gradient into the predecessors' dH rows
__global__ void __launch_bounds__(256) k_bwd_attn(const AttnP A, const float* __restrict__ alpha, const float* __restrict__ dM, float* __restrict__ dH,
                                                  int a, int b, float* __restrict__ dwk /*[H]*/, float* __restrict__ dca /*[2]*/,
                                                  float* __restrict__ dvid /*[nvid]*/) {
  extern __shared__ float dwk_s[];                 // [H]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int u = threadIdx.x; u < A.H; u += blockDim.x) dwk_s[u] = 0.f;
  __syncthreads();
  const float* wk = A.attn_w + A.Dq;
  const int lstart = a;
  float c0 = 0.f, c1 = 0.f;
  for (int p = a + blockIdx.x * wpb + (threadIdx.x >> 5); p < b; p += gridDim.x * wpb) {
    const int e0 = A.rowptr[p], e1 = A.rowptr[p + 1];
    const float* dm = dM + (size_t)p * A.ldh;
    // S = sum_e alpha_e d alpha_e,  d alpha_e = dm . h_e (0 for a predecessor that is not in an earlier level)
    float S = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int sp = A.col[e];
      if (sp < lstart) {
        float dt = 0.f;
        const float* h = A.Hs + (size_t)sp * A.ldh;
        for (int u = lane; u < A.H; u += 32) dt = fmaf(dm[u], h[u], dt);
        S = fmaf(alpha[e], warp_sum(dt), S);
      }
    }
    for (int e = e0; e < e1; ++e) {
      const int sp = A.col[e];
      const bool valid = sp < lstart;
      const float al = alpha[e];
      float da = 0.f;
      const float* h = A.Hs + (size_t)sp * A.ldh;
      if (valid) {
        float dt = 0.f;
        for (int u = lane; u < A.H; u += 32) dt = fmaf(dm[u], h[u], dt);
        da = warp_sum(dt);
      }
      const float ds = al * (da - S);
      if (valid) {
        float* dh = dH + (size_t)sp * A.ldh;
        for (int u = lane; u < A.H; u += 32) {
          atomicAdd(dh + u, fmaf(al, dm[u], ds * __ldg(wk + u)));
          atomicAdd(dwk_s + u, ds * h[u]);
        }
      }
      if (lane == 0) {
        if (A.eattr) { c0 = fmaf(ds, A.eattr[2 * (size_t)e], c0); c1 = fmaf(ds, A.eattr[2 * (size_t)e + 1], c1); }
        if (A.nvid > 0) atomicAdd(dvid + (A.perm[sp] % A.nvid), ds);
      }
    }
  }
  if (lane == 0 && A.eattr) { atomicAdd(dca, c0); atomicAdd(dca + 1, c1); }
  __syncthreads();
  for (int u = threadIdx.x; u < A.H; u += blockDim.x)
    if (dwk_s[u] != 0.f) atomicAdd(dwk + u, dwk_s[u]);
}

// Short levels (the tail of the level chain: a handful of rows each): the three steps of a level in ONE launch, one CTA per
// row — gate gradients, dm = z dh + dGh W_hh as a matrix-vector product against W_hh streamed from L2 (coalesced: thread =
// hidden unit), softmax / score backward over the row's in-edges by the CTA's 8 warps. dGi / dGh rows still go to global
// memory (the weight-gradient GEMMs read them at the end).
constexpr int kFusedMaxH = 768;                    // 12 H + 16 floats of shared memory per CTA stay under the 48 KB default
__global__ void __launch_bounds__(256) k_bwd_level_fused(const AttnP A, const float* __restrict__ G1, const float* __restrict__ G2, long long ldg,
                                                         const float* __restrict__ M, const float* __restrict__ alpha, const float* __restrict__ W_hh,
                                                         float* __restrict__ dH, float* __restrict__ dGi, float* __restrict__ dGh, int a,
                                                         float* __restrict__ dwk, float* __restrict__ dca, float* __restrict__ dvid) {
  extern __shared__ float sm[];                    // [3H] dGh row | [H] dm row | [8] partial S | [1] S | pad | [8][H] GEMV partials
  const int H = A.H;
  float* sgh = sm;
  float* sdm = sm + 3 * H;
  float* spart = sdm + H;
  const int p = a + blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* g1 = G1 + (size_t)p * ldg;
  for (int u = tid; u < H; u += 256) {
    const float r = g1[u], z = g1[H + u], n = g1[2 * H + u], hn = G2[(size_t)p * ldg + 2 * H + u];
    const float m = M[(size_t)p * A.ldh + u], dh = dH[(size_t)p * A.ldh + u];
    const float dn = dh * (1.f - z), dz = dh * (m - n);
    const float dpn = dn * (1.f - n * n);
    const float dar = dpn * hn * r * (1.f - r), daz = dz * z * (1.f - z);
    float* gi = dGi + (size_t)p * ldg;
    float* gh = dGh + (size_t)p * ldg;
    gi[u] = dar; gi[H + u] = daz; gi[2 * H + u] = dpn;
    gh[u] = dar; gh[H + u] = daz; gh[2 * H + u] = dpn * r;
    sgh[u] = dar; sgh[H + u] = daz; sgh[2 * H + u] = dpn * r;
    sdm[u] = dh * z;
  }
  __syncthreads();
  // dm[u] += sum_k dGh[k] W_hh[k, u]: warp w takes the rows k = w, w + 8, ... of W_hh, a lane 4 consecutive units per 128-unit
  // slab (coalesced 16-byte loads), 8 rows in flight per lane; the 8 partial rows are summed through shared memory
  float* spw = sm + 4 * H + 16;                    // [8][H] partial sums of the warps
  for (int ub = 0; ub < H; ub += 128) {
    const int u = ub + 4 * lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u + 4 <= H && (H & 3) == 0) {
      int k = warp;
      for (; k + 120 < 3 * H; k += 128) {              // 16 rows of W_hh in flight per lane: the product is bound by L2 latency
        float4 wv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(W_hh + (size_t)(k + 8 * j) * H + u));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float g = sgh[k + 8 * j];
          acc.x = fmaf(g, wv[j].x, acc.x); acc.y = fmaf(g, wv[j].y, acc.y); acc.z = fmaf(g, wv[j].z, acc.z); acc.w = fmaf(g, wv[j].w, acc.w);
        }
      }
      for (; k < 3 * H; k += 8) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(W_hh + (size_t)k * H + u));
        const float g = sgh[k];
        acc.x = fmaf(g, wv.x, acc.x); acc.y = fmaf(g, wv.y, acc.y); acc.z = fmaf(g, wv.z, acc.z); acc.w = fmaf(g, wv.w, acc.w);
      }
    } else {
      for (int k = warp; k < 3 * H; k += 8) {
        const float g = sgh[k];
        if (u < H) acc.x = fmaf(g, __ldg(W_hh + (size_t)k * H + u), acc.x);
        if (u + 1 < H) acc.y = fmaf(g, __ldg(W_hh + (size_t)k * H + u + 1), acc.y);
        if (u + 2 < H) acc.z = fmaf(g, __ldg(W_hh + (size_t)k * H + u + 2), acc.z);
        if (u + 3 < H) acc.w = fmaf(g, __ldg(W_hh + (size_t)k * H + u + 3), acc.w);
      }
    }
    if (u < H) spw[warp * H + u] = acc.x;
    if (u + 1 < H) spw[warp * H + u + 1] = acc.y;
    if (u + 2 < H) spw[warp * H + u + 2] = acc.z;
    if (u + 3 < H) spw[warp * H + u + 3] = acc.w;
  }
  __syncthreads();
  for (int u = tid; u < H; u += 256) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += spw[w8 * H + u];
    sdm[u] += t;
  }
  __syncthreads();
  // ---- attention backward: edges of the row dealt to the 8 warps
  const int e0 = A.rowptr[p], e1 = A.rowptr[p + 1];
  const int lstart = a;
  const float* wk = A.attn_w + A.Dq;
  float Sp = 0.f;
  for (int e = e0 + warp; e < e1; e += 8) {
    const int sp = A.col[e];
    if (sp < lstart) {
      float dt = 0.f;
      const float* h = A.Hs + (size_t)sp * A.ldh;
      for (int u = lane; u < H; u += 32) dt = fmaf(sdm[u], h[u], dt);
      Sp = fmaf(alpha[e], warp_sum(dt), Sp);
    }
  }
  if (lane == 0) spart[warp] = Sp;
  __syncthreads();
  if (tid == 0) {
    float S = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) S += spart[w8];
    spart[8] = S;
  }
  __syncthreads();
  const float S = spart[8];
  float c0 = 0.f, c1 = 0.f;
  for (int e = e0 + warp; e < e1; e += 8) {
    const int sp = A.col[e];
    const bool valid = sp < lstart;
    const float al = alpha[e];
    float da = 0.f;
    const float* h = A.Hs + (size_t)sp * A.ldh;
    if (valid) {
      float dt = 0.f;
      for (int u = lane; u < H; u += 32) dt = fmaf(sdm[u], h[u], dt);
      da = warp_sum(dt);
    }
    const float ds = al * (da - S);
    if (valid) {
      float* dh = dH + (size_t)sp * A.ldh;
      for (int u = lane; u < H; u += 32) {
        atomicAdd(dh + u, fmaf(al, sdm[u], ds * __ldg(wk + u)));
        atomicAdd(dwk + u, ds * h[u]);
      }
    }
    if (lane == 0) {
      if (A.eattr) { c0 = fmaf(ds, A.eattr[2 * (size_t)e], c0); c1 = fmaf(ds, A.eattr[2 * (size_t)e + 1], c1); }
      if (A.nvid > 0) atomicAdd(dvid + (A.perm[sp] % A.nvid), ds);
    }
  }
  if (lane == 0 && A.eattr && (c0 != 0.f || c1 != 0.f)) { atomicAdd(dca, c0); atomicAdd(dca + 1, c1); }
}

// dst[perm[p], :] += src[p, :]   (position order -> node order, one writer per element)
__global__ void __launch_bounds__(256) k_scatter_rows_add(const int* __restrict__ perm, const float* __restrict__ src, long long lds,
                                                          float* __restrict__ dst, long long ldd, int N, int W) {
  const long long total = (long long)N * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / W), c = (int)(idx % W);
    dst[(size_t)perm[p] * ldd + c] += src[(size_t)p * lds + c];
  }
}
// dst[p, :] = src[perm[p], :]
__global__ void __launch_bounds__(256) k_gather_rows(const int* __restrict__ perm, const float* __restrict__ src, long long lds,
                                                     float* __restrict__ dst, long long ldd, int N, int W) {
  const long long total = (long long)N * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / W), c = (int)(idx % W);
    dst[(size_t)p * ldd + c] = src[(size_t)perm[p] * lds + c];
  }
}
// dst[p, :] += src[p, :]
__global__ void __launch_bounds__(256) k_rows_add(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, int N, int W) {
  const long long total = (long long)N * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / W), c = (int)(idx % W);
    dst[(size_t)p * ldd + c] += src[(size_t)p * lds + c];
  }
}
// out[c] = sum_p X[p, c]: block (x = 32-column group, y = row slice); out zeroed by the caller
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, long long ld, int N, int W, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < W)
    for (int p = blockIdx.y * 8 + w; p < N; p += gridDim.y * 8) s += X[(size_t)p * ld + c];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}
// attention parameter gradients from the accumulators: d attn_w[Dq + u] = dwk[u] + sum_c W_e[u, c] dca[c]; d W_e[u, c] = wk[u] dca[c];
// the vertex-id part; zeros elsewhere (query part: no gradient)
__global__ void __launch_bounds__(256) k_bwd_attn_params(const float* __restrict__ attn_w, int Dq, int H, int nvid, const float* __restrict__ edge_w,
                                                         const float* __restrict__ dwk, const float* __restrict__ dca, const float* __restrict__ dvid,
                                                         float* __restrict__ d_attn_w, float* __restrict__ d_edge_w) {
  const int total = Dq + H + nvid;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    float g = 0.f;
    if (j >= Dq && j < Dq + H) {
      const int u = j - Dq;
      g = dwk[u];
      if (edge_w) {
        g += edge_w[2 * u] * dca[0] + edge_w[2 * u + 1] * dca[1];
        d_edge_w[2 * u] = attn_w[Dq + u] * dca[0];
        d_edge_w[2 * u + 1] = attn_w[Dq + u] * dca[1];
      }
    } else if (j >= Dq + H) {
      g = dvid[j - Dq - H];
    }
    d_attn_w[j] = g;
  }
}

// ---- readout backward: mirrors k_readout (embed_readout.cu): grid (B, nblocks), the gradient of graph g's pooled row goes to the
// selected nodes (max: the rows that attain the maximum; mean: 1/count each; add: each)
struct ReadoutBwdArgs {
  int nblocks, pool;
  DagnnReadoutBlock blk[DAGNN_MAX_READOUT_BLOCKS];      // src = the GRADIENT buffer of the block's source, filter etc. as in the forward
  const float* fwd_src[DAGNN_MAX_READOUT_BLOCKS];       // the forward source (max pool compares against the pooled value)
  const int* pos[DAGNN_MAX_DIRS];
  const int* gptr;
};
__global__ void __launch_bounds__(256) k_readout_bwd(const __grid_constant__ ReadoutBwdArgs a, const float* __restrict__ out, const float* __restrict__ dout,
                                                     long long ldo) {
  __shared__ int cnt_s;
  const int g = blockIdx.x;
  const DagnnReadoutBlock& b = a.blk[blockIdx.y];
  const float* fsrc = a.fwd_src[blockIdx.y];
  float* dsrc = const_cast<float*>(b.src);
  const int v0 = a.gptr[g], v1 = a.gptr[g + 1];
  const int* pos = b.index_mode ? a.pos[b.dir] : nullptr;
  int vb = v0, ve = v1;
  if (b.filter == 2) vb = max(v0, v1 - 1);
  if (b.filter == 3) ve = min(v1, v0 + 1);
  if (threadIdx.x == 0) cnt_s = 0;
  __syncthreads();
  if (a.pool == 1) {
    int c = 0;
    for (int v = vb + threadIdx.x; v < ve; v += blockDim.x) c += (b.filter != 1 || b.filter_lvl[v] == 0) ? 1 : 0;
    if (c) atomicAdd(&cnt_s, c);
    __syncthreads();
  }
  const float scale = (a.pool == 1 && cnt_s > 0) ? 1.f / (float)cnt_s : 1.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = vb + warp; v < ve; v += 8) {
    if (b.filter == 1 && b.filter_lvl[v] != 0) continue;
    const long long row = pos ? (long long)pos[v] : (long long)v;
    for (int c = lane; c < b.width; c += 32) {
      const float go = dout[(size_t)g * ldo + b.out_col + c];
      float gr = go * scale;
      if (a.pool == 0) gr = (fsrc[(size_t)row * b.ld + c] == out[(size_t)g * ldo + b.out_col + c]) ? go : 0.f;
      if (gr != 0.f) atomicAdd(dsrc + (size_t)row * b.ld + c, gr);
    }
  }
}

// ---- node encoder backward: dT[x0] += dX, dA[x1] += dX, dP[min(depth, max_depth)] += dX
__global__ void __launch_bounds__(256) k_embed_bwd(const int64_t* __restrict__ x, const int64_t* __restrict__ depth, int max_depth, long long n_types,
                                                   long long n_attrs, int N, int D, const float* __restrict__ dX, long long ldx,
                                                   float* __restrict__ dT, float* __restrict__ dA, float* __restrict__ dP) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int v = blockIdx.x * wpb + (threadIdx.x >> 5); v < N; v += gridDim.x * wpb) {
    const long long t = x[2 * (size_t)v], at = x[2 * (size_t)v + 1];
    long long dp = depth[v];
    dp = dp > max_depth ? max_depth : dp;
    if (t < 0 || t >= n_types || at < 0 || at >= n_attrs || dp < 0) continue;
    const float* g = dX + (size_t)v * ldx;
    for (int c = lane; c < D; c += 32) {
      const float gv = g[c];
      atomicAdd(dT + (size_t)t * D + c, gv);
      atomicAdd(dA + (size_t)at * D + c, gv);
      atomicAdd(dP + (size_t)dp * D + c, gv);
    }
  }
}

static size_t balign(size_t x) { return (x + 255) / 256 * 256; }

struct BwdWs {
  float *M, *G1, *G2, *dGi, *dGh, *dM, *alpha, *Xpos, *dinp, *acc;
  long long ldg;
  size_t bytes;
};
static BwdWs carve_bwd(void* ws, int Din, int H, int64_t N, int64_t E, int nvid) {
  BwdWs w;
  const size_t ldh = (size_t)round_up(H, 4);
  w.ldg = round_up(3 * H, 4);
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = ws ? reinterpret_cast<float*>(static_cast<char*>(ws) + off) : nullptr;
    off += balign(nfloat * sizeof(float));
    return p;
  };
  w.M = take((size_t)N * ldh);
  w.G1 = take((size_t)N * w.ldg);
  w.G2 = take((size_t)N * w.ldg);
  w.dGi = take((size_t)N * w.ldg);
  w.dGh = take((size_t)N * w.ldg);
  w.dM = take((size_t)N * ldh);
  w.alpha = take((size_t)(E > 0 ? E : 1));
  w.Xpos = take((size_t)N * round_up(Din, 4));
  w.dinp = take((size_t)N * round_up(Din > H ? Din : H, 4));
  w.acc = take((size_t)round_up(H, 4) + 4 + round_up(nvid > 0 ? nvid : 1, 4));
  w.bytes = off;
  return w;
}

}  // namespace dagnn

using namespace dagnn;

// one scratch set per direction: the two directions are independent and run concurrently on two streams
extern "C" size_t dagnn_sweep_backward_workspace_bytes(int32_t Din, int32_t H, int32_t nvid, int64_t N, int64_t E) {
  if (Din < 1 || H < 1 || N < 0 || E < 0 || nvid < 0) return 0;
  return DAGNN_MAX_DIRS * carve_bwd(nullptr, Din, H, N, E, nvid).bytes;
}

namespace {
struct BwdStreams {                 // per device: the second direction's stream and the fork / join events
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
BwdStreams g_bwd_streams[dagnn::kMaxDevices];
dagnn::PerDeviceOnce g_bwd_once;
}  // namespace

static int grid_for(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

extern "C" int dagnn_sweep_backward_f32(const DagnnSweepBwdArgs* A, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(A && A->sched, "sweep backward: null args");
  const DagnnSchedule* S = A->sched;
  const int dirs = S->dirs, layers = A->num_layers, H = A->H, Din = A->Din, N = (int)S->N, L = A->num_levels;
  DAGNN_REQUIRE(dirs >= 1 && dirs <= DAGNN_MAX_DIRS && layers >= 1 && layers <= DAGNN_MAX_LAYERS, "sweep backward: dirs / layers");
  DAGNN_REQUIRE(A->X && A->ldx >= Din && H >= 1 && Din >= 1 && A->ldh >= H && L >= 1 && L <= S->max_levels, "sweep backward: sizes");
  DAGNN_REQUIRE(A->workspace && ((uintptr_t)A->workspace & 255) == 0, "sweep backward: workspace must be 256-byte aligned");
  const size_t ws_one = carve_bwd(nullptr, Din, H, N, S->E, A->nvid).bytes;
  if (A->workspace_bytes < dirs * ws_one) return set_err(DAGNN_E_WORKSPACE, "sweep backward: workspace too small");
  const long long ldh = A->ldh;
  const int ldx_pos = round_up(Din, 4), ld_dinp = round_up(Din > H ? Din : H, 4);
  // direction 1 runs on a side stream (forked from / joined into the caller's): the per-level launches of the two directions
  // interleave on the device instead of queueing behind each other. dX is the only shared output: its two scatter-adds are
  // ordered by the join (direction 1 adds its part on the caller's stream after the join).
  cudaStream_t main_st = st;
  int dev = 0;
  if (int rc = per_device_once(g_bwd_once, &dev, [&](int dv) {
        DAGNN_CUDA_OK(cudaStreamCreateWithFlags(&g_bwd_streams[dv].side, cudaStreamNonBlocking));
        DAGNN_CUDA_OK(cudaEventCreateWithFlags(&g_bwd_streams[dv].fork, cudaEventDisableTiming));
        DAGNN_CUDA_OK(cudaEventCreateWithFlags(&g_bwd_streams[dv].join, cudaEventDisableTiming));
        return (int)DAGNN_OK;
      }))
    return rc;
  const BwdStreams& bs = g_bwd_streams[dev];
  if (dirs == 2) {
    DAGNN_CUDA_OK(cudaEventRecord(bs.fork, main_st));
    DAGNN_CUDA_OK(cudaStreamWaitEvent(bs.side, bs.fork, 0));
  }
  float* dinp1_deferred = nullptr;
  for (int d = 0; d < dirs; ++d) {
    st = d == 0 ? main_st : bs.side;
    BwdWs w = carve_bwd(static_cast<char*>(A->workspace) + (size_t)d * ws_one, Din, H, N, S->E, A->nvid);
    const long long ldg = w.ldg;
    DAGNN_REQUIRE(A->lvl_off_host[d], "sweep backward: host level offsets");
    const int32_t* lo = A->lvl_off_host[d];
    for (int i = layers - 1; i >= 0; --i) {
      const DagnnCellParams& pr = A->params[d][i];
      const DagnnCellGrads& gr = A->grads[d][i];
      DAGNN_REQUIRE(A->Hs[d][i] && A->dHs[d][i] && pr.weight_ih && pr.weight_hh && pr.bias_ih && pr.bias_hh && pr.attn_w, "sweep backward: pointers");
      DAGNN_REQUIRE(gr.weight_ih && gr.weight_hh && gr.bias_ih && gr.bias_hh && gr.attn_w, "sweep backward: gradient pointers");
      DAGNN_REQUIRE(!(A->use_edge_attr && pr.edge_w) || gr.edge_w, "sweep backward: edge_w gradient");
      const int Di = i == 0 ? Din : H;
      const float* Hcur = A->Hs[d][i];
      float* dH = A->dHs[d][i];
      AttnP ap;
      ap.rowptr = S->rowptr[d]; ap.col = S->col[d]; ap.eattr = (A->use_edge_attr && pr.edge_w) ? S->eattr[d] : nullptr; ap.perm = S->perm[d];
      ap.lvl_off = S->lvl_off[d]; ap.Hs = Hcur; ap.ldh = ldh; ap.attn_w = pr.attn_w; ap.Dq = pr.Dq; ap.edge_w = pr.edge_w; ap.H = H;
      ap.nvid = A->nvid; ap.L = L;
      // ---- layer input in position order
      const float* inp; long long ldi;
      if (i == 0) {
        k_gather_rows<<<grid_for((long long)N * Din, 256), 256, 0, st>>>(S->perm[d], A->X, A->ldx, w.Xpos, ldx_pos, N, Din);
        if (int rc = check_launch("k_gather_rows")) return rc;
        inp = w.Xpos; ldi = ldx_pos;
      } else { inp = A->Hs[d][i - 1]; ldi = ldh; }
      // ---- recompute aggregates, gates
      DAGNN_CUDA_OK(cudaMemsetAsync(w.M, 0, (size_t)N * ldh * sizeof(float), st));
      DAGNN_CUDA_OK(cudaMemsetAsync(w.acc, 0, (size_t)(round_up(H, 4) + 4 + round_up(A->nvid > 0 ? A->nvid : 1, 4)) * sizeof(float), st));
      const int p1 = L > 1 ? lo[1] : N;
      if (p1 < N && S->E > 0) {
        k_bwd_gather<<<grid_for(N - p1, 8), 256, 0, st>>>(ap, p1, N, w.alpha, w.M);
        if (int rc = check_launch("k_bwd_gather")) return rc;
      }
      if (int rc = gemm_f16x3(inp, ldi, 1, pr.weight_ih, Di, 1, w.G1, ldg, nullptr, N, 3 * H, Di, 0, st)) return rc;
      if (int rc = gemm_f16x3(w.M, ldh, 1, pr.weight_hh, H, 1, w.G2, ldg, nullptr, N, 3 * H, H, 0, st)) return rc;
      k_bwd_gates<<<grid_for((long long)N * H, 256), 256, 0, st>>>(w.G1, w.G2, ldg, pr.bias_ih, pr.bias_hh, N, H);
      if (int rc = check_launch("k_bwd_gates")) return rc;
      // ---- reverse level loop
      float* dwk = w.acc; float* dca = w.acc + round_up(H, 4); float* dvid = dca + 4;
      for (int l = L - 1; l >= 0; --l) {
        const int a = lo[l], b = lo[l + 1];
        if (b <= a) continue;
        if (l > 0 && b - a <= 148 && H <= kFusedMaxH && S->E > 0) {             // the tail of the level chain: one launch per level
          k_bwd_level_fused<<<b - a, 256, (size_t)(12 * H + 16) * sizeof(float), st>>>(ap, w.G1, w.G2, ldg, w.M, w.alpha, pr.weight_hh, dH, w.dGi,
                                                                                       w.dGh, a, dwk, dca, dvid);
          if (int rc = check_launch("k_bwd_level_fused")) return rc;
          continue;
        }
        k_bwd_cell<<<grid_for((long long)(b - a) * H, 256), 256, 0, st>>>(w.G1, w.G2, ldg, w.M, dH, ldh, w.dGi, w.dGh, w.dM, a, b, H);
        if (int rc = check_launch("k_bwd_cell")) return rc;
        if (l == 0) break;
        // dM[rows] += dGh[rows] W_hh   (contraction over the 3H gate rows: W_hh is the [K, N] operand)
        if (int rc = gemm_f16x3(w.dGh + (size_t)a * ldg, ldg, 1, pr.weight_hh, H, 0, w.dM + (size_t)a * ldh, ldh, nullptr, b - a, H, 3 * H, 1, st))
          return rc;
        if (S->E > 0) {
          k_bwd_attn<<<grid_for(b - a, 8), 256, (size_t)H * sizeof(float), st>>>(ap, w.alpha, w.dM, dH, a, b, dwk, dca, dvid);
          if (int rc = check_launch("k_bwd_attn")) return rc;
        }
      }
      // ---- input gradient: d inp = dGi W_ih
      if (i > 0 || A->dX) {
        if (int rc = gemm_f16x3(w.dGi, ldg, 1, pr.weight_ih, Di, 0, w.dinp, ld_dinp, nullptr, N, Di, 3 * H, 0, st)) return rc;
        if (i > 0) {
          k_rows_add<<<grid_for((long long)N * H, 256), 256, 0, st>>>(w.dinp, ld_dinp, A->dHs[d][i - 1], ldh, N, H);
          if (int rc = check_launch("k_rows_add")) return rc;
        } else if (d == 0) {
          k_scatter_rows_add<<<grid_for((long long)N * Din, 256), 256, 0, st>>>(S->perm[d], w.dinp, ld_dinp, A->dX, A->lddx, N, Din);
          if (int rc = check_launch("k_scatter_rows_add")) return rc;
        } else {
          dinp1_deferred = w.dinp;               // added on the caller's stream behind the join
        }
      }
      // ---- parameter gradients
      if (int rc = gemm_f16x3(w.dGi, ldg, 0, inp, ldi, 0, gr.weight_ih, Di, nullptr, 3 * H, Di, N, 0, st)) return rc;
      if (int rc = gemm_f16x3(w.dGh, ldg, 0, w.M, ldh, 0, gr.weight_hh, H, nullptr, 3 * H, H, N, 0, st)) return rc;
      DAGNN_CUDA_OK(cudaMemsetAsync(gr.bias_ih, 0, (size_t)3 * H * sizeof(float), st));
      DAGNN_CUDA_OK(cudaMemsetAsync(gr.bias_hh, 0, (size_t)3 * H * sizeof(float), st));
      dim3 cg((unsigned)ceil_div(3 * H, 32), (unsigned)(N >= 4096 ? 64 : (N >= 256 ? 8 : 1)));
      k_colsum<<<cg, 256, 0, st>>>(w.dGi, ldg, N, 3 * H, gr.bias_ih);
      if (int rc = check_launch("k_colsum")) return rc;
      k_colsum<<<cg, 256, 0, st>>>(w.dGh, ldg, N, 3 * H, gr.bias_hh);
      if (int rc = check_launch("k_colsum")) return rc;
      k_bwd_attn_params<<<ceil_div(pr.Dq + H + A->nvid, 256), 256, 0, st>>>(pr.attn_w, pr.Dq, H, A->nvid, ap.eattr ? pr.edge_w : nullptr, dwk, dca,
                                                                              dvid, gr.attn_w, gr.edge_w);
      if (int rc = check_launch("k_bwd_attn_params")) return rc;
    }
  }
  if (dirs == 2) {
    DAGNN_CUDA_OK(cudaEventRecord(bs.join, bs.side));
    DAGNN_CUDA_OK(cudaStreamWaitEvent(main_st, bs.join, 0));
    if (dinp1_deferred) {
      k_scatter_rows_add<<<grid_for((long long)N * Din, 256), 256, 0, main_st>>>(S->perm[1], dinp1_deferred, ld_dinp, A->dX, A->lddx, N, Din);
      if (int rc = check_launch("k_scatter_rows_add")) return rc;
    }
  }
  return DAGNN_OK;
}

extern "C" int dagnn_readout_backward_f32(const DagnnSchedule* s, const DagnnReadoutBlock* grad_blocks, const float* const* fwd_src, int32_t nblocks,
                                          int32_t pool, const float* out, const float* dout, int64_t ldo, void* stream_) {
  DAGNN_REQUIRE(s && grad_blocks && fwd_src && out && dout, "readout backward: null pointer");
  DAGNN_REQUIRE(nblocks > 0 && nblocks <= DAGNN_MAX_READOUT_BLOCKS && pool >= 0 && pool <= 2, "readout backward: nblocks / pool");
  DAGNN_REQUIRE(s->B > 0 && s->gptr, "readout backward: schedule has no graph pointers");
  ReadoutBwdArgs a;
  a.nblocks = nblocks; a.pool = pool; a.gptr = s->gptr;
  for (int d = 0; d < DAGNN_MAX_DIRS; ++d) a.pos[d] = d < s->dirs ? s->pos[d] : nullptr;
  for (int i = 0; i < nblocks; ++i) {
    a.blk[i] = grad_blocks[i];
    a.fwd_src[i] = fwd_src[i];
    DAGNN_REQUIRE(grad_blocks[i].src && fwd_src[i] && grad_blocks[i].width > 0 && grad_blocks[i].ld >= grad_blocks[i].width, "readout backward: block");
    DAGNN_REQUIRE(grad_blocks[i].filter >= 0 && grad_blocks[i].filter <= 3 && (grad_blocks[i].filter != 1 || grad_blocks[i].filter_lvl), "readout backward: filter");
    DAGNN_REQUIRE(!grad_blocks[i].index_mode || (grad_blocks[i].dir >= 0 && grad_blocks[i].dir < s->dirs), "readout backward: dir");
  }
  k_readout_bwd<<<dim3((unsigned)s->B, (unsigned)nblocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a, out, dout, ldo);
  return check_launch("k_readout_bwd");
}

extern "C" int dagnn_embed_backward_f32(const int64_t* x, const int64_t* depth, int max_depth, int64_t n_types, int64_t n_attrs, int64_t N, int D,
                                        const float* dX, int64_t ldx, float* d_type_tab, float* d_attr_tab, float* d_depth_tab, void* stream_) {
  DAGNN_REQUIRE(x && depth && dX && d_type_tab && d_attr_tab && d_depth_tab, "embed backward: null pointer");
  DAGNN_REQUIRE(N > 0 && N < (1ll << 31) && D > 0 && ldx >= D && max_depth >= 0, "embed backward: sizes");
  const int blocks = (int)((N + 7) / 8 < 148 * 16 ? (N + 7) / 8 : 148 * 16);
  k_embed_bwd<<<blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, depth, max_depth, n_types, n_attrs, (int)N, D, dX, ldx, d_type_tab, d_attr_tab,
                                                                     d_depth_tab);
  return check_launch("k_embed_bwd");
}
