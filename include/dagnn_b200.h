/*
 * dagnn_b200.h — C ABI of libdagnn_sm100.so: the DAGNN layer-wise (topological-level) forward on B200.
 *
 * The reference (vthost/DAGNN @ b065cd5) has NO FFI / plugin ABI for this path: its boundary is the Python
 * nn.Module API. This header is the drop-in boundary a maintainer would bind instead of the
 * library calls listed in SURVEY.md §2.1 (K1-K13); each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + sizes + leading dimensions; no torch / C++ types cross the boundary;
 *   - the caller owns every buffer (nothing is allocated inside; scratch comes in through *workspace*
 *     arguments sized by the matching *_bytes() query);
 *   - every call is asynchronous on the given CUDA stream (cudaStream_t passed as void*), no hidden sync;
 *   - return value: 0 on success, negative DAGNN_E_* otherwise; dagnn_last_error() gives a thread-local
 *     message. Nothing throws across the ABI;
 *   - no global mutable state except per-process cached function attributes and a launch counter;
 *     safe to call from several host threads on different streams / devices.
 *   - fp32 everywhere ("f32" suffix); integer schedule arrays are int32 on the device (inputs int64 as PyG
 *     provides them).
 */
#ifndef DAGNN_B200_H_
#define DAGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAGNN_ABI_VERSION 10
#define DAGNN_MAX_LAYERS 8          /* stacked GRU layers per direction                         */
#define DAGNN_MAX_DIRS 2            /* 0 = forward sweep, 1 = sweep on the reversed DAG          */
#define DAGNN_K_CHUNK 64            /* K granularity of the packed weight images (one swizzle row of fp16) */
#define DAGNN_MAX_READOUT_BLOCKS 20 /* column blocks of one readout call                         */

enum {
  DAGNN_OK = 0,
  DAGNN_E_INVALID = -1,   /* bad argument (null pointer, size, alignment)                    */
  DAGNN_E_CUDA = -2,      /* a CUDA runtime call / launch failed                             */
  DAGNN_E_WORKSPACE = -3, /* workspace too small                                             */
  DAGNN_E_UNSUPPORTED = -4
};

int dagnn_abi_version(void);
const char* dagnn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
int64_t dagnn_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Node encoder:  X[v,:] = T[x[v,0],:] + A[x[v,1],:] + P[min(depth[v], max_depth),:]
 * replaces ASTNodeEncoder.forward, ogbg-code/utils.py:26-28 (called at ogbg-code/model/dagnn.py:139).
 * x int64 [N,2] row-major, depth int64 [N]; tables fp32 row-major with leading dimension D; X fp32 [N, ldx].
 * Unlike the reference it does not clamp `depth` in place. n_types / n_attrs = rows of the two tables: a row whose index
 * is outside its table (nn.Embedding raises IndexError there) or whose depth is negative becomes NaN — no out-of-bounds
 * read, and every output depending on it is NaN.
 * x_image (optional, NULL = off): a second copy of X as fp16 hi / lo halves in the tcgen05 operand-image layout
 * (128-node tiles x 64-wide k chunks, dagnn_operand_image_bytes(N, D) bytes, 1024-byte aligned; needs D % 4 == 0) that
 * dagnn_sweep_forward_f32 can bulk-copy for its first projection (DagnnSweepArgs.X_image).
 * --------------------------------------------------------------------------------------------------------- */
size_t dagnn_operand_image_bytes(int64_t N, int32_t D);
int dagnn_embed_f32(const int64_t* x, const int64_t* depth, const float* type_tab, const float* attr_tab,
                    const float* depth_tab, int max_depth, int64_t n_types, int64_t n_attrs, int64_t N, int D, float* X,
                    int64_t ldx, void* x_image, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Integer pre-pass (bit-exact): level-sorted node order + in-edge CSR per direction.
 * replaces the per-level boolean-mask selects and the per-node O(E) edge scans of
 * ogbg-code/model/dagnn.py:130-137,146-147,151-157 (dvae/dagnn.py:104,111-112,116-122).
 *
 * For direction d, "position" p enumerates nodes sorted by (level_d, index in the level array) — i.e.
 * positions [lvl_off[d][l], lvl_off[d][l+1]) are exactly the reference's `layer` list of level l, in its order.
 * Row p of the CSR lists the edges e with edge_index[1-d][e] == perm[d][p] in ascending e (the reference's
 * `le_idx` order); col = position (in direction d's order) of the neighbour edge_index[d][e].
 * --------------------------------------------------------------------------------------------------------- */
typedef struct DagnnSchedule {
  int64_t N, E, B;
  int32_t dirs;        /* 1 or 2                                                              */
  int32_t max_levels;  /* capacity of lvl_off (entries: max_levels + 1)                        */
  int32_t* perm[DAGNN_MAX_DIRS];    /* [N]   position -> node id                               */
  int32_t* pos[DAGNN_MAX_DIRS];     /* [N]   node id  -> position                              */
  int32_t* lvl_off[DAGNN_MAX_DIRS]; /* [max_levels+1] first position of each level; [L] = N    */
  int32_t* rowptr[DAGNN_MAX_DIRS];  /* [N+1] CSR row pointers, rows indexed by position        */
  int32_t* col[DAGNN_MAX_DIRS];     /* [E]   neighbour position                                */
  int32_t* eid[DAGNN_MAX_DIRS];     /* [E]   original edge id (ascending inside a row)         */
  float* eattr[DAGNN_MAX_DIRS];     /* [E,2] edge_attr rows in CSR order, or NULL              */
  int32_t* gptr;                    /* [B+1] first node id of each graph (batch vector sorted) */
  int32_t* gdepth;                  /* [B]   number of levels of each graph (max forward level + 1), or NULL: the cluster sweep
                                             then cuts its graph groups by node count alone                         */
  /* summary[0]=num_levels of dir 0, [1]=num_levels of dir 1, [2]=status (0 ok, 1 level >= max_levels,
   * 2 node id / edge endpoint out of range), [3] = 1 when a level array carries node ids other than 0..N-1 in order
   * (positions inside a level are then not sorted by node id), [4..7] reserved.                       */
  int32_t* summary;                 /* [8]                                                     */
} DagnnSchedule;

size_t dagnn_schedule_workspace_bytes(int64_t N, int64_t E, int32_t max_levels);

/* edge_index int64 [2,E]; lvl{0,1} int64 [N] levels; nid{0,1} int64 [N] node id of each level-array entry
 * (NULL = identity; the reference's _bi_layer_index{0,1}); edge_attr fp32 [E,2] or NULL; batch int64 [N]
 * (sorted, PyG convention) or NULL when B == 0. All output arrays of *sched must be allocated by the caller. */
int dagnn_schedule_build(const int64_t* edge_index, const int64_t* lvl0, const int64_t* lvl1,
                         const int64_t* nid0, const int64_t* nid1, const float* edge_attr, const int64_t* batch,
                         const DagnnSchedule* sched, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Input side: longest-path level of every node of a batch of DAGs, on the edges as given
 * (lvl_fwd) and on the reversed edges (lvl_bwd) — what the reference computes per graph on the host when a data set is
 * built and stores as `_bi_layer_idx0/1` / `bi_layer_index[d][0]`: `top_sort` src/utils_dag.py:8-35,
 * `add_order_info_01` :39-52, `add_order_info` :70-76 (the node-id rows are arange(N)). OGB: pass the raw AST edges,
 * not the augmented ones (ogb/io/read_graph_pyg.py:51 runs before augment_edge2).
 * edge_index int64 [2, E] (row 0 sources, row 1 targets), outputs int64 [N]. Bit-exact (integers).
 * summary: device int32 [4], written by the call: [0] number of levels (max lvl_fwd + 1), [1] status — 0 ok, 1 not at the
 * fixed point after max_passes passes over the edges (batch deeper than max_passes - 1, or a cycle: call again with
 * more passes; more than N passes means the edge list is not acyclic), 2 an edge endpoint outside [0, N) —,
 * [2] max lvl_bwd + 1. Asynchronous; max_passes + 2 small launches, passes after the fixed point return at once.
 * --------------------------------------------------------------------------------------------------------- */
size_t dagnn_levels_workspace_bytes(int64_t N, int32_t max_passes);
int dagnn_levels_build(const int64_t* edge_index, int64_t N, int64_t E, int32_t max_passes, int64_t* lvl_fwd, int64_t* lvl_bwd,
                       int32_t* summary, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Input side: D-VAE text rows -> collated batch, on the device. Replaces decode_ENAS_to_pygraph / decode_BN_to_pygraph
 * (dvae/util.py:343-385 / :290-339) and the collation of dvae/batch.py:26-145 for a batch of B rows of n variables each.
 * rows int32 [B, n, n]: rows[g][i][0] = raw type of variable i, rows[g][i][1 + j] (j < i) = its flag for variable j (the text
 * row [[t0], [t1, f10], [t2, f20, f21], ...] padded to n columns). kind 0 = ENAS / NA, 1 = BN. Graph g gets n + 2 nodes
 * (start, variables, end), node ids g (n + 2) + v.
 * Outputs (caller-allocated): x fp32 [B (n + 2), nvt] one-hot; edge_index int64 [2, ecap] (edges in the reference's order:
 * non-zeros of the adjacency row-major, graph after graph); bi_layer_index int64 [2, 2, N] ([d][0] levels, [d][1] node ids);
 * batch int64 [N]; counts int32 [4]: [0] = number of edges E (<= ecap or the tail was dropped: (n + 2)(n + 1) / 2 per graph
 * is always enough), [1] = status (2: a node type outside [0, nvt)). Bit-exact integers.
 * --------------------------------------------------------------------------------------------------------- */
size_t dagnn_dvae_rows_workspace_bytes(int64_t B);
int dagnn_dvae_rows_build(const int32_t* rows, int64_t B, int32_t n, int32_t kind, int32_t nvt, float* x, int64_t* edge_index, int64_t ecap,
                          int64_t* bi_layer_index, int64_t* batch, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Parameter packing for one (direction, layer): GRU weights -> fp16 hi/lo split, pre-swizzled shared-memory images
 * that the projection GEMM bulk-copies and feeds to tcgen05.mma; attention vector -> key part + edge-type coefficients.
 * Sources: nn.GRUCell weight_ih [3H,Din], weight_hh [3H,H], bias_ih/bias_hh [3H]   (dagnn.py:79-81),
 *          attn_lin.weight [1, Dq + H (+nvid)] (dagnn.py:359; dvae/dagnn.py:47-48,357),
 *          edge_encoder.weight [H,2] (dagnn.py:356) or NULL,
 *          weight_ih of the NEXT stacked layer [3H,H] (NULL for the last layer): the state of layer i is the operand of
 *          both W_hh^i (its own recurrence) and W_ih^{i+1} (the next layer's input), so the two ride in one image.
 * The query part of attn_lin (first Dq columns), attn_lin.bias and edge_encoder.bias add the same constant
 * to every in-edge score of a node and cancel in the softmax, so they are not packed.
 * Layout of `packed` (4-byte units), all offsets from dagnn_pack_layout():
 *   bias  [4][HP]               b_r=b_ir+b_hr, b_z=b_iz+b_hz, b_in, b_hn; HP = 64*ceil(H/64), zero padded
 *   wk    [HP]                  key weights on the hidden state
 *   attnc [4]                   {wk.W_e[:,0], wk.W_e[:,1], 0, 0}
 *   vidk  [nvid]                key weights on the one-hot vertex id (D-VAE NA), nvid may be 0
 *   imgx  [Mc/64][Kin64/64][hi,lo][64 rows][64 halfs]        W_ih over the layer input (first layer only)
 *   imgh  [(1|2) Mc/64][Kh64/64][hi,lo][64 rows][64 halfs]   [W_hh ; W_ih of the next layer] over this layer's state
 *         column c of a matrix = gate * Hq + unit (gates r, z, n; Hq = roundup(H,4)), zero-padded to Mc = roundup(3 Hq, 64);
 *         K-major SWIZZLE_128B tiles (dagnn_b200/csrc/tc.cuh) of w = hi + lo, hi = rn_f16(w), lo = rn_f16(w - hi).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct DagnnPackLayout {
  int32_t Din, H, nvid;
  int32_t first_layer, last_layer;
  int32_t Hq, Mc;                      /* roundup(H,4); padded columns of one projected matrix                 */
  int32_t Kin64, Kh64, HP;             /* K of the input / state operand padded to the 64-wide chunk; padded units */
  int64_t bias_off, wk_off, attnc_off, vidk_off, imgx_off, imgh_off, total_floats;
} DagnnPackLayout;

int dagnn_pack_layout(int32_t Din, int32_t H, int32_t nvid, int32_t first_layer, int32_t last_layer, DagnnPackLayout* out);

/* `packed` must be 16-byte aligned (the images are bulk-copied into swizzled shared memory). */
int dagnn_pack_params_f32(const float* weight_ih, const float* weight_hh, const float* bias_ih,
                          const float* bias_hh, const float* attn_w, int32_t Dq, const float* edge_w,
                          const float* weight_ih_next, const DagnnPackLayout* layout, float* packed, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * The level sweep (the hot path): for every direction d, level l (sequential) and stacked layer i,
 *   m_v   = sum_e softmax_e( wk.h_e + wk.W_e a_e [+ vidk[nbr mod nvid]] ) * h_e     over ALL in-edges e of v,
 *           h_e = H[d][i][nbr(e)] if level_d[nbr(e)] < l else 0      (level 0: m_v = 0, edges ignored)
 *   inp_v = GRUCell_{d,i}(inp_v, m_v);  H[d][i][v] = inp_v          (inp_v starts as X[v])
 * replaces dagnn.py:144-182 incl. AttnConv (:362-373), PyG propagate/softmax/scatter-add, nn.GRUCell (:181) and
 * the index_put at :182; D-VAE variants dvae/dagnn.py:109-145, dvae/dagnn_bn.py:108-136.
 * ONE persistent cooperative kernel (one CTA per SM) runs the whole sweep. Since W_hh (sum_e a_e h_e) = sum_e a_e (W_hh h_e),
 * every node is projected ONCE, right after its state is final: P^i = W_hh^i h^i and Gi^{i+1} = W_ih^{i+1} h^i (one
 * weight image [W_hh ; W_ih next]), Gi^0 = W_ih^0 X. Wavefront step s = level + layer handles all (d, layer) pairs of
 * the step in two phases separated by grid barriers:
 *   gate phase  one warp per node (the whole CTA for long in-edge lists): softmax over the in-edges from per-node scalar
 *               key scores, weighted sum of the predecessors' P and h rows, GRU pointwise with the node's own Gi row;
 *               the new state is stored as fp32 and as fp16 hi/lo halves in the tcgen05 operand-image layout;
 *   proj phase  the rows the gate phase produced (contiguous: positions are level-sorted) x the weight image as
 *               128/256-row tiles: operand and weight tiles arrive by bulk copy (cp.async.bulk + mbarrier), the GEMM
 *               runs on tcgen05 (fp16 x 3 split, fp32 accumulators in TMEM), the epilogue stores P / Gi rows.
 * Level offsets and the level count (reference: max level of direction 0, + 1 — dagnn.py:137) are read from the
 * schedule's DEVICE arrays, so a forward needs no host synchronisation. If sched->summary[2] != 0 (bad input) the
 * kernel does nothing; the caller checks the status. States are stored in POSITION order: H[d][i] is fp32 [N, ldh], row p = node perm[d][p].
 * Inputs must satisfy |x| < 65504 (fp16 range of the hi part); states are in (-1, 1) by construction.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct DagnnSweepArgs {
  const DagnnSchedule* sched;
  int32_t num_layers;
  int32_t Din, H, nvid;                /* input width, hidden width, #vertex-id columns (0 = none) */
  const float* X;                      /* [N, ldx] node features in NODE order                      */
  int64_t ldx;
  const void* X_image;                 /* optional operand image of X (dagnn_embed_f32), NULL = X rows are gathered */
  float* Hs[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];               /* [N, ldh] each, position order      */
  int64_t ldh;                         /* >= roundup(H,4), multiple of 4                            */
  const float* packed[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];     /* dagnn_pack_params_f32 outputs       */
  int32_t use_edge_attr;               /* 1: add the edge-type score term (sched->eattr must exist) */
  void* workspace;                     /* device scratch, dagnn_sweep_workspace_bytes(), 256-byte aligned */
  size_t workspace_bytes;
  void* trace;                         /* optional profiling buffer (NULL = off), dagnn_sweep_trace_bytes(steps):
                                          int64 [steps + 1][256][16] clock64 values per (phase pair, CTA): entry k = proj
                                          phase of the input (k = 0) or gate + proj phase of step k - 1. Slots: 0 begin,
                                          8 gate phase done, 9 grid barrier passed, 1 first tile: operands ready,
                                          2 first tile: accumulators ready, 3 first tile stored, 4 all tiles done,
                                          5 second grid barrier passed, 6 #tiles of this CTA, 7 columns | rows << 12 of a
                                          tile; issuer warp, first tile: 10 start, 11 first operands landed, 13 / 14 chunk 0 / 1
                                          issued, 12 all MMAs issued (10..15 are gate-phase stage stamps instead in builds
                                          with -DDAGNN_GATE_TRACE)  */
} DagnnSweepArgs;

size_t dagnn_sweep_workspace_bytes(int32_t dirs, int32_t layers, int32_t Din, int32_t H, int64_t N, int64_t E, int32_t max_levels);
size_t dagnn_sweep_trace_bytes(int32_t max_steps);
int dagnn_sweep_forward_f32(const DagnnSweepArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Readout: out[g, out_col + c] = pool over selected nodes v of graph g of src[row(v), c]
 * replaces dagnn.py:119-126,184-202 (cat over layers + mask/gather of output nodes + global_{max,mean,add}_pool)
 * and the last/first-node gathers of dvae/dagnn.py:147-161, dvae/dagnn_bn.py:138-152.
 *   row(v)  = v (index_mode 0, node order) or sched->pos[dir][v] (index_mode 1, position order)
 *   filter  : 0 all nodes, 1 nodes with filter_lvl[v] == 0, 2 last node of the graph, 3 first node
 *   pool    : 0 max, 1 mean, 2 add      (empty selection -> 0, like torch_scatter)
 * --------------------------------------------------------------------------------------------------------- */
typedef struct DagnnReadoutBlock {
  const float* src;
  int64_t ld;
  int32_t width;
  int32_t index_mode;
  int32_t dir;
  int32_t filter;
  const int64_t* filter_lvl;
  int32_t out_col;
  int32_t reserved;
} DagnnReadoutBlock;

int dagnn_readout_f32(const DagnnSchedule* sched, const DagnnReadoutBlock* blocks, int32_t nblocks, int32_t pool,
                      float* out, int64_t ldo, void* stream);

/* Diagnostics of the tcgen05 building blocks the level kernels use (K-major SWIZZLE_128B operand tiles, fp16 x 3 split on
 * kind::f16, TMEM accumulators), fp32 row-major in and out:
 *   f16x3: C[M,N] = A[M,K] * B[N,K]^T, both operands from shared memory. N % 16 == 0, 16 <= N <= 256, K % 64 == 0.
 *   ts   : C[N,R] = X[N,K] * W[R,K]^T with W resident in TMEM as the A operand (hi rows on lanes 0..63, lo rows on lanes
 *          64..127) and the rows of X as the shared-memory B operand — the arrangement of the cluster sweep.
 *          R <= 64, N % 16 == 0, 16 <= N <= 256, K % 16 == 0, K <= 256. */
int dagnn_tc_selftest_f16x3(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream);
int dagnn_tc_selftest_ts(const float* W, const float* X, float* C, int32_t R, int32_t N, int32_t K, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Small dense layers on the tensor cores (fp16 x 3 split, fp32-grade accuracy), no cuBLAS:
 *   dagnn_linear_f32   y[M, N] = x[M, K] w[N, K]^T + bias[N]      nn.Linear.forward — the heads (ogbg-code/model/dagnn.py:209-215),
 *                      out_linear / hg_unify (dvae/dagnn.py:156,161), fc1 / fc2 (:183). bias may be NULL.
 *   dagnn_gemm_f32     C[M, N] (+)= A B^T with each operand either [rows, K] row-major (kmajor = 1) or its transposed view
 *                      [K, rows] (kmajor = 0): the backward of the above (dx = dy w: B = w as [K = N_out, rows = K_in];
 *                      dw = dy^T x: both operands transposed views) and of the GRU cells.
 * --------------------------------------------------------------------------------------------------------- */
int dagnn_linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y, int64_t ldy, int32_t M,
                     int32_t N, int32_t K, void* stream);
int dagnn_gemm_f32(const float* A, int64_t lda, int32_t a_kmajor, const float* B, int64_t ldb, int32_t b_kmajor, float* C, int64_t ldc,
                   int32_t M, int32_t N, int32_t K, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Backward of the path: what torch.autograd derives from ogbg-code/model/dagnn.py:141-202 and
 * ogbg-code/utils.py:26-28, so that loss.backward() (main_pyg.py:55-65, dvae/train.py:255-264) runs on this library.
 *   dagnn_readout_backward_f32  gradient of the pooled readout into the gradient buffers of its sources. grad_blocks[k] describes
 *                               block k as in the forward with `src` = the GRADIENT buffer (same shape / ld as the forward source,
 *                               accumulated with atomicAdd: zero it first), fwd_src[k] = the forward source, out / dout [B, ldo].
 *   dagnn_sweep_backward_f32    reverse-level BPTT through GRU cells and attention. dHs[d][i] holds d loss / d H[d][i] (position
 *                               order) on entry and is clobbered; parameter gradients are overwritten; dX (node order, may be
 *                               NULL) is accumulated (+=). Needs HOST copies of the level offsets (the forward's schedule has
 *                               been finalized by then). The query part of attn_lin.weight gets zeros (its exact gradient).
 *   dagnn_embed_backward_f32    dT[x0] += dX, dA[x1] += dX, dP[min(depth, max_depth)] += dX (atomicAdd; zero the tables' gradients first).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct DagnnCellParams {      /* raw parameters of one (direction, layer): device pointers, row-major like torch */
  const float* weight_ih;             /* [3H, Din_i]                                  */
  const float* weight_hh;             /* [3H, H]                                      */
  const float* bias_ih;               /* [3H]                                         */
  const float* bias_hh;               /* [3H]                                         */
  const float* attn_w;                /* attn_lin.weight [1, Dq + H + nvid]           */
  const float* edge_w;                /* edge_encoder.weight [H, 2] or NULL           */
  int32_t Dq;
  int32_t reserved;
} DagnnCellParams;
typedef struct DagnnCellGrads {       /* same shapes; all written by dagnn_sweep_backward_f32 */
  float* weight_ih;
  float* weight_hh;
  float* bias_ih;
  float* bias_hh;
  float* attn_w;
  float* edge_w;                      /* NULL when there is no edge encoder */
} DagnnCellGrads;
typedef struct DagnnSweepBwdArgs {
  const DagnnSchedule* sched;
  int32_t num_layers, Din, H, nvid;
  const float* X;                     /* [N, ldx] node order (the forward's input) */
  int64_t ldx;
  const float* Hs[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];   /* forward states, position order, [N, ldh] */
  float* dHs[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];        /* in: gradient wrt the states; clobbered    */
  int64_t ldh;
  DagnnCellParams params[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
  DagnnCellGrads grads[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
  float* dX;                          /* [N, lddx] node order, accumulated; NULL = not needed */
  int64_t lddx;
  int32_t use_edge_attr;
  int32_t num_levels;                 /* summary[0] of the schedule                                */
  const int32_t* lvl_off_host[DAGNN_MAX_DIRS];         /* HOST copies of sched->lvl_off[d], num_levels + 1 entries */
  void* workspace;                    /* dagnn_sweep_backward_workspace_bytes(), 256-byte aligned  */
  size_t workspace_bytes;
} DagnnSweepBwdArgs;

size_t dagnn_sweep_backward_workspace_bytes(int32_t Din, int32_t H, int32_t nvid, int64_t N, int64_t E);
int dagnn_sweep_backward_f32(const DagnnSweepBwdArgs* args, void* stream);
int dagnn_readout_backward_f32(const DagnnSchedule* sched, const DagnnReadoutBlock* grad_blocks, const float* const* fwd_src, int32_t nblocks,
                               int32_t pool, const float* out, const float* dout, int64_t ldo, void* stream);
int dagnn_embed_backward_f32(const int64_t* x, const int64_t* depth, int max_depth, int64_t n_types, int64_t n_attrs, int64_t N, int D,
                             const float* dX, int64_t ldx, float* d_type_tab, float* d_attr_tab, float* d_depth_tab, void* stream);

/* Un-permute states for inspection / tests: dst[v,:] = src[pos[dir][v],:]  (fp32 [N,H]) */
int dagnn_states_to_node_order_f32(const DagnnSchedule* sched, int32_t dir, const float* src, int64_t lds,
                                   int32_t H, float* dst, int64_t ldd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAGNN_B200_H_ */
