/*
 * demf_b200.h -- C ABI of libdemf_b200.so, the B200 (sm_100a) replacement for the
 * native ops that haoy945/DeMF executes on its data-parallel hot path.
 *
 * The reference repository contains no native code: every entry point below
 * replaces a pybind function of a pinned third-party extension
 * (requirements.txt:2-4: mmcv_full==1.3.18, mmdet3d==0.18.1) that the reference
 * reaches through
 *     demf/modeling/heads/class_agnostic_vote_head.py:13   (build_sa_module, furthest_point_sample)
 *     demf/modeling/heads/class_agnostic_vote_head.py:429  (furthest_point_sample call)
 *     demf/modeling/heads/class_agnostic_vote_head.py:455  (vote aggregation SA module)
 *     demf/modeling/layers/transformer.py:9,73-78          (MultiScaleDeformableAttention)
 *     configs/demf/demf_votenet.py:48-62                   (PointNet2SASSG backbone)
 * Each declaration names the upstream binding it stands in for.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; device pointers unless stated otherwise;
 *   - all tensors dense, row-major ("contiguous"), float32 / int32 / int64;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the library never allocates, never synchronises, never exits the process;
 *   - return value: 0 = launched; >0 = cudaError_t of the failed launch;
 *     <0 = argument check failed (DEMF_E_*); demf_last_error_string() explains;
 *   - the caller owns every buffer; outputs marked "pre-zeroed" must be zero on
 *     entry (the kernels accumulate into them with atomics).
 *
 * The CPU oracle (oracle/demf_oracle.c, test infrastructure only) exports
 * host-pointer twins `demf_ref_*` with the same argument lists minus
 * workspace/stream.
 */
#ifndef DEMF_B200_H_
#define DEMF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMF_B200_VERSION 100 /* major*10000 + minor*100 + patch */

enum {
  DEMF_OK = 0,
  DEMF_E_NULL = -1,        /* a required pointer is NULL */
  DEMF_E_SIZE = -2,        /* a size is negative / zero where it must not be / overflows int32 indexing */
  DEMF_E_UNSUPPORTED = -3, /* configuration outside what the kernels implement */
  DEMF_E_WORKSPACE = -4    /* workspace required but not supplied */
};

/* Library version (DEMF_B200_VERSION of the build). */
int demf_version(void);

/* Human-readable description of the last non-zero return on this thread. */
const char* demf_last_error_string(void);

/* Number of kernels this library has launched in this process (all streams);
 * used by bench.py for the "gpu_launches" claim. */
uint64_t demf_launch_count(void);

/* ---------------------------------------------------------------- FPS ---- */
/* replaces mmdet3d furthest_point_sample_ext.furthest_point_sampling_wrapper
 * (call sites: class_agnostic_vote_head.py:429-430; every SA module of
 * configs/demf/demf_votenet.py:48-62).
 *   xyz (B,N,3) f32 -> idx (B,m) i32.  idx[b,0] = 0; iterative D-FPS with the
 * upstream block-tree tie rule (see oracle/demf_oracle.c: demf_ref_fps).
 * `workspace` is only read when demf_fps_workspace_bytes() > 0. */
size_t demf_fps_workspace_bytes(int B, int N, int m);
int demf_fps(const float* xyz, int B, int N, int m, void* workspace, int32_t* idx,
             float* new_xyz /* optional (B,m,3): coordinates of the picked points, or NULL */, void* stream);

/* ----------------------------------------------------------- ball query -- */
/* replaces mmdet3d ball_query_ext.ball_query_wrapper (QueryAndGroup inside each
 * PointSAModule; demf_votenet.py:52-53,155-162).
 *   xyz (B,N,3), new_xyz (B,M,3) -> idx (B,M,nsample) i32. Every row is written
 * (rows whose ball is empty are written as zeros, which is what upstream's
 * pre-zeroed output holds). */
int demf_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M,
                    float min_radius, float max_radius, int nsample, int32_t* idx, void* stream);

/* ---------------------------------------------------- grouping / gather -- */
/* replaces mmdet3d group_points_ext.forward / backward.
 *   features (B,C,N), idx (B,M,ns) -> out (B,C,M,ns); bwd: grad_out (B,C,M,ns)
 *   -> grad_features (B,C,N), pre-zeroed. */
int demf_group_fwd(const float* features, const int32_t* idx, int B, int C, int N, int M, int ns,
                   float* out, void* stream);
int demf_group_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N, int M, int ns,
                   float* grad_features, void* stream);

/* replaces mmdet3d gather_points_ext.gather_points_wrapper / _grad_wrapper.
 *   features (B,C,N), idx (B,M) -> out (B,C,M); bwd -> grad_features (B,C,N), pre-zeroed. */
int demf_gather_fwd(const float* features, const int32_t* idx, int B, int C, int N, int M,
                    float* out, void* stream);
int demf_gather_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                    float* grad_features, void* stream);

/* Fused QueryAndGroup (ball query + group xyz + centre subtraction + radius
 * normalisation + feature group + channel concat) = mmdet3d
 * ops/group_points/group_points.py:QueryAndGroup.forward as one launch.
 *   xyz (B,N,3), features (B,C,N) or NULL (C=0), new_xyz (B,M,3)
 *   -> idx (B,M,ns) i32 and out (B, (use_xyz?3:0)+C, M, ns) f32
 * The xyz rows come first on the channel axis (upstream torch.cat order). */
int demf_query_and_group_fwd(const float* xyz, const float* features, const float* new_xyz,
                             int B, int N, int M, int C, float min_radius, float max_radius, int ns,
                             int use_xyz, int normalize_xyz, int32_t* idx, float* out, void* stream);

/* --------------------------------------------- three_nn / interpolate ---- */
/* replaces mmdet3d interpolate_ext.three_nn_wrapper.
 *   unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (SQUARED distances, the
 *   Python wrapper takes the square root exactly as upstream's does), idx (B,n,3) i32. */
int demf_three_nn(const float* unknown, const float* known, int B, int n, int m,
                  float* dist2, int32_t* idx, void* stream);

/* replaces interpolate_ext.three_interpolate_wrapper / _grad_wrapper.
 *   features (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n);
 *   bwd: grad_out (B,C,n) -> grad_features (B,C,m), pre-zeroed. */
int demf_three_interpolate_fwd(const float* features, const int32_t* idx, const float* weight,
                               int B, int C, int m, int n, float* out, void* stream);
int demf_three_interpolate_bwd(const float* grad_out, const int32_t* idx, const float* weight,
                               int B, int C, int n, int m, float* grad_features, void* stream);

/* ------------------------------------------- point-major ("rows") layout -- */
/* The layout the B200 backbone runs on: features are kept point-major, (B,N,C), so that a
 * neighbour is ONE contiguous row. These entry points replace the same upstream bindings as
 * demf_query_and_group_fwd / demf_group_bwd / demf_three_interpolate_* (mmdet3d
 * ball_query_ext + group_points_ext + interpolate_ext, reached from PointSAModule /
 * PointFPModule of configs/demf/demf_votenet.py:48-62,155-162) for a caller that keeps
 * activations as GEMM-ready rows.
 *
 * Row width K = demf_group_rows_width(C) = roundup(C,4) + 4 floats:
 *   out[b,m,s,:] = [ feat[b,idx,0:C] | 0-pad to a multiple of 4 | (xyz[idx]-centre)(/r) | 0 ]
 * query != 0: run the ball query (idx is an OUTPUT, same rows as demf_ball_query);
 * query == 0: idx is an INPUT. */
int demf_group_rows_width(int C);
int demf_query_and_group_rows_fwd(const float* xyz, const float* feat_rows, const float* new_xyz,
                                  int B, int N, int M, int C, float min_radius, float max_radius,
                                  int ns, int normalize_xyz, int query,
                                  const void* grid /* demf_ball_grid_build workspace or NULL */,
                                  int32_t* idx, float* out, void* stream);
/* grad_out (B,M,ns,K) -> grad_feat_rows (B,N,C) pre-zeroed (may be NULL when C == 0),
 * grad_xyz (B,N,3) pre-zeroed or NULL, grad_centre (B,M,3) fully written or NULL.
 * xyz_scale = 1/max_radius when normalize_xyz was set, else 1. */
int demf_group_rows_bwd(const float* grad_out, const int32_t* idx, int B, int N, int M, int C, int ns,
                        float xyz_scale, float* grad_feat_rows, float* grad_xyz, float* grad_centre,
                        void* stream);
/* feat_rows (B,m,C), idx (B,n,3), weight (B,n,3) -> out (B,n,C); C % 4 == 0.
 * bwd: grad_out (B,n,C) -> grad_feat_rows (B,m,C) pre-zeroed. */
int demf_three_interpolate_rows_fwd(const float* feat_rows, const int32_t* idx, const float* weight,
                                    int B, int C, int m, int n, float* out, void* stream);
int demf_three_interpolate_rows_bwd(const float* grad_out, const int32_t* idx, const float* weight,
                                    int B, int C, int n, int m, float* grad_feat_rows, void* stream);

/* development: CTA-level trace of the instrumented kernels into a caller-owned DEVICE buffer of
 * `capacity` 32-byte records {int kernel, block, smid, pad; uint64 start_ns, end_ns} plus a device
 * counter (zero it first); records == NULL switches tracing off. Kernel ids: 1 fps, 2 sa_fused,
 * 3 ball_query_grid, 4 ball_grid_build. */
int demf_trace_set(void* records, unsigned* counter, unsigned capacity);

/* ------------------------------------------- exact grid ball query ------- */
/* The same (B,M,ns) index rows as demf_ball_query, but each centre only tests the points of its
 * 3x3x3 cell neighbourhood in a uniform grid (cell edge >= radius) instead of the whole cloud.
 * demf_ball_grid_build bins xyz (B,N,3) once (counting sort, one CTA per scene) into `workspace`
 * (demf_ball_grid_workspace_bytes(B,N) bytes, 16-byte aligned, caller-owned); the workspace then
 * answers any query on the same xyz with max_radius <= radius (larger radii fall back to the full
 * scan inside the kernel). Accepted as the optional `grid` argument of demf_sa_fused_fwd and
 * demf_query_and_group_rows_fwd. ns <= 256. */
size_t demf_ball_grid_workspace_bytes(int B, int N);
int demf_ball_grid_build(const float* xyz, int B, int N, float radius, void* workspace, void* stream);
int demf_ball_query_grid(const float* xyz, const float* new_xyz, const void* grid, int B, int N, int M,
                         float min_radius, float max_radius, int ns, int32_t* idx, void* stream);
/* Furthest point sampling through the same workspace (any build radius): identical indices to
 * demf_fps (replaces furthest_point_sampling_wrapper for large clouds). The cloud sits cell-ordered
 * in the shared memory of a 2/4/8-CTA cluster per scene; a new sample only revisits the 32-point
 * blocks whose bounding box it can reach. N <= ~80k points. */
int demf_fps_grid(const float* xyz, const void* grid, int B, int N, int m, int32_t* idx,
                  float* new_xyz /* optional (B,m,3) or NULL */, void* stream);
/* The sampling CHAIN of PointNet2SASSG (configs/demf/demf_votenet.py:51: 20000 -> 2048 -> 1024 -> 512 -> 256, and
 * the head's 1024 -> 256 seed sampling, class_agnostic_vote_head.py:429-430): every level after the first samples
 * a cloud that is the previous level's pick sequence. demf_fps_grid_prefix also writes, per scene,
 * unique_prefix[b] = the first iteration whose arg-max was not provably unique (`certify` if none of the first
 * `certify` iterations -- only those are tracked). demf_fps_prefix, given
 * that certificate for a cloud that is such a pick sequence, returns idx = 0..m-1 (and the first m points) without
 * iterating whenever unique_prefix[b] >= m -- the greedy pick of iteration k on the full cloud is then also the
 * unique greedy pick on the subset -- and runs the ordinary kernel for the other scenes. Results are identical to
 * demf_fps in every case. */
int demf_fps_grid_prefix(const float* xyz, const void* grid, int B, int N, int m, int32_t* idx, float* new_xyz,
                         int32_t* unique_prefix, int certify /* iterations to certify, <= m */, void* stream);
int demf_fps_prefix(const float* xyz, int B, int N, int m, void* workspace, int32_t* idx, float* new_xyz,
                    const int32_t* unique_prefix, void* stream);

/* ------------------------- fused set abstraction (inference), tcgen05 --- */
/* replaces, for one PointSAModule forward in eval mode (mmdet3d
 * ops/pointnet_modules/point_sa_module.py; configs/demf/demf_votenet.py:48-62,155-162;
 * call site class_agnostic_vote_head.py:455): ball_query_wrapper + 2x group_points_wrapper
 * + the three Conv2d(1x1)+BN2d+ReLU layers + max_pool2d, in ONE launch.
 *   xyz (B,N,3), feat_rows (B,N,C) point-major or NULL when C == 0, new_xyz (B,M,3)
 *   -> out (B,M,c3) point-major = max over the ns neighbours of relu(W3 relu(W2 relu(W1 g + b1) + b2) + b3),
 *   g = [feat | 0-pad to 4 | (xyz[idx]-centre)(/r) | 0]  (the row of demf_query_and_group_rows_fwd).
 * wpack: the three BN-folded weight matrices (c1 x K, c2 x c1, c3 x c2; K = demf_group_rows_width(C),
 * columns in row order), each packed by demf_sa_pack_weights and concatenated; bias: c1+c2+c3 floats.
 * query != 0: the ball query runs in the kernel; idx (B,M,ns) is an optional OUTPUT (may be NULL);
 * query == 0: idx is an INPUT. TF32 tensor-core products, fp32 accumulation; indices exact.
 * Supported: ns in {16,32,64}; c1,c2,c3 multiples of 32 in [32,256] (demf_sa_fused_supported).
 * demf_sa_fused_error(): synchronising debug read of the kernel's protocol-time-out flag (0 = none). */
long demf_sa_pack_floats(int Cout, int Cin);
int demf_sa_pack_weights(const float* w /* (Cout,Cin) device */, int Cout, int Cin, float* packed,
                         void* stream);
int demf_sa_fused_supported(int C, int ns, int c1, int c2, int c3);
int demf_sa_fused_fwd(const float* xyz, const float* feat_rows, const float* new_xyz, int B, int N, int M,
                      int C, float min_radius, float max_radius, int ns, int normalize_xyz, int query,
                      const float* wpack, const float* bias, int c1, int c2, int c3,
                      const void* grid /* demf_ball_grid_build workspace of xyz, or NULL */,
                      int32_t* idx, float* out, void* stream);
int demf_sa_fused_error(void);
/* The same level as a warp-specialised PIPELINE over 128-row tiles (csrc/sa_pipe.cu) for the geometry whose weights
 * are resident: C = 1 feature channel, widths (64, 64, 128), ns in {16, 32, 64}, M a multiple of 128/ns -- the
 * backbone's first level. idx = the ball-query rows (B,M,ns) of the same query (demf_ball_query / _grid);
 * wpack / bias as for demf_sa_fused_fwd. Same results as demf_sa_fused_fwd. points4 (optional, may be NULL): the
 * same cloud packed as (B,N,4) rows [x y z feat], 16-byte aligned -- one load per neighbour instead of four. */
int demf_sa_pipe_supported(int C, int ns, int c1, int c2, int c3, int M);
int demf_sa_pipe_error(void);
int demf_sa_pipe_fwd(const float* xyz, const float* feat_rows, const float* points4, const float* new_xyz,
                     const int32_t* idx, int B, int N, int M, int ns, float max_radius, int normalize_xyz,
                     const float* wpack, const float* bias, float* out, void* stream);
/* debug only: device buffer of 64 int64 receiving [count, clock64 stamps of one worker thread of
 * CTA (0,0): start, after ball query, then per tile: gathered, acc0 ready, act1 written, acc1 ready,
 * act2 written, acc2 ready, stored]; NULL switches it off. */
int demf_sa_fused_set_profile(long long* device_buffer);
/* development knobs: most tile pipelines ("lanes") per CTA (1, 2 or 4; default 4) and the worker
 * warps' back-off between mbarrier polls in ns (default 0 = spin). */
int demf_sa_fused_tune(int max_lanes, int sleep_ns);

/* ----------------------------------------------------- fused glue (inference) --- */
/* Each replaces a chain of tiny library launches in the upstream Python modules:
 *  demf_chain_indices: PointNet2SASSG's sa_indices[i+1] = gather(sa_indices[i], 1, idx.long()) for
 *    up to 4 levels (mmdet3d models/backbones/pointnet2_sa_ssg.py); idx_l (B,m_l) i32 -> out_l i64.
 *  demf_interp_cat_rows_fwd: PointFPModule's sqrt -> 1/(d+1e-8) -> normalise -> three_interpolate ->
 *    cat([interpolated, skip]) from three_nn's SQUARED distances: src_rows (B,m,C1), skip_rows (B,n,C2)
 *    or NULL, idx (B,n,3), dist2 (B,n,3) -> out (B,n,C1+C2); C1, C2 multiples of 4.
 *  demf_decode_boxes: one prediction stage of DeMFClassAgnosticBBoxCoder.decode + the score softmaxes
 *    (demf/core/bbox/coders/class_agnostic_bbox_coder.py:168-194; upstream VoteHead.get_bboxes):
 *    row tensors (B,Q,*) given as pointer + row stride in floats -> box (B,out_rows,7),
 *    obj_prob (B,out_rows), sem_prob (B,out_rows,classes) at rows [out_offset, out_offset+Q). */
int demf_chain_indices(int B, int levels, const int32_t* idx0, int m0, const int32_t* idx1, int m1,
                       const int32_t* idx2, int m2, const int32_t* idx3, int m3, int64_t* out0,
                       int64_t* out1, int64_t* out2, int64_t* out3, void* stream);
int demf_interp_cat_rows_fwd(const float* src_rows, const float* skip_rows, const int32_t* idx,
                             const float* dist2, int B, int C1, int C2, int m, int n, float* out,
                             void* stream);
int demf_decode_boxes(const float* center, int s_center, const float* size, int s_size,
                      const float* dir_class, int s_dir_class, const float* dir_res, int s_dir_res,
                      const float* obj, int s_obj, const float* sem, int s_sem, int B, int Q, int bins,
                      int classes, int out_rows, int out_offset, float* box, float* obj_prob,
                      float* sem_prob, void* stream);

/* ------------------------------------------------ training BatchNorm rows --- */
/* Training-mode BatchNorm (+ ReLU) over point-major rows x (R,C): replaces, for the 1x1-conv ConvModules of
 * the shared MLPs (mmcv ConvModule conv -> BN -> ReLU as built by mmdet3d PointSAModule / PointFPModule /
 * VoteModule / BaseConvBboxHead), torch's batch_norm collect_statistics / update_stats / transform_input /
 * clamp forward and threshold / backward_reduce / backward_elemt backward kernels by two launches each way.
 *   fwd: y = [relu]((x - mean_batch) * invstd * gamma + beta); save_mean / save_invstd (C) for backward;
 *        running_mean / running_var (may both be NULL) updated with `momentum` (unbiased variance).
 *   bwd: grad_x (R,C), grad_gamma (C), grad_beta (C) from grad_y, the forward's y (ReLU mask), x and the
 *        saved statistics; coef = scratch of 2*C floats.
 * state: persistent device block of demf_bn_rows_state_bytes(C) bytes owned by the layer, ZERO before the
 * first call (each call leaves it zero); calls sharing a state must be stream-ordered.
 * C = 4 * 2^k <= 1024 (demf_bn_rows_supported), all pointers 16-byte aligned. */
int demf_bn_rows_supported(int C);
long demf_bn_rows_state_bytes(int C);
int demf_bn_rows_fwd(const float* x, long R, int C, const float* gamma, const float* beta, float eps,
                     float momentum, int relu, float* running_mean, float* running_var, void* state,
                     float* save_mean, float* save_invstd, float* y, void* stream);
int demf_bn_rows_bwd(const float* grad_y, const float* y, const float* x, long R, int C, const float* gamma,
                     const float* save_mean, const float* save_invstd, int relu, void* state, float* coef,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* stream);

/* The LAST layer of a set-abstraction MLP in training, with the max over the neighbourhood folded in
 * (mmdet3d PointSAModule: ... -> BN -> ReLU -> F.max_pool2d over nsample): x (M*ns, C) rows, ns consecutive rows per
 * centre -> pooled (M,C) = max_r relu(bn(x)) and arg (M,C) u8 = the row attaining it. The normalised (M*ns,C)
 * tensor is never written; backward takes grad_pooled (M,C) and returns the dense grad_x (M*ns,C). */
int demf_bn_max_rows_fwd(const float* x, long M, int ns, int C, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, void* state, float* save_mean,
                         float* save_invstd, float* pooled, uint8_t* arg, void* stream);
int demf_bn_max_rows_bwd(const float* grad_pooled, const float* pooled, const uint8_t* arg, const float* x, long M,
                         int ns, int C, const float* gamma, const float* save_mean, const float* save_invstd,
                         void* state, float* coef, float* grad_x, float* grad_gamma, float* grad_beta,
                         void* stream);

/* Scaled-dot-product attention of one nn.MultiheadAttention call in inference (csrc/mha.cu), exact fp32: q (Lq*B, .)
 * rows with stride ldq floats -- token l of scene b is row l*B + b (b*L + l when batch_first), head h occupies columns
 * [h*D, h*D + D) -- k / v
 * (Lk*B, .) likewise; o (Lq*B, .) = softmax(q k^T * scale) v per (scene, head). D in {32, 36, 64}; row starts and
 * strides 16-byte aligned; K and V of one head (Lk * D * 8 bytes) must fit 200 KB of shared memory. Replaces torch's
 * scaled_dot_product_attention (an sm80 memory-efficient kernel) under the decoder layer's self-attention among the
 * proposals (demf/modeling/layers/transformer.py:55-80; configs/demf/demf_votenet.py:76-78). */
int demf_mha_supported(int D);
int demf_mha_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, float* o, long ldo,
                 int Lq, int Lk, int B, int H, int D, float scale, int batch_first, void* stream);
/* out[c] += sum over the R rows of x (R, N), row stride ld floats: the bias gradient of a Linear / 1x1 convolution
 * accumulated in place (autograd's grad.sum(0) + accumulate in one launch). */
int demf_col_sum_add(const float* x, long R, int N, long ld, float* out, void* stream);
/* The normalise(+ReLU)(+max) half alone, when mean / invstd came from demf_bn_finalize (statistics accumulated by
 * the epilogue of the GEMM that produced x). */
int demf_bn_rows_apply(const float* x, long R, int C, const float* gamma, const float* beta, const float* mean,
                       const float* invstd, int relu, float* y, void* stream);
int demf_bn_max_rows_apply(const float* x, long M, int ns, int C, const float* gamma, const float* beta,
                           const float* mean, const float* invstd, float* pooled, uint8_t* arg, void* stream);
int demf_bn_rows_bwd_apply(const float* g, const float* x, long R, int C, const float* gamma, const float* mean,
                           const float* invstd, const float* coef, float* grad_x, void* stream);

/* ------------------------------------------- training GEMMs (tcgen05 + TMA) --- */
/* The 1x1 convolutions of mmcv ConvModule inside the shared MLPs in TRAINING (upstream: cuDNN / cuBLAS behind
 * torch.nn.Conv1d/Conv2d forward and backward; mmdet3d ops/pointnet_modules/point_sa_module.py,
 * point_fp_module.py, models/model_utils/vote_module.py, models/dense_heads/base_conv_bbox_head.py, built from
 * configs/demf/demf_votenet.py:48-62,142-162). Row-major fp32 matrices with row strides (`ld*`, in floats,
 * multiples of 4), TF32 products with fp32 accumulation on the tcgen05 tensor cores, operands and results moved
 * by TMA tensor maps.
 *   fwd   : y (R,N) = x (R,K) w(N,K)^T [+ bias] [ReLU]; bn_state != NULL (N <= 256): the per-column sum and sum
 *           of squares of y are ADDED to the (2,N) double accumulators at the head of the BatchNorm layer's
 *           state block (demf_bn_rows_state_bytes); demf_bn_finalize turns them into mean / invstd.
 *   dgrad : dx (R,K) = dy (R,N) w(N,K)
 *   wgrad : dw (N,K) += dy (R,N)^T x (R,K)   (accumulates: the caller zeroes or owns the running gradient)
 * demf_gemm_error(): nonzero after an internal pipeline time-out (sticky; never expected). */
int demf_gemm_supported(int K, int N);
int demf_gemm_error(void);
int demf_gemm_tune(int epilogue_groups); /* development: 1 or 2 epilogue warp groups (default 2) */
int demf_gemm_debug_mn(int sbo, int layout, int tma_swizzle); /* development: MN-major operand encoding */
int demf_gemm_rows_fwd(const float* x, long ldx, const float* w, long ldw, const float* bias, long R, int K, int N,
                       int relu, void* bn_state, float* y, long ldy, void* stream);
int demf_gemm_rows_dgrad(const float* dy, long lddy, const float* w, long ldw, long R, int N, int K, float* dx,
                         long lddx, void* stream);
int demf_gemm_wgrad(const float* dy, long lddy, const float* x, long ldx, long R, int N, int K, float* dw, int ldw,
                    void* stream);
int demf_bn_finalize(void* state, long R, int C, float eps, float momentum, float* save_mean, float* save_invstd,
                     float* running_mean, float* running_var, void* stream);
/* Data gradient THROUGH the previous layer's BatchNorm + ReLU: g (R,K) = (dy (R,N) w(N,K)) masked where
 * relu(bn(y_prev)) was inactive (y_prev (R,K) = that layer's pre-BN activations), and sum g / sum g*y_prev per
 * channel ADDED to the (2,K) accumulators of `bn_state` -- the two reductions of the BatchNorm backward, for
 * demf_bn_bwd_finalize (-> grad_gamma, grad_beta, coef (2,K)) and demf_bn_rows_bwd_apply (-> grad of y_prev).
 * K <= 256. Replaces mmcv ConvModule's conv-backward-data + threshold_backward + batch_norm_backward_reduce. */
int demf_gemm_rows_dgrad_bn(const float* dy, long lddy, const float* w, long ldw, long R, int N, int K,
                            const float* y_prev, long ldy_prev, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, void* bn_state, float* g, long ldg, void* stream);
int demf_bn_bwd_finalize(void* state, long R, int C, const float* mean, const float* invstd, float* grad_gamma,
                         float* grad_beta, float* coef, void* stream);

/* ------------------------------------------- training: per-stage detection loss --- */
/* The seven weighted sums of DeMFVoteHead._loss (demf/modeling/heads/class_agnostic_vote_head.py:622-712;
 * mmdet CrossEntropyLoss / SmoothL1Loss, mmdet3d AxisAlignedIoULoss with reduction='sum', loss weights and betas of
 * configs/demf/demf_votenet.py:113-141) over `rows` = B*Q proposals in one launch, and their gradient with respect
 * to every prediction tensor in one more. Predictions (rows, C) fp32 contiguous; targets as the head assigns them
 * (int64 class targets, fp32 weights). sem / sem_t / g_sem may be NULL together (no semantic loss).
 * cfg (HOST, 12 floats): objectness class weights (2); loss_weight of objectness, dir_class, dir_res, size, center,
 * semantic, iou (7; iou 0 = no IoU loss); SmoothL1 beta of dir_res, size, center (3).
 * fwd: losses (7 device floats, ZERO on entry) += weighted sums in the order objectness, dir_class, dir_res, size,
 * center, semantic, iou. bwd: upstream (7 device floats) = dL/dloss_i; g_* get d(sum_i upstream_i loss_i)/d(pred). */
int demf_stage_loss_fwd(const float* center, const float* size, const float* dir_class, const float* dir_res_norm,
                        const float* obj, const float* sem, const int64_t* obj_t, const float* obj_w,
                        const float* box_w, const float* size_t_, const float* center_t, const int64_t* dir_class_t,
                        const float* dir_res_t, const int64_t* sem_t, long rows, int num_dir_bins, int num_sem,
                        const float* cfg, float* losses, void* stream);
int demf_stage_loss_bwd(const float* center, const float* size, const float* dir_class, const float* dir_res_norm,
                        const float* obj, const float* sem, const int64_t* obj_t, const float* obj_w,
                        const float* box_w, const float* size_t_, const float* center_t, const int64_t* dir_class_t,
                        const float* dir_res_t, const int64_t* sem_t, long rows, int num_dir_bins, int num_sem,
                        const float* cfg, const float* upstream, float* g_center, float* g_size, float* g_dir_class,
                        float* g_dir_res_norm, float* g_obj, float* g_sem, void* stream);

/* ------------------------------------------- inference post-processing --- */
/* The two per-scene loops of mmdet3d 0.18.1 VoteHead.multiclass_nms_single, reached from
 * DeMFVoteHead.get_bboxes (demf/modeling/heads/class_agnostic_vote_head.py:739-743):
 * demf_box_point_count replaces `bbox.points_in_boxes(points).T.sum(1)` (roiaware_pool3d
 * points_in_boxes_batch): counts (B,K) i32 = points of scene b inside box k; boxes (B,K,7) =
 * (x,y,z_bottom,dx,dy,dz,yaw), points rows of point_stride floats starting with xyz.
 * demf_aligned_3d_nms replaces mmdet3d core/post_processing aligned_3d_nms for a whole batch:
 * minmax (B,K,6) axis-aligned (x1,y1,z1,x2,y2,z2); scores (B,K); classes (B,K) i64; valid (B,K) u8 =
 * boxes taking part; keep (B,K) u8 <- 1 for picked boxes. Greedy by descending score (ties: larger
 * index first), a box is dropped unless iou * (same class) <= thresh. K <= 4096. */
int demf_box_point_count(const float* points, int point_stride, const float* boxes, int B, int N, int K,
                         int gravity_centre, int32_t* counts, void* stream);
/* gravity_centre != 0: boxes carry the decoded gravity centre instead of the bottom centre.
 * demf_nms_select: the rest of multiclass_nms_single up to the boolean selection, three launches for the batch:
 * per box the axis-aligned hull of its corners (-> minmax (B,K,6)), argmax class (-> classes i64) and
 * counts > min_points (-> valid u8); then demf_aligned_3d_nms; then selected &= obj_scores > score_thresh and
 * num_selected (B) i32 = selected boxes per scene (the one number the host has to read).
 * boxes (B,K,7) gravity-centre, sem_scores (B,K,C) probabilities. */
int demf_nms_select(const float* boxes, const float* obj_scores, const float* sem_scores, const int32_t* counts,
                    int B, int K, int C, int min_points, float nms_thresh, float score_thresh, float* minmax,
                    int64_t* classes, uint8_t* valid, uint8_t* selected, int32_t* num_selected, void* stream);
int demf_aligned_3d_nms(const float* minmax, const float* scores, const int64_t* classes, const uint8_t* valid,
                        int B, int K, float thresh, uint8_t* keep, void* stream);

/* Tail of mmdet3d VoteModule.forward (models/model_utils/vote_module.py; built by DeMFVoteHead,
 * class_agnostic_vote_head.py:371) for vote_per_seed = 1, inference: from the conv_out rows
 * votes (rows, ldv) = [offset(3) | residual(C) | padding] to vote_xyz (rows,3) = seed_xyz + clamp(offset),
 * offset (rows,3) and vote_rows (rows,C) = (seed_rows + residual), L2-normalised per row when norm_feats.
 * xyz_range: HOST pointer to 3 floats (vote_xyz_range) or NULL. C a multiple of 128 up to 512. */
int demf_vote_tail(const float* votes, int ldv, const float* seed_xyz, const float* seed_rows, long rows, int C,
                   const float* xyz_range, int norm_feats, float* vote_xyz, float* offset, float* vote_rows,
                   void* stream);

/* DeMFVoteHead.get_reference_points (class_agnostic_vote_head.py:524-547) for the whole batch: xyz (B,Q,3)
 * proposal centres, mats (B,3,4) and affs (B,4) = the per-scene projection chain folded on the host
 * (geometry.fold_projection) -> out (B,Q,2) image coordinates normalised to [0,1] and clamped. */
int demf_project_points(const float* xyz, const float* mats, const float* affs, int B, int Q, float* out,
                        void* stream);

/* Image pyramid (num_levels <= 8 tensors (B,C,H_l*W_l) f32, HOST array of device pointers and HOST array of
 * H_l*W_l) -> token rows out (B, sum H_l*W_l, C): the flatten(2).transpose(1,2) + cat of
 * DeMFVoteHead.prepare_decoder_inputs (demf/modeling/heads/class_agnostic_vote_head.py:570-591) and of
 * DeformableDetrEncoder.transformer (deform_detr_encoder.py:107-121) in one coalesced launch. */
int demf_levels_to_rows(const float* const* levels, const int* hw, int num_levels, int B, int C, float* out,
                        void* stream);

/* LayerNorm over rows with the preceding bias / residual adds folded in: replaces, for inference, the
 * `dropout(out) + identity` add and the nn.LayerNorm after every attention and FFN block of mmcv's
 * BaseTransformerLayer (post-norm) that the image-branch encoder runs
 * (demf/modeling/layers/deform_detr_encoder.py:141-151, configs/demf/demf_votenet.py:33-39).
 *   out[r,:] = LN(x[r,:] + bias[:] + residual[r,:]) * gamma + beta; bias and residual may be NULL.
 *   post_add / out2 (both or neither): out2[r,:] = out[r,:] + post_add[r,:], the `query + query_pos` of the
 *   attention that follows (mmcv MultiScaleDeformableAttention / MultiheadAttention forward).
 * C a multiple of 128 up to 1024, pointers 16-byte aligned; out may alias x. */
int demf_bias_layer_norm_rows(const float* x, const float* bias, const float* residual, const float* gamma,
                              const float* beta, long rows, int C, float eps, float* out, const float* post_add,
                              float* out2, void* stream);

/* ------------------------------------ multi-scale deformable attention --- */
/* replaces mmcv _ext.ms_deform_attn_forward / ms_deform_attn_backward
 * (MultiScaleDeformableAttnFunction; reached from transformer.py:73-78).
 *   value (B,S,H,D) f32; spatial_shapes (L,2) i64 [h,w]; level_start_index (L) i64;
 *   sampling_loc (B,Q,H,L,P,2) f32 [x,y in 0..1]; attn_weight (B,Q,H,L,P) f32
 *   -> out (B,Q,H*D) f32.
 * spatial_shapes / level_start_index are DEVICE pointers (as in upstream).
 * bwd: grad_out (B,Q,H*D) -> grad_value (B,S,H,D) pre-zeroed, grad_sampling_loc,
 * grad_attn_weight (both fully written). */
int demf_msda_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                  const float* sampling_loc, const float* attn_weight,
                  int B, int S, int H, int D, int Q, int L, int P, float* out, void* stream);
/* Inference form fed by the attention module's projections: replaces, in one launch, the chain
 * mmcv's MultiScaleDeformableAttention.forward runs between its Linear layers and the sampling
 * kernel (multi_scale_deform_attn.py:322-349: view, softmax over L*P, offsets / (W_l,H_l),
 * + reference points, contiguous, ms_deform_attn_forward).
 *   proj (B*Q, H*L*P*3) f32 rows = [sampling_offsets (H,L,P,2) | attention logits (H,L,P)];
 *   proj_add: NULL, or a second tensor of proj's shape added to it element by element first (the
 *   projection of the positional embedding, which is constant per image geometry);
 *   ref_points (B,Q,L,ref_dim) f32, ref_dim 2 (x,y) or 4 (x,y,w,h), already scaled by valid ratios;
 *   -> out (B,Q,H*D).
 * Supported when D = 4*2^k <= 128 and L*P is a power of two <= 32 dividing the warp's record count
 * (demf_msda_proj_fwd_supported returns 1); anything else returns DEMF_E_UNSUPPORTED and the
 * caller composes the steps around demf_msda_fwd. */
int demf_msda_proj_fwd_supported(int D, int L, int P);
int demf_msda_proj_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const float* proj, const float* proj_add, const float* ref_points, int ref_dim,
                       int B, int S, int H, int D, int Q, int L, int P, float* out, void* stream);
int demf_msda_bwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                  const float* sampling_loc, const float* attn_weight, const float* grad_out,
                  int B, int S, int H, int D, int Q, int L, int P,
                  float* grad_value, float* grad_sampling_loc, float* grad_attn_weight, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMF_B200_H_ */
