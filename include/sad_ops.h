/* sad_ops.h -- C ABI of libsad_b200.so: the set-abstraction / size-adaptive clustering
 * hot path of a 3DSAD-class detector, hand-written CUDA for sm_100a (B200).
 *
 * Reference interface replaced: NONE EXISTS TO CITE.  /root/reference holds only
 * README.md:1 ("# 3DSAD-main") and README.md:2 ("Size Adaptive Clustering for 3D
 * object detection in Point Clouds"); there is no extension module, setup.py or FFI
 * file.  Each entry point below therefore cites the SURVEY.md section 8(a) row whose
 * operator it implements (the PointNet++-lineage torch.autograd.Function surface that
 * BASELINE.json's north_star names); INTEGRATION.md shows the ctypes binding.
 *
 * Conventions:
 *  - plain C types only; every pointer is a DEVICE pointer owned by the caller;
 *    the library never allocates, frees or retains device memory across calls;
 *  - all tensors are dense, contiguous, row-major with the stated shape; float = IEEE
 *    fp32, idx = int32; bf16 tensors are passed as const void* / void*;
 *  - every call is asynchronous on `stream` (a cudaStream_t / CUstream handle; NULL =
 *    legacy default stream), performs no host<->device synchronisation and keeps no
 *    global mutable state besides a thread-local error string: scheduling choices (FPS policy / variant, fused-MLP
 *    grid width) are explicit arguments, and the library never reads the environment;
 *  - return value: SAD_OK (0) or a negative SAD_E* code; never throws, never aborts;
 *  - arithmetic contract: fp32, no FMA contraction,
 *    d2 = ((dx*dx)+(dy*dy))+(dz*dz), strict d2 < r*r, ties to the lowest index.
 */
#ifndef SAD_OPS_H_
#define SAD_OPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAD_OK 0
#define SAD_EINVAL (-1)       /* bad shape / null pointer / misalignment            */
#define SAD_ECUDA (-2)        /* CUDA launch / runtime error (see error string)      */
#define SAD_EUNSUPPORTED (-3) /* shape outside what the kernels are built for        */

#define SAD_ABI_VERSION 1

typedef void* sad_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define SAD_API __attribute__((visibility("default")))
#else
#define SAD_API
#endif

/* Library / ABI version (SAD_ABI_VERSION). */
SAD_API int sad_version(void);
/* Thread-local, NUL-terminated description of the last non-zero return on this thread. */
SAD_API const char* sad_last_error_string(void);

/* a1  furthest_point_sample(xyz, npoint): xyz (B,N,3) f32 -> idx (B,npoint) i32.
 * sel[0]=0, mind=1e10, next = argmax min-distance, ties -> lowest index.
 * N <= 204800 (register-resident cluster kernel); larger -> SAD_EUNSUPPORTED. */
SAD_API int sad_furthest_point_sample_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx,
                                  sad_stream_t stream);
/* a1 over PREFIX-ORDERED input: row k of xyz is the k-th pick of a farthest-point sampling of a superset (e.g. the
 * previous stage's new_xyz).  Returns exactly what sad_furthest_point_sample_fwd returns; scenes whose first npoint
 * rows hold no exact duplicate (checked on the device, contract arithmetic) get the identity without sampling, the
 * others run the sampler.  flags = B device int32 of scratch (1 = identity taken). */
SAD_API int sad_furthest_point_sample_prefix_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx, int* flags,
                                                 sad_stream_t stream);

/* ---- scene grid: spatial sort shared by the exact culled FPS and the grid ball query ----------
 * workspace (caller-owned, 16-byte aligned, sad_scene_grid_workspace_bytes(B,N) bytes) receives, per
 * scene, the points sorted by cell of a 32^3 grid (x,y,z,original index) and the cell offsets.  It
 * is valid for exactly the xyz it was built from.  Results of the *_grid_fwd entry points are
 * bit-identical to sad_furthest_point_sample_fwd / sad_ball_query(_adaptive)_fwd. */
SAD_API long long sad_scene_grid_workspace_bytes(int B, int N);
SAD_API int sad_scene_grid_build(int B, int N, const float* xyz, void* workspace, sad_stream_t stream);
/* a1 over the grid: bounding-box culling of the per-pick update; points resident in the shared memory
 * of a 1..16-CTA cluster (or, beyond that capacity, one CTA over the L2-resident sorted array: the
 * workspace's min-distance scratch is then written, so two FPS calls must not share a workspace).
 * N <= sad_fps_grid_max_points(), else SAD_EUNSUPPORTED. */
SAD_API int sad_fps_grid_max_points(void);
SAD_API int sad_furthest_point_sample_grid_fwd(int B, int N, int npoint, const float* xyz,
                                               void* grid_workspace, int32_t* idx, sad_stream_t stream);
/* Same result, explicit scheduling policy.  SAD_FPS_LATENCY (what sad_furthest_point_sample_grid_fwd uses): the
 * fewest-CTA cluster whose shared memory holds the scene -- shortest time per scene (40k points: 4 SMs, 0.70 us per
 * pick).  SAD_FPS_THROUGHPUT: ONE SM per scene over the L2-resident sorted array (1.1 us per pick, i.e. about 0.4 of
 * the SM-time per scene): what a pipelined caller with other kernels to overlap wants (writes the workspace's
 * min-distance scratch, so two calls must not share a workspace).
 * SAD_FPS_THROUGHPUT_PAIRED: two scenes share one SM (16 warps each): ~30 % less SM-time again, longer per scene.
 * `variant` (tests / tools; results never depend on it): 0 = the policy's kernel; 1,2,4,8,16 = the cluster kernel with
 * at least that many CTAs per scene; -1 = the single-SM kernel; -2 = the single-SM kernel restricted to its 16-warp
 * instances. */
#define SAD_FPS_LATENCY 0
#define SAD_FPS_THROUGHPUT 1
#define SAD_FPS_THROUGHPUT_PAIRED 2
SAD_API int sad_furthest_point_sample_grid_policy_fwd(int B, int N, int npoint, const float* xyz,
                                                      void* grid_workspace, int32_t* idx, int policy, int variant,
                                                      sad_stream_t stream);
/* a3 / a4 over the grid: radius_t (B,npoint) per-query radius or NULL (then `radius`). */
SAD_API int sad_ball_query_grid_fwd(int B, int N, int npoint, float radius, const float* radius_t,
                                    int nsample, const float* xyz, const void* grid_workspace,
                                    const float* new_xyz, int32_t* idx, sad_stream_t stream);

/* a2  gather_operation: out[b,c,j] = features[b,c,idx[b,j]].
 * features (B,C,N) f32, idx (B,npoint) i32 -> out (B,C,npoint) f32. */
SAD_API int sad_gather_operation_fwd(int B, int C, int N, int npoint, const float* features,
                             const int32_t* idx, float* out, sad_stream_t stream);
/* a2 backward: grad_features (B,C,N) is ZEROED by the call, then scatter-added. */
SAD_API int sad_gather_operation_bwd(int B, int C, int N, int npoint, const float* grad_out,
                             const int32_t* idx, float* grad_features, sad_stream_t stream);

/* a3  ball_query(radius, nsample, xyz, new_xyz): xyz (B,N,3), new_xyz (B,npoint,3)
 * -> idx (B,npoint,nsample) i32; first hit pads, no hit -> zeros. */
SAD_API int sad_ball_query_fwd(int B, int N, int npoint, float radius, int nsample, const float* xyz,
                       const float* new_xyz, int32_t* idx, sad_stream_t stream);
/* a4  ball_query_adaptive: as a3 with a per-query radius radius_t (B,npoint) f32
 * (the size-adaptive clustering search). */
SAD_API int sad_ball_query_adaptive_fwd(int B, int N, int npoint, const float* radius_t, int nsample,
                                const float* xyz, const float* new_xyz, int32_t* idx,
                                sad_stream_t stream);

/* a5  grouping_operation: out[b,c,j,s] = features[b,c,idx[b,j,s]].
 * features (B,C,N) f32, idx (B,npoint,nsample) i32 -> out (B,C,npoint,nsample) f32. */
SAD_API int sad_grouping_operation_fwd(int B, int C, int N, int npoint, int nsample,
                               const float* features, const int32_t* idx, float* out,
                               sad_stream_t stream);
/* a5 backward: grad_features (B,C,N) is ZEROED by the call, then scatter-added. */
SAD_API int sad_grouping_operation_bwd(int B, int C, int N, int npoint, int nsample,
                               const float* grad_out, const int32_t* idx, float* grad_features,
                               sad_stream_t stream);

/* a8  three_nn(unknown, known): unknown (B,n,3), known (B,m,3), m >= 3
 * -> dist (B,n,3) f32 (sqrt of d2, ascending), idx (B,n,3) i32. */
SAD_API int sad_three_nn_fwd(int B, int n, int m, const float* unknown, const float* known, float* dist,
                     int32_t* idx, sad_stream_t stream);
/* a8 + the FP module's inverse-distance weights in the same launch: weight (B,n,3) = r / ((r0 + r1) + r2),
 * r = 1 / (dist + 1e-8), evaluated in that order (bit-equal to the oracle's interpolation_weights). */
SAD_API int sad_three_nn_weights_fwd(int B, int n, int m, const float* unknown, const float* known, float* dist,
                                     int32_t* idx, float* weight, sad_stream_t stream);
/* a4 helper: predicted box size (rows,3) -> per-cluster radius = clamp(alpha/2 * ||size||_2, r_min, r_max), norm
 * evaluated as sqrt(((sx*sx)+(sy*sy))+(sz*sz)) (formula unpinned by the reference: SURVEY a4 DECISION). */
SAD_API int sad_size_to_radius(long long rows, const float* size, float alpha, float r_min, float r_max, float* radius,
                               sad_stream_t stream);
/* new_xyz (B,npoint,3) = xyz[b, inds[b,j], :] from the (B,N,3) layout in one launch; new_xyzw (optional, (B,npoint) float4
 * {x,y,z,0}) is the padded copy sad_sa_mlp_fwd gathers from. */
SAD_API int sad_gather_points_fwd(int B, int N, int npoint, const float* xyz, const int32_t* inds, float* new_xyz,
                                  void* new_xyzw, sad_stream_t stream);

/* a9  three_interpolate: out[b,c,i] = ((w0*f[i0]) + (w1*f[i1])) + (w2*f[i2]).
 * features (B,C,m) f32, idx (B,n,3) i32, weight (B,n,3) f32 -> out (B,C,n) f32. */
SAD_API int sad_three_interpolate_fwd(int B, int C, int m, int n, const float* features,
                              const int32_t* idx, const float* weight, float* out,
                              sad_stream_t stream);
/* a9 backward: grad_features (B,C,m) is ZEROED by the call, then scatter-added. */
SAD_API int sad_three_interpolate_bwd(int B, int C, int n, int m, const float* grad_out,
                              const int32_t* idx, const float* weight, float* grad_features,
                              sad_stream_t stream);

/* ---- a6  shared point-wise MLP (+ max-pool), fused with the neighbourhood gather ------------
 *
 * Weights travel as pre-packed bf16 "images" (the exact tcgen05 K-major SWIZZLE_128B
 * shared-memory byte layout, so the TMA engine can stream them with plain bulk copies).
 * sad_mlp_weight_image_bytes / sad_mlp_pack_weights are HOST functions (no CUDA call):
 *   W      (cout x cin) fp32 row-major (BN already folded),
 *   perm   kpad int32: perm[k] = column of W feeding packed K index k, -1 = zero column
 *          (layer 1 K order: [feat_cl C0 | feat2_cl C1in | special chunk of 64: dx,dy,dz,
 *          extras..., zeros]; later layers: identity over the previous layer's outputs),
 *   kpad   K rounded up to a multiple of 64,
 *   is_last  0 = hidden layer; final layer of the stack: 1 when the stage pools (S > 1, evaluated
 *          transposed in 128-channel blocks), 2 when S == 1 (plain, rows padded to 32, <= 512).
 * Returns bytes (or -1) / SAD_OK. */
SAD_API long long sad_mlp_weight_image_bytes(int cout, int kpad, int is_last);
SAD_API int sad_mlp_pack_weights(const float* W, int cout, int cin, const int32_t* perm, int kpad,
                                 int is_last, void* out_image_host);

/* One launch = gather + 2..3 layer MLP (bias, ReLU; last ReLU optional) + max over S.
 *   rows           r = (b, j, s), b<B, j<P, s<S;   S power of two <= 128 (1 = no pooling)
 *   feat_cl        (B,N,C0) bf16 channel-last, row picked by idx (C0 % 64 == 0; NULL if C0 == 0)
 *   feat2_cl       (B,P,C1in) bf16 channel-last, row-aligned second source (needs S == 1)
 *   xyz,new_xyz    (B,N,3),(B,P,3) f32: adds (xyz[idx]-new_xyz[j]) (/ radius if normalize_xyz)
 *   idx            (B,P,S) i32, or NULL = identity (needs S == 1 and N == P)
 *   radius_t       (B,P) f32 per-cluster radius (overrides `radius`) or NULL
 *   extra          (B,N,E) f32 scalar features appended after xyz in the special chunk (E <= 13)
 *   w_img, bias    n_layers device pointers (packed images / f32 biases), c_out widths;
 *                  hidden widths % 64 == 0 and <= 256
 *   out_cl_bf16    (B,P,c_last) bf16 channel-last and/or out_cf_f32 (B,c_last,P) f32 (either may be NULL)
 *   tile_counter   one device int32 the CALLER ZEROES before every call (stream-ordered): tiles are then
 *                  handed to the persistent CTAs dynamically; NULL = static round-robin
 * bf16 operands, fp32 accumulate/bias/ReLU/max (tolerance 2e-2 vs the fp32 oracle). */
/* ---- a6, shape-specialised fast path (csrc/mlp_sa.cu): the pooled SA stages of the detector, compiled per shape.
 * All weights pinned in shared memory (CTA pairs / cta_group::2 for the 128-wide stages), feature rows by cp.async
 * straight into the swizzled operand layout, 2-4 tiles in flight in TMEM.  Same math and tolerance as sad_shared_mlp_fwd; layer-1 and hidden biases
 * are rounded to bf16 where they ride on a constant-1 K column.
 *   sad_sa_mlp_query        instance id for (C0 gathered channels, hidden widths h1 == h2, c3 outputs, nsample S,
 *                           E <= 4 scalar features, relative xyz present), or -1: use sad_shared_mlp_fwd.
 *                           prefer: 0 = the library's choice, 1 = a single-CTA instance, 2 = a CTA-pair instance,
 *                           where one exists (tools / tests).
 *   sad_sa_mlp_image_bytes  size of the packed weight image of an instance
 *   sad_sa_mlp_pack         HOST: fp32 W1 (h x cin1), W2 (h x h), W3 (c3 x h), b1, b2 -> image.  perm_feat[C0]: source
 *                           column of gathered K index k (-1 = zero); perm_sp[7]: source columns of dx,dy,dz,e0..e3
 *   sad_pack_xyzw           (B,N,3) xyz (+ optional (B,N) scalar feature) -> (B,N) float4 {x,y,z,f}: optional gathered
 *                           source of sad_sa_mlp_fwd (`xyzw`, needs E <= 1): one 16-byte load per neighbour row
 *   sad_sa_mlp_fwd          launch.  bias3_padded has the instance's full output width (zero padded);
 *                           npoint P must be a power of two; `sched` = 2 device int32 that are ZERO before the first
 *                           launch that uses them (the kernel re-zeroes them when it finishes; launches sharing the
 *                           words must be stream-ordered); tiles_per_cta = scheduling hint (>= 1, never changes
 *                           results): minimum tiles per CTA, i.e. a narrower grid for small stages. */
SAD_API int sad_sa_mlp_query(int C0, int h1, int h2, int c3, int S, int E, int has_xyz, int prefer);
SAD_API int sad_sa_mlp_instance_info(int instance, int* out5); /* {CTAs per MMA group, gathered 64-wide chunks, hidden width, max outputs, nsample} */
SAD_API long long sad_sa_mlp_image_bytes(int instance);
SAD_API int sad_sa_mlp_pack(int instance, const float* W1, int cin1, const int32_t* perm_feat, const int32_t* perm_sp,
                            const float* b1, const float* W2, const float* b2, const float* W3, int c3, void* out_image);
SAD_API int sad_pack_xyzw(int B, int N, const float* xyz, const float* extra1, void* out_xyzw, sad_stream_t stream);
SAD_API int sad_sa_mlp_fwd(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                           const float* new_xyz, const int32_t* idx, float radius, const float* radius_t,
                           int normalize_xyz, const float* extra, int E, const void* w_image, const float* bias3_padded, int c3, void* out_cl_bf16,
                           float* out_cf_f32, int* sched, int tiles_per_cta, sad_stream_t stream);

/* Same stage, DUPLICATE-FREE.  A ball query pads a neighbourhood that has fewer than nsample hits with copies of its first
 * hit, and a max-pool ignores copies: only the leading samples need to go through the MLP.  A small plan kernel turns
 * every point into a run of 1, 2 or 4 slots of 8 or 16 samples (the smallest run after which every remaining sample
 * equals sample 0 -- checked per point, so any idx is handled exactly) and the stage's sibling instance with that nsample
 * runs over the slots, max-combining the slots of a point in its epilogue.  slot_samples: 0 = the smallest
 * slot the library has an instance for (16 today; an 8-sample variant measured slower and is not compiled in).  Results are bit-identical to sad_sa_mlp_fwd.
 * `instance`: the stage's ordinary single-CTA instance with nsample 32 or 64 (else SAD_EUNSUPPORTED); same weight image.
 * workspace: sad_sa_mlp_dedup_workspace_bytes(B, P) bytes, 16-byte aligned, caller-owned scratch.
 * sched: 32 zero-initialised ints that the kernel re-zeroes ([0..1] tile scheduler, [16..18] the plan's counters). */
SAD_API long long sad_sa_mlp_dedup_workspace_bytes(int B, int P);
SAD_API int sad_sa_mlp_dedup_fwd(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                                 const float* new_xyz, const int32_t* idx, float radius, const float* radius_t,
                                 int normalize_xyz, const float* extra, int E, const void* w_image, const float* bias3_padded,
                                 int c3, void* out_cl_bf16, float* out_cf_f32, int* sched, void* workspace,
                                 int slot_samples, int tiles_per_cta, sad_stream_t stream);

/* ---- a6 / a9, shape-specialised POINT-WISE fast path (csrc/mlp_pw.cu): nsample == 1 stages with 256-wide layers.
 *   kind 0  FP module: [three_interpolate(known) (256) | skip (256)] -> 256 -> c_last (<= 256), ReLU everywhere.  The
 *           interpolation is computed inside the kernel from known_cl (B*m,256) bf16, nn_idx (B*n,3) (indices into
 *           the m known points of the same batch element) and nn_w (B*n,3): no three_interpolate launch, no
 *           interpolated tensor.  src_cl (B*n,256) bf16 = the skip features.
 *   kind 1  voting module: seed features (256) -> 256 -> 256 -> 3 + 256 (linear), fused with vote = seed + y:
 *           vote_xyz (B*n,3) = seed_xyz + y[0:3]; out_cf (B,256,n) = seed_cf + y[3:]; out_cl its bf16 twin.
 * Weights stream through shared memory (sad_pw_mlp_pack: fp32 row-major W1 (256 x K0), W2 (256 x 256, kind 1), Wlast
 * (c_last x 256) -> image of sad_pw_mlp_image_bytes(kind) bytes).  n must be a multiple of 128.  bias_last_padded has
 * 256 (kind 0) / 384 (kind 1) entries.  Same math and tolerance as sad_shared_mlp_fwd. */
SAD_API long long sad_pw_mlp_image_bytes(int kind);
SAD_API int sad_pw_mlp_pack(int kind, const float* W1, const float* W2, const float* Wlast, int c_last, void* out_image);
SAD_API int sad_pw_mlp_fwd(int kind, int B, int n, int m, const void* src_cl, const void* known_cl, const int32_t* nn_idx,
                           const float* nn_w, const void* w_image, const float* bias1, const float* bias2,
                           const float* bias_last_padded, int c_last, float* out_cf, void* out_cl, const float* seed_xyz,
                           const float* seed_cf, float* vote_xyz, int tiles_per_cta, sad_stream_t stream);

/* ---- a2 / a5 / a9 backward, deterministic mode (csrc/scatter.cu).  The default *_bwd entry points scatter
 * with fp32 atomics (summation order varies from run to run).  These turn the scatter into a gather with a fixed order:
 *   sad_scatter_plan_build  idx (B,PS) destinations in [0,N) -> order (B,PS), offsets (B,N+1): the source positions of
 *                           every destination, ascending (stable counting sort; out-of-range indices are skipped);
 *   sad_interp_plan_build   the same for three_interpolate's idx (B,n,3) over m destinations, position key t*n + i;
 *   sad_scatter_add_det     grad_features (B,C,N) = per destination the sequential fp32 sum of grad_out (B,C,PS) over
 *                           its segment (grouping: PS = npoint*nsample; gather: PS = npoint); no memset needed;
 *   sad_three_interpolate_bwd_det  ... of grad_out[b,c,i] * weight[b,i,t].
 * Results are reproducible bit for bit and equal the oracle's index-order accumulation. */
SAD_API int sad_scatter_plan_build(int B, int N, long long PS, const int32_t* idx, int32_t* order, int32_t* offsets,
                                   sad_stream_t stream);
SAD_API int sad_interp_plan_build(int B, int n, int m, const int32_t* idx, int32_t* order, int32_t* offsets,
                                  sad_stream_t stream);
SAD_API int sad_scatter_add_det(int B, int C, int N, long long PS, const float* grad_out, const int32_t* order,
                                const int32_t* offsets, float* grad_features, sad_stream_t stream);
SAD_API int sad_three_interpolate_bwd_det(int B, int C, int n, int m, const float* grad_out, const float* weight,
                                          const int32_t* order, const int32_t* offsets, float* grad_features,
                                          sad_stream_t stream);

/* ---- a6, tf32 mode (csrc/mlp_tf32.cu): the same fused stage with fp32 channel-last activations and kind::tf32 MMAs
 * (north_star "tcgen05 bf16/tf32 GEMM"); 1e-3-class agreement with the fp32 oracle instead of the bf16 path's 2e-2.
 * Layer-1 operand, K order [interp | feat | special]:
 *   interp   CI channels = three_interpolate(known_cl (B,m,CI), nn_idx (B*P,3), nn_w (B*P,3)) computed in the kernel
 *            (nsample == 1 only), or CI == 0;
 *   feat     CF channels of feat_cl (B,N,CF), row idx[b,p,s] (idx (B,P,S)), or row p itself when idx == NULL
 *            (nsample == 1, N == P), or CF == 0;
 *   special  when xyz != NULL: [(xyz[idx]-new_xyz)/r, extra[idx][0..E-1], 0..] as one 8-wide K step (E <= 4); r = radius
 *            or radius_t[b,p] when normalize_xyz, else 1.
 * CI, CF multiples of 4; each part is zero-padded to 32-channel chunks.  2-3 layers, hidden widths % 32 == 0 and
 * <= 256, last width <= 512, nsample a power of two <= 128.  ReLU after every layer except (last_relu == 0) the last;
 * max over nsample.  out_cf (B,c_last,P) f32 and / or out_cl (B,P,c_last) f32.
 * Weights: per layer one image from sad_mlp_tf32_pack(W (c_out x kc*32) fp32 row-major, input channels in operand
 * order, zero-padded per chunk; kc = that layer's 32-channel K chunks; last = 1 for the last layer) of
 * sad_mlp_tf32_image_bytes(c_out, kc, last) bytes, device resident.  bias[l]: c_out[l] floats (last layer: padded
 * with zeros to a multiple of 128). */
SAD_API long long sad_mlp_tf32_image_bytes(int c_out, int kc, int last);
SAD_API int sad_mlp_tf32_pack(const float* W, int c_out, int kc, int last, void* out);
SAD_API int sad_mlp_tf32_fwd(int B, int N, int P, int S, const float* known_cl, int m, int CI, const int32_t* nn_idx,
                             const float* nn_w, const float* feat_cl, int CF, const int32_t* idx, const float* xyz,
                             const float* new_xyz, float radius, const float* radius_t, int normalize_xyz,
                             const float* extra, int E, int n_layers, const void* const* w_img, const float* const* bias,
                             const int* c_out, int last_relu, float* out_cf, float* out_cl, sad_stream_t stream);

/* Scheduling options of one sad_shared_mlp_fwd launch (never change results); NULL = defaults.
 *   tiles_per_cta  at least this many 128-row tiles per CTA, i.e. a narrower grid for the small stages.  1 (default) =
 *                  one CTA per SM whenever there are that many tiles: shortest time for one launch.  A pipelined
 *                  caller with other streams to fill the SMs wants ~6 (fewer per-CTA prologues, less SM-time).
 *   super_tiles    0 = automatic, 1 = never, 2 = always (when the shape allows) two tiles per context and phase. */
typedef struct sad_mlp_opts {
  int tiles_per_cta;
  int super_tiles;
} sad_mlp_opts;
SAD_API int sad_shared_mlp_fwd(int B, int N, int P, int S, const void* feat_cl, int C0,
                               const void* feat2_cl, int C1in, const float* xyz, const float* new_xyz,
                               const int32_t* idx, float radius, const float* radius_t,
                               int normalize_xyz, const float* extra, int E, int n_layers,
                               const void* const* w_img, const float* const* bias, const int* c_out,
                               int last_relu, void* out_cl_bf16, float* out_cf_f32, int* tile_counter, const sad_mlp_opts* opts,
                               sad_stream_t stream);

/* a9 on the internal layout: features (B,m,C) bf16 channel-last -> out (B,n,C) bf16 (C % 8 == 0). */
SAD_API int sad_three_interpolate_cl_fwd(int B, int C, int m, int n, const void* feat_cl_bf16,
                                         const int32_t* idx, const float* weight, void* out_cl_bf16,
                                         sad_stream_t stream);

/* Layout bridge for the drop-in surface: (B,C,N) f32 channel-first -> (B,N,C) bf16 channel-last. */
SAD_API int sad_cf_to_cl_bf16(int B, int C, int N, const float* in_cf, void* out_cl_bf16,
                              sad_stream_t stream);

/* Number of kernels this library has launched in this process (all threads, monotonic). */
SAD_API unsigned long long sad_launch_count(void);

/* ---- pipelined executor, host side (engine.PipelinedHotPath) ------------------------------------------------------
 * One batch = [stream waits for `wait_event`] -> n_in async copies (host-pinned or device sources; the direction is
 * inferred from the pointers) -> launch of the slot's instantiated CUDA graph -> n_out async copies of the results
 * -> `done_event` recorded, all queued on `stream` by ONE call (no Python between the driver calls: the per-batch
 * host cost is what bounds how fast a pipeline of `slots` batches fills).  Handles are the CUDA runtime's
 * (cudaGraphExec_t / cudaEvent_t); NULL wait_event / done_event are skipped.  Launches no kernel of its own. */
#define SAD_SUBMIT_MAX_COPIES 4
typedef struct {
  void* graph_exec;                            /* cudaGraphExec_t */
  void* wait_event;                            /* cudaEvent_t or NULL */
  void* done_event;                            /* cudaEvent_t or NULL */
  int n_in, n_out;
  void* in_dst[SAD_SUBMIT_MAX_COPIES];
  const void* in_src[SAD_SUBMIT_MAX_COPIES];
  size_t in_bytes[SAD_SUBMIT_MAX_COPIES];
  void* out_dst[SAD_SUBMIT_MAX_COPIES];
  const void* out_src[SAD_SUBMIT_MAX_COPIES];
  size_t out_bytes[SAD_SUBMIT_MAX_COPIES];
} sad_submit_desc;
SAD_API int sad_engine_submit(const sad_submit_desc* desc, sad_stream_t stream);

/* a1 with an explicit thread-block-cluster size (tests / tools: 1,2,4,8,16; 0 = the built-in heuristic, i.e.
 * sad_furthest_point_sample_fwd).  Results never depend on it. */
SAD_API int sad_furthest_point_sample_cs_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx, int cluster_size,
                                             sad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAD_OPS_H_ */
