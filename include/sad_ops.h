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
 *    global mutable state besides a thread-local error string;
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

/* a9  three_interpolate: out[b,c,i] = ((w0*f[i0]) + (w1*f[i1])) + (w2*f[i2]).
 * features (B,C,m) f32, idx (B,n,3) i32, weight (B,n,3) f32 -> out (B,C,n) f32. */
SAD_API int sad_three_interpolate_fwd(int B, int C, int m, int n, const float* features,
                              const int32_t* idx, const float* weight, float* out,
                              sad_stream_t stream);
/* a9 backward: grad_features (B,C,m) is ZEROED by the call, then scatter-added. */
SAD_API int sad_three_interpolate_bwd(int B, int C, int n, int m, const float* grad_out,
                              const int32_t* idx, const float* weight, float* grad_features,
                              sad_stream_t stream);

/* Number of kernels this library has launched in this process (all threads, monotonic). */
SAD_API unsigned long long sad_launch_count(void);

/* Test / benchmark hook: force the FPS thread-block-cluster size for subsequent calls on
 * this thread (1,2,4,8,16; 0 = built-in heuristic).  Results never depend on it. */
SAD_API void sad_fps_force_cluster_size(int cluster_size);

#ifdef __cplusplus
}
#endif
#endif /* SAD_OPS_H_ */
