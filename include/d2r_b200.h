/*
 * d2r_b200.h -- C ABI of libd2r_b200.so: the B200-native replacement for the native code on
 * Dream2Real's imagination-and-scoring hot path.
 *
 * What it replaces (paths relative to /root/reference):
 *   - the pybind11 module `pyngp` as far as the path uses it
 *       reconstruction/instant-ngp/src/python_api.cu:123-201 (Testbed::render = render_to_cpu),
 *       :382-566 (load_snapshot, set_nerf_camera_matrix, set_camera_to_training_view, ...);
 *   - the NumPy/cv2 compositing in reconstruction/combined_rendering.py:117-155;
 *   - the PIL/HF preprocessing + CLIPModel.forward + logit arithmetic in clip_scoring.py:145-203.
 *
 * Conventions:
 *   - every function returns 0 on success, a negative d2r_status otherwise; the message is
 *     available through d2r_last_error() (thread-local).  No C++ exception crosses the boundary.
 *   - plain pointers and sizes only.  Pointers named *_dev are DEVICE pointers owned by the
 *     caller (typically torch tensors); pointers named *_host are host pointers.
 *   - `stream` is a cudaStream_t passed as void*.  The per-candidate calls (d2r_render*, d2r_clip_*,
 *     d2r_score) only enqueue work: they never wait for the device.  The candidate cameras are a HOST
 *     array (the pose grid is built on the host): the library plans every candidate's screen rectangle
 *     from it on the CPU and uploads cameras + plan with one pinned asynchronous copy, so no launch
 *     needs a read-back.  Exceptions, all outside the steady state: a call whose scratch must grow
 *     (first launch of a size class) goes through cudaMalloc/cudaFree; a host that runs more than 8
 *     launches ahead of the device waits for the oldest upload slot; the set-up calls
 *     (d2r_model_load, d2r_view_prepare, d2r_clip_load) and the *_get_* read-backs synchronise.
 *   - handles are not re-entrant; different handles / streams / devices may be used concurrently
 *     (launch scratch is kept per stream).  The caller's current device is never left changed.
 */
#ifndef D2R_B200_H
#define D2R_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    D2R_OK = 0,
    D2R_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    D2R_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed      */
    D2R_ERR_NOMEM = -3
} d2r_status;

typedef struct d2r_model d2r_model;   /* one Instant-NGP NeRF (hash grid + 2 MLPs + occupancy bitfield) */
typedef struct d2r_view d2r_view;     /* one (training-view intrinsics, W x H) ray table                */
typedef struct d2r_clip d2r_clip;     /* CLIP vision tower weights + workspaces                         */

/* ---- model ---------------------------------------------------------------------------------- */
/* Mirrors what Testbed::load_snapshot + reset_network leave on the device
 * (reconstruction/instant-ngp/src/testbed.cu:4757-4871, :3613-3673; testbed_nerf.cu:2195-2231). */
typedef struct {
    int32_t n_levels;             /* 8   (configs/nerf/base.json:23-29)           */
    int32_t n_features_per_level; /* 4                                             */
    int32_t log2_hashmap_size;    /* 19                                            */
    int32_t base_resolution;      /* 16                                            */
    float per_level_scale;        /* exp(log(2048*aabb_scale/16)/7), testbed.cu:3668 */
    int32_t max_cascade;          /* log2(aabb_scale), testbed_nerf.cu:2221-2224    */
    float aabb_min[3], aabb_max[3];               /* m_aabb (training box, warp_position) */
    float render_aabb_min[3], render_aabb_max[3]; /* m_render_aabb                        */
    float render_aabb_to_local[9];                /* row-major 3x3                        */
    float cone_angle_constant;    /* 1/256 if aabb_scale > 1 else 0, testbed_nerf.cu:2228 */
    float min_transmittance;      /* nerf.render_min_transmittance, testbed.h:766 (0.01)  */
    float depth_scale;            /* 1 / dataset.scale, testbed_nerf.cu:1888              */
} d2r_model_cfg;

/* params_f16_host: the snapshot's `params_binary` blob as-is: fp16
 *   [density MLP 64x32,16x64 | rgb MLP 64x32,64x64,16x64 | hash grid]  (nerf_network.h:356-372).
 * density_grid_f32_host: (max_cascade+1)*128^3 floats (snapshot `density_grid_binary` widened),
 *   converted here to the occupancy bitfield exactly like testbed_nerf.cu:284-331,2355-2373.   */
int d2r_model_load(const void* params_f16_host, size_t n_params,
                   const float* density_grid_f32_host, size_t n_grid_cells,
                   const d2r_model_cfg* cfg, int device, d2r_model** out);
void d2r_model_free(d2r_model* m);
/* nerf.render_min_transmittance may be changed after loading (combined_rendering.py:49). */
int d2r_model_set_min_transmittance(d2r_model* m, float min_transmittance);
/* Copy the 8-cascade occupancy bitfield (128^3/8*8 bytes) to the host -- parity tests. */
int d2r_model_get_bitfield(const d2r_model* m, uint8_t* bitfield_host, size_t n_bytes);
/* Tight box (in NGP coordinates) around every occupied cell of every cascade, clipped to the
 * render aabb: rays that miss it can never take a sample. out6 = min xyz, max xyz.           */
int d2r_model_get_occupied_aabb(const d2r_model* m, float* out6_host);

/* ---- view ----------------------------------------------------------------------------------- */
/* State after set_camera_to_training_view(v) for a W x H render (testbed.cu:453-468,4065-4072):
 * focal = metadata.focal_length / resolution[fov_axis] * {W,H}[fov_axis] * zoom; screen_center =
 * principal point fraction; lens = the view's OpenCV (k1,k2,p1,p2) or perspective.
 * Builds the per-pixel camera-plane direction table (Newton lens undistortion,
 * common_device.cuh:289-333,393-431) once; every candidate pose then only rotates it.         */
typedef struct {
    int32_t width, height;
    float focal[2];
    float screen_center[2];
    int32_t lens_mode;       /* 0 = perspective, 1 = OpenCV */
    float lens_params[4];
} d2r_camera;
int d2r_view_prepare(const d2r_camera* cam, int device, d2r_view** out);
void d2r_view_free(d2r_view* v);
int d2r_view_get_dirs(const d2r_view* v, float* dirs_xy_host /* [H*W*2] */);

/* ---- render --------------------------------------------------------------------------------- */
#define D2R_MODE_SHADE 1
#define D2R_MODE_DEPTH 2
/* K renders of one model from K cameras = K x Testbed::render(W,H,1,linear=True) per mode
 * (python_api.cu:123-201).  cams_ngp_host: [K,3,4] float, ALREADY in NGP convention
 * (nerf_matrix_to_ngp applied, nerf_loader.h:101-121).  background_rgba: Testbed.background_color.
 * rgba_out_dev / depth_out_dev: [K,H,W,4] float32 linear premultiplied (NULL = skip that mode);
 * both modes come out of ONE march.  n_samples_out_dev (optional, uint64 device): sample counter. */
int d2r_render(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K,
               const float background_rgba[4], float* rgba_out_dev, float* depth_out_dev,
               unsigned long long* n_samples_out_dev, void* stream);

/* d2r_render plus the reference's Cost render mode (testbed_nerf.cu:1322-1326: payload.n_steps): cost_out_dev [K,H,W]
 * float32 = the step count of every ray the reference keeps (final alpha > 0.001), 0 elsewhere; NULL = skip.
 * The parity tests use it to tell single-sample flips at occupancy-cell boundaries from arithmetic differences. */
int d2r_render_ex(const d2r_model* m, const d2r_view* v, const float* cams_ngp_host, int K,
                  const float background_rgba[4], float* rgba_out_dev, float* depth_out_dev,
                  float* cost_out_dev, unsigned long long* n_samples_out_dev, void* stream);

/* Fused candidate render + depth-test composite + colour post-process =
 * reconstruction/combined_rendering.py:117-155 for K candidate poses of the movable object:
 *   fg Shade + fg Depth (one march), fg_d<0.05 -> 100, bg_d<0.05 -> 100, fg where fg_d < bg_d,
 *   rgb/a, linear_to_srgb (scripts/common.py:142-144), u8 = clip*255+0.5, alpha_u8 < 130 -> 0.
 * bg_rgba_dev [H,W,4] / bg_depth_dev [H,W]: the cached background render and depth map.
 * rgb_u8_out_dev: [K,H,W,3] uint8.                                                             */
int d2r_render_composite(const d2r_model* fg, const d2r_view* v, const float* cams_ngp_host, int K,
                         const float fg_background_rgba[4], const float* bg_rgba_dev,
                         const float* bg_depth_dev, uint8_t* rgb_u8_out_dev,
                         unsigned long long* n_samples_out_dev, void* stream);

/* Same, and additionally reports what the preprocessing can exploit: rects_out_dev [K,4] int32 = the screen
 * rectangle (x0, y0, x1, y1 inclusive; x1 < x0 = none) outside which candidate k's frame equals the composited
 * background, and bg_u8_out_dev [H,W,3] = that background frame (the composite of an absent object).  Both optional. */
int d2r_render_composite_ex(const d2r_model* fg, const d2r_view* v, const float* cams_ngp_host, int K,
                            const float fg_background_rgba[4], const float* bg_rgba_dev,
                            const float* bg_depth_dev, uint8_t* rgb_u8_out_dev, int* rects_out_dev,
                            uint8_t* bg_u8_out_dev, unsigned long long* n_samples_out_dev, void* stream);

/* ---- CLIP preprocessing ----------------------------------------------------------------------
 * np.rot90(k=1, axes=(1,2)) (clip_scoring.py:145) then transformers CLIPImageProcessor (PIL
 * backend): resize shortest side -> R with PIL BICUBIC (antialiased, u8 fixed-point two-pass),
 * centre crop R, /255, (x-mean)/std.  Output is written patch-major for the ViT patch-embed GEMM:
 * patches_out_dev [K * (R/P)^2, 3*P*P] fp16, row = k*(R/P)^2 + py*(R/P)+px, col = c*P*P+iy*P+ix.
 * pixels_f32_out_dev (optional): [K,3,R,R] float32 "pixel_values" for parity checks.            */
int d2r_clip_preprocess(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P,
                        const float mean[3], const float std[3], void* patches_out_dev,
                        float* pixels_f32_out_dev, void* stream);

/* d2r_clip_preprocess for frames that equal a common background outside per-candidate rectangles (the output of
 * d2r_render_composite_ex): the background is resized once, per candidate only the outputs whose filter windows
 * touch its rectangle are recomputed (same integer arithmetic on the same bytes => bit-identical to
 * d2r_clip_preprocess on the full frames).  bg_u8_dev [H,W,3], rects_dev [K,4] int32 (device).              */
int d2r_clip_preprocess_delta(const uint8_t* rgb_u8_dev, int K, int H, int W, int rot90, int R, int P,
                              const float mean[3], const float std_[3], const uint8_t* bg_u8_dev,
                              const int* rects_dev, void* patches_out_dev, void* stream);

/* ---- CLIP vision tower ----------------------------------------------------------------------- */
typedef struct {
    int32_t image_size;    /* 224 | 336 */
    int32_t patch_size;    /* 32  | 14  */
    int32_t hidden;        /* 768 | 1024 */
    int32_t heads;         /* 12  | 16  */
    int32_t layers;        /* 12  | 24  */
    int32_t mlp;           /* 3072 | 4096 */
    int32_t proj;          /* 512 | 768 */
    float ln_eps;          /* 1e-5 */
    int32_t max_batch;     /* images per forward call (workspace sizing) */
} d2r_clip_cfg;

/* Weight order (all float32 host pointers, HF transformers CLIPVisionModelWithProjection names):
 *  0 embeddings.patch_embedding.weight [hidden,3,P,P]   1 embeddings.class_embedding [hidden]
 *  2 embeddings.position_embedding.weight [T,hidden]    3,4 pre_layrnorm.{weight,bias}
 *  then per layer l (16 tensors): ln1.{w,b}, q.{w,b}, k.{w,b}, v.{w,b}, out_proj.{w,b},
 *                                 ln2.{w,b}, fc1.{w,b}, fc2.{w,b}
 *  then post_layernorm.{weight,bias}, visual_projection.weight [proj,hidden].                  */
int d2r_clip_load(const d2r_clip_cfg* cfg, const float* const* weights_host, int n_weights,
                  int device, d2r_clip** out);
void d2r_clip_free(d2r_clip* c);
/* patches_dev: output of d2r_clip_preprocess for K <= max_batch images.
 * embeds_out_dev: [K, proj] float32, L2-normalised image embeddings (CLIPModel.forward).       */
int d2r_clip_encode(d2r_clip* c, const void* patches_dev, int K, float* embeds_out_dev, void* stream);

/* logits_per_image = exp(logit_scale) * img . txt^T; score = mean(goal logits)/mean(norm logits)
 * (clip_scoring.py:180-203).  n_goal = number of leading goal captions (1, or #templates);
 * C == n_goal -> score is the mean goal logit.  logits_out_dev optional [K,C].                 */
int d2r_score(const float* img_embeds_dev, const float* txt_embeds_dev, int K, int C, int D,
              float logit_scale_exp, int n_goal, float* scores_out_dev, float* logits_out_dev,
              void* stream);

/* ---- pre-render physics filter ------------------------------------------------
 * The per-pose part of `unsupcol_check` (vision_3d/physics_utils.py:305-372) for all N poses of the grid in one launch: in
 * collision -> invalid; moved `unsup_thresh` down along gravity the object must touch a support unless the pose lies below
 * the table plane; lowered, it must still touch one when pushed `p_dist` along +-x and +-y.  The collision primitive is an
 * occupancy overlap: fg_points_world_dev [n_pts,4] float = occupied cell centres of the movable object's density grid (world
 * frame, initial pose; w unused), moved by rel_3x4_dev [N,12] = pose . init_pose^-1, against the occupancy bitfield of `bg`.
 * pose_z_dev [N] = z of the candidate pose; valid_in_dev / valid_out_dev [N] uint8.                                   */
typedef struct {
    float dataset_scale;       /* world -> NGP: p * scale + offset, xyz <- yzx (nerf_loader.h:148-151) */
    float dataset_offset[3];
    float scene_centre_z;      /* task_model.scene_model.scene_centre[2] */
    float unsup_thresh;        /* 0.02 */
    float p_dist;              /* 0.04 */
    int32_t stability_check;
} d2r_phys_cfg;
int d2r_phys_check(const d2r_model* bg, const float* fg_points_world_dev, int n_pts, const float* rel_3x4_dev,
                   const float* pose_z_dev, const uint8_t* valid_in_dev, int N, const d2r_phys_cfg* cfg,
                   uint8_t* valid_out_dev, void* stream);

/* ---- building block exported for tests: the tcgen05 GEMM every ViT contraction runs on ----------
 * out = A[M,K] . B[N,K]^T (+bias[N]); A, B fp16 K-major device pointers with leading dimensions
 * lda/ldb (elements); K % 64 == 0, N % 64 == 0.  mode: 0 fp16 out, 1 fp16 quick-GELU out,
 * 2 fp32 in-place residual add, 3 fp32 out.                                                     */
int d2r_gemm_f16(const void* a_dev, int lda, const void* b_dev, int ldb, int M, int N, int K,
                 const float* bias_dev, int mode, void* out_dev, int ldo, void* stream);

/* ---- measurement hooks (bench.py roofline): CUDA events around the march kernel alone, recorded on
 * the launching stream, plus device counters of network samples and 16x8 ray tiles processed.   */
int d2r_profile_enable(int device, int on);   /* also resets the accumulated state */
int d2r_profile_read(int device, float* march_ms_total, int* n_launches,
                     unsigned long long* n_samples, unsigned long long* n_tiles);
/* The raw counters of the profiled launches (synchronises): [0] samples, [1] rays, then the role statistics of k_march_ws
 * in SM cycles summed over warps -- [2] gather total, [3] gather waiting for its MLP round, [4] occupancy walk, [5] hash-grid
 * gather, [6] slot refill, [7] epilogue total, [8] epilogue waiting for an MMA, [9] epilogue items, [10] MMA thread total,
 * [11] MMA thread issuing, [12] gather rounds.  tools/ws_stats.py prints them.                                          */
int d2r_profile_read_stats(int device, unsigned long long* out16_host);

/* ---- misc ------------------------------------------------------------------------------------ */
const char* d2r_last_error(void);
/* number of kernels this library has launched on the calling thread since the last reset */
unsigned long long d2r_launch_count(int reset);
const char* d2r_version(void);

#ifdef __cplusplus
}
#endif
#endif /* D2R_B200_H */
