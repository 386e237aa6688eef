/* include/vrb200.h -- C ABI of libvrb200.so, the B200-native (sm_100a) replacement of the GPU hot path of
 * lquatrin/cpp_volume_rendering.  Plain pointers and sizes only; no C++, GL or torch types cross this boundary.
 *
 * What each entry point replaces (file:line relative to the reference tree):
 *   - the GL texture / image objects and glDispatchCompute calls made by the BaseVolumeRenderer subclasses
 *     (cppvolrend/volrenderbase.h:25-98) and by their helpers;
 *   - the CPU pre-passes those subclasses run in Init().
 * The C++ host mirror under cpp_volume_rendering_b200/host/ (same class / method names as the reference)
 * calls ONLY these functions; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (VRB_OK) or a VRB_ERR_* code; vrb_last_error() gives the message of the last
 *     failure on the calling thread.  Nothing ever calls exit() (the reference does: libs/gl_utils/utils.cpp:11-30).
 *   - a vrb_ctx is bound to one CUDA device, is not thread-safe, and owns all device memory it creates.
 *   - host arrays passed in are borrowed for the duration of the call only.
 *   - calls are ordered on the context's CUDA stream; functions that write to HOST memory synchronise it.
 *   - images are W x H, row 0 = bottom row (GL image coordinates, storePos = gl_GlobalInvocationID.xy),
 *     premultiplied RGBA, stored on the device as RGBA16F like the reference's output texture
 *     (libs/vis_utils/renderoutputframe.cpp:64-87).
 */
#ifndef VRB200_H
#define VRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRB_OK              0
#define VRB_ERR_INVALID     1   /* bad argument */
#define VRB_ERR_CUDA        2   /* CUDA runtime error (message has the cudaError string) */
#define VRB_ERR_STATE       3   /* call order violated (e.g. render before volume upload) */
#define VRB_ERR_UNSUPPORTED 4

typedef struct vrb_ctx vrb_ctx;

/* ---- uniforms shared by every marcher ------------------------------------------------------------------ */

/* CameraEye, u_CameraLookAt / ViewMatrix (glm::lookAt, column major), u_TanCameraFovY / fov_y_tangent,
 * u_CameraAspectRatio / aspect_ratio -- rc1prenderer.cpp:91-101, ebsrenderer.cpp:206-219. */
typedef struct vrb_camera {
  float eye[3];
  float lookat[16];
  float tan_fovy;
  float aspect;
} vrb_camera;

/* Blinn-Phong + light uniforms (renderingparameters.cpp:17-32,130-165; lightsourcelist.cpp:21-37). */
typedef struct vrb_lighting {
  float ka, kd, ks, shininess;         /* Kambient, Kdiffuse, Kspecular, Nshininess */
  float ispecular[3];                  /* Ispecular */
  float light_pos[3];                  /* WorldLightingPos / LightSourcePosition */
  float light_forward[3];              /* LightCamForward */
  float light_up[3];                   /* LightCamUp */
  float light_right[3];                /* LightCamRight */
  float spot_angle_deg;                /* SpotLightMaxAngle */
  int   apply_phong;                   /* ApplyPhongShading / ApplyGradientPhongShading: 1 = gradient Blinn-Phong branch of
                                          ShadeSample (needs vrb_gradient_build); the host passes
                                          (m_apply_gradient_shading && GetCurrentGradientTexture()) ? 1 : 0 */
} vrb_lighting;

/* Sort-first image partition: the image is cut into tile_w x tile_h tiles, numbered row
 * major; this context renders the tiles with (tile_index % nranks) == rank and leaves the others untouched (0).
 * nranks <= 1 renders everything. */
typedef struct vrb_partition {
  int rank, nranks;
  int tile_w, tile_h;
} vrb_partition;

/* ---- context --------------------------------------------------------------------------------------------- */
int  vrb_ctx_create(int device, vrb_ctx** out);
int  vrb_ctx_destroy(vrb_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t passed as void*; NULL = the context's own stream). */
int  vrb_ctx_set_stream(vrb_ctx* ctx, void* cuda_stream);
int  vrb_ctx_synchronize(vrb_ctx* ctx);
int  vrb_ctx_set_partition(vrb_ctx* ctx, const vrb_partition* part);
/* Texture filtering of the marchers (the reference's GL_LINEAR / GL_LINEAR_MIPMAP_LINEAR sampler state,
 * libs/volvis_utils/utils.cpp:20-56, gl_utils/texture3d.cpp):
 *   VRB_FILTER_EXACT    blends in fp32 in software, bit-reproducible against oracle/ (loop counts included);
 *   VRB_FILTER_HARDWARE uses the GPU's texture units, as the reference's GL path does (fixed-point blend weights):
 *                       same image within the parity tolerance (2/255, PSNR >= 50 dB), not bit-reproducible.
 * The SAT queries of rc1pextbsd always use exact texels (fp32 prefix sums do not survive fixed-point weights).
 * Default: VRB_FILTER_EXACT, or the VRB_FILTER environment variable ("exact" / "hardware") at context creation. */
#define VRB_FILTER_EXACT    0
#define VRB_FILTER_HARDWARE 1
int  vrb_ctx_set_filter(vrb_ctx* ctx, int mode);
int  vrb_ctx_get_filter(const vrb_ctx* ctx);
const char* vrb_last_error(void);
const char* vrb_version(void);
/* Number of kernels this library launched on the context since creation (bench.py's gpu_launches). */
uint64_t vrb_launch_count(const vrb_ctx* ctx);
/* Loop iterations ("primary samples", ray_marching_1p.comp:124 body) executed by the last render call made with
 * count_samples != 0. */
uint64_t vrb_last_sample_count(const vrb_ctx* ctx);
/* Secondary work items of that same call: SAT box queries (rc1pextbsd), cone taps (rc1pdosct / rc1pvctsg),
 * secondary-ray steps (rc1pcrtgt). */
uint64_t vrb_last_aux_count(const vrb_ctx* ctx);
/* Device time (CUDA events, ms) of the kernels of the last pre-pass build (the three SAT scan passes), scratch
 * allocation excluded; and the SAT layout the marcher samples (1 linear loads, 2/4 packed, 8 texture-gather atlas). */
float vrb_last_prepass_ms(const vrb_ctx* ctx);
/* Per-kernel timing for the roofline report (no reference counterpart: the reference times whole frames with
 * glFinish, renderingmanager.cpp:645-718).  on != 0: every render call records CUDA events on the context's stream
 * around its DOMINANT kernel (the marcher; for the deferred lit renderers the shading kernel).  vrb_last_kernel_ms
 * waits for that kernel of the last render call and returns its duration and name (static string). */
int   vrb_ctx_set_kernel_timing(vrb_ctx* ctx, int on);
int   vrb_last_kernel_ms(vrb_ctx* ctx, float* ms, const char** kernel_name);
int   vrb_sat_layout(const vrb_ctx* ctx);
/* Measured rooflines (GB/s): L1-resident 128-bit loads on every SM; device-to-device copy (read + write bytes). */
int  vrb_measure_l1_bandwidth(vrb_ctx* ctx, double* gb_per_s);
int  vrb_measure_hbm_bandwidth(vrb_ctx* ctx, double* gb_per_s);
/* Texture-pipe ceiling of the SAT box queries: lane-level tex2Dgather operations per second (1e9/s, 16 B each) on a
 * cache-resident R32F array with the 32 lanes of a warp on an 8x4 patch. */
int  vrb_measure_gather_rate(vrb_ctx* ctx, double* ggathers_per_s);
/* Texture-pipe ceiling of the hardware-filter marchers (VRB_FILTER_HARDWARE): lane-level trilinear tex3D fetches per
 * second (1e9/s) on a cache-resident R16F 3-D array, 32 lanes on neighbouring texels at fractional positions. */
int  vrb_measure_tex3d_rate(vrb_ctx* ctx, double* gfetches_per_s);
/* Load-pipe ceiling of the exact-filter marchers: lane-level 16-bit loads per second (1e9/s) on an L1-resident fp16 brick,
 * eight per software trilinear footprint, 32 lanes on neighbouring texels. */
int  vrb_measure_ldg16_rate(vrb_ctx* ctx, double* gloads_per_s);

/* ---- inputs ---------------------------------------------------------------------------------------------- */
/* Replaces vis::GenerateRTexture (libs/volvis_utils/utils.cpp:20-56): voxels x-fastest, u8 (bytes_per_voxel 1)
 * or little-endian u16 (2); voxel scale as StructuredGridVolume::GetScale().  The device copy holds the same
 * values an R16F texture would: half(float(double(v)/255.0)) resp. /65535.0. */
int  vrb_volume_upload(vrb_ctx* ctx, const void* voxels, int w, int h, int d, int bytes_per_voxel,
                       const float scale[3]);
/* Same, but the voxel array already lives in device memory (sort-last bricks generated on the GPU). */
int  vrb_volume_upload_device(vrb_ctx* ctx, const void* dev_voxels, int w, int h, int d, int bytes_per_voxel,
                              const float scale[3]);
/* Replaces TransferFunction1D::GenerateTexture_1D_RGBt and _RGBA (transferfunction1d.cpp:89-118,58-87):
 * the two GL_FLOAT client arrays (n x RGBA; .a = extinction resp. opacity); rounded to RGBA16F on upload. */
int  vrb_tf_upload(vrb_ctx* ctx, const float* rgbt, const float* rgba, int n);

/* ---- output frame (vis::RenderFrameToScreen, renderoutputframe.cpp:64-87,187-202) --------------------------- */
int  vrb_frame_resize(vrb_ctx* ctx, int width, int height);          /* UpdateScreenResolution */
int  vrb_frame_clear(vrb_ctx* ctx);                                  /* ClearTexture */
/* glGetTexImage(GL_RGBA, GL_FLOAT) of the output texture (renderingmanager.cpp:637-640): W*H*4 floats. */
int  vrb_frame_read_rgba32f(vrb_ctx* ctx, float* host_out);
/* Pipelined form of the same read (the reference reads back synchronously, renderingmanager.cpp:637-640 and twice per
 * iteration in crtgtrenderer.cpp:272-325): the fp16 -> fp32 conversion is queued behind the render on the context's
 * stream and the device -> host copy on a separate copy stream, so the copy of frame i overlaps the render of frame
 * i+1.  host_out must stay valid (and should be page-locked) until vrb_frame_read_wait reports it complete.  At most
 * two reads are in flight.  vrb_frame_read_wait(ctx, 1) returns once all but the most recent read have landed,
 * vrb_frame_read_wait(ctx, 0) once all have. */
int  vrb_frame_read_rgba32f_async(vrb_ctx* ctx, float* host_out);
int  vrb_frame_read_wait(vrb_ctx* ctx, int max_in_flight);
/* Sort-first without a frame reduce.  vrb_frame_set_target redirects this context's pixel stores (and its read-backs) to
 * another RGBA16F buffer of the frame's size: a buffer of this context (vrb_frame_extra) or the display rank's buffer
 * mapped with vrb_ipc_import, so that the image is assembled by the marchers' own stores over NVLink and the only
 * collective left is a barrier.  While a target is set the buffer is NOT cleared per render (other contexts write into
 * it too) and every owned pixel is stored, zeros where the ray misses.  NULL restores the context's own frame.
 * vrb_frame_extra returns one of two context-owned buffers (double buffering: frame i+1 must not land in the buffer
 * frame i is still being read from); they are freed by vrb_frame_resize. */
int  vrb_frame_set_target(vrb_ctx* ctx, void* dev_rgba16f);
int  vrb_frame_extra(vrb_ctx* ctx, int index, void** dev_rgba16f);
/* Device pointer of the RGBA16F image (the analogue of GetScreenTextureID(), volrenderbase.h:73-75). */
int  vrb_frame_device_ptr(vrb_ctx* ctx, void** dev_rgba16f, int* width, int* height);

/* ---- rc1pass: single-pass ray casting (rc1pass/ray_marching_1p.comp:85-179) --------------------------------- */
typedef struct vrb_rc1pass_params {
  float step_size;          /* StepSize (rc1prenderer.cpp:62-63) */
  int   count_samples;      /* != 0: also count executed loop iterations (vrb_last_sample_count) */
  int   skip_empty;         /* != 0: result-preserving empty-space skipping; 0 = as the reference */
} vrb_rc1pass_params;
int  vrb_rc1pass_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_rc1pass_params* p);

/* Gradient texture of the current volume.  Replaces DataManager::GenerateStructuredGradientTexture (libs/volvis_utils/
 * datamanager.cpp:332-352) and its three generators: vis::GenerateSobelFeldmanGradientTexture (utils.cpp:287-350),
 * vis::GenerateGradientTexture with default arguments (utils.cpp:146-284) and the compute-shader Sobel
 * (datamanager.cpp:623-717, sobelfeldman_generator.comp).  Texels are RGB16F like the reference's; VRB_GRADIENT_NONE
 * frees the texture (DeleteGradientData).  A new vrb_volume_upload drops it. */
enum { VRB_GRADIENT_NONE = 0, VRB_GRADIENT_SOBEL_FELDMAN = 1, VRB_GRADIENT_FINITE_DIFFERENCES = 2, VRB_GRADIENT_COMPUTE_SHADER_SOBEL = 3 };
int  vrb_gradient_build(vrb_ctx* ctx, int mode);
int  vrb_gradient_mode(const vrb_ctx* ctx);
int  vrb_gradient_read(vrb_ctx* ctx, float* host_xyz);   /* w*h*d*3 floats, x fastest */
/* rc1pass with the lighting uniforms (rc1prenderer.cpp:112-135): light->apply_phong == 1 runs ShadeBlinnPhong
 * (ray_marching_1p.comp:48-81) on every non-transparent sample; apply_phong == 0 is vrb_rc1pass_render. */
int  vrb_rc1pass_render_lit(vrb_ctx* ctx, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_lighting* light);

/* Adaptive-step isosurface ray casting: replaces the dispatch of rc1pisoadapt/ray_marching_1p_iso_adapt.comp made by
 * RayCasting1PassIsoAdapt::Redraw (rc1pisoadaptrenderer.cpp:168-179), uniforms of ::Update (:113-165) and the
 * constructor defaults (:13-22).  No transfer function; light->apply_phong shades the hits on the gradient texture. */
typedef struct vrb_iso_params {
  float isovalue;           /* Isovalue, normalised density (default 0.5) */
  float step_size_small;    /* StepSizeSmall: while |previous sample - isovalue| < step_size_range (default 0.05) */
  float step_size_large;    /* StepSizeLarge: elsewhere (default 1.0) */
  float step_size_range;    /* StepSizeRange (default 0.1) */
  float color[4];           /* Color of the surface, alpha included (default 0.66, 0.6, 0.05, 1) */
  int   count_samples;
} vrb_iso_params;
int  vrb_iso_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_iso_params* p);

/* ---- pixel multi-scaling (BaseVolumeRenderer::MULTISCALING, volrenderbase.h:28-33) -------------------------------- */
/* Replaces RenderFrameToScreen::UpdateScreenResolutionMultiScaling (libs/vis_utils/renderoutputframe.cpp:89-145): the
 * marchers' frame becomes (screen_w * mw, screen_h * mh) for positive multipliers and (screen_w / |mw|, screen_h / |mh|)
 * for negative ones (the reference uses +-MULTISAMPLE_NUMBEROFSAMPLES = +-2, cppvolrend/defines.h:16-17), and a
 * screen_w x screen_h filtered frame is created next to it.  vrb_frame_resize returns to one ray per pixel. */
int  vrb_frame_resize_multiscaling(vrb_ctx* ctx, int screen_w, int screen_h, int mw, int mh);
/* The image-space pass that follows the dispatch in MultiSampleRedraw / DownScalingRedraw / UpScalingRedraw:
 * DrawMultiSampleHigherResolutionMode (:265-301, multisample_filter.comp), DrawHigherResolutionWithDownScale (:303-418,
 * downscaling_filter.comp + digital filter AFTER it for the cardinal kernels), DrawLowerResolutionWithUpScale (:420-540,
 * digital filter in place on the rendered frame BEFORE upscaling_filter.comp).  kernel = vis::IMAGE_FILTER_KERNEL
 * (libs/vis_utils/filters/utils.hpp:8-15; the reference's default is K2_HAT), ignored by the multisample pass. */
enum { VRB_FILTER_PASS_MULTISAMPLE = 1, VRB_FILTER_PASS_DOWNSCALE = 2, VRB_FILTER_PASS_UPSCALE = 3 };
enum { VRB_KERNEL_BOX = 0, VRB_KERNEL_HAT = 1, VRB_KERNEL_CATMULL_ROM = 2, VRB_KERNEL_MITCHELL_NETRAVALI = 3,
       VRB_KERNEL_CARDINAL_BSPLINE_3 = 4, VRB_KERNEL_CARDINAL_OMOMS3 = 5 };
int  vrb_frame_filter(vrb_ctx* ctx, int pass, int kernel);
int  vrb_filtered_frame_info(vrb_ctx* ctx, void** dev_rgba16f, int* w, int* h);
int  vrb_filtered_frame_read_rgba32f(vrb_ctx* ctx, float* host_out);

/* ---- sort-last: one brick of a volume that does not fit / is split over GPUs ------------ */
/* The context holds ONE brick: the voxel array given to vrb_volume_upload covers the owned region plus the ghost
 * layers listed here (one ghost layer on every interior face is enough for trilinear sampling).  The marcher walks the
 * ray of the WHOLE volume with the single-GPU sample positions s_k and composites only the samples whose voxel cell
 * lies in the owned region, so the per-brick partial results concatenate along the ray. */
typedef struct vrb_brick {
  int global_dims[3];       /* resolution of the whole volume */
  int origin[3];            /* first owned voxel in global coordinates */
  int owned[3];             /* owned extent */
  int ghost_lo[3];          /* ghost layers present in the uploaded array below the owned region (0 or more) */
  int ghost_hi[3];          /* ... and above it */
} vrb_brick;
/* rc1pass over the brick; writes premultiplied float RGBA (W*H*4 floats, 0 where the ray does not cross the brick)
 * into the context's partial frame (vrb_partial_device_ptr). */
int  vrb_rc1pass_render_brick(vrb_ctx* ctx, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* brick);
int  vrb_partial_device_ptr(vrb_ctx* ctx, void** dev_rgba32f);
/* Exact two-pass variant (the 0.99 cut falls on the same sample as on one GPU): pass 1 writes the opacity of the brick's
 * segment (float per pixel, vrb_brick_alpha_device_ptr); pass 2 starts from the opacity accumulated by the bricks in
 * FRONT (their pass-1 buffers in visibility order, local or peer memory).  The pass-2 partial frames are combined with
 * vrb_composite_sum: colours add, alpha is the maximum. */
int  vrb_rc1pass_brick_alpha(vrb_ctx* ctx, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* brick);
int  vrb_brick_alpha_device_ptr(vrb_ctx* ctx, void** dev_alpha32f);
int  vrb_rc1pass_render_brick_exact(vrb_ctx* ctx, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* brick,
                                    const void* const* front_alphas_in_order, int n_front);
int  vrb_composite_sum(vrb_ctx* ctx, const void* const* partials, int n, int row0, int rows);
/* Front-to-back "over" of n partial frames given IN VISIBILITY ORDER (device pointers, possibly peer memory mapped
 * with vrb_ipc_import): rows [row0, row0+rows) are composited with the reference's 0.99 opacity cut applied between
 * segments and written to the context's RGBA16F frame.  One kernel reads the peers' buffers directly (P2P loads over
 * NVLink): transfer and compositing are the same pass. */
int  vrb_composite_ordered(vrb_ctx* ctx, const void* const* partials_in_order, int n, int row0, int rows);
/* CUDA IPC plumbing for one-process-per-GPU peers (handle = cudaIpcMemHandle_t, 64 bytes). */
int  vrb_ipc_export(vrb_ctx* ctx, const void* dev_ptr, unsigned char handle[64]);
int  vrb_ipc_import(vrb_ctx* ctx, const unsigned char handle[64], void** dev_ptr);
int  vrb_ipc_close(vrb_ctx* ctx, void* dev_ptr);

/* Sort-last bricks of the voxel-cone-tracing renderer (BASELINE config 5: rc1pass + VCT shadows over bricks).  The
 * uploaded window must carry ghost layers that cover the cone reach plus the trilinear footprint of the coarsest
 * level the cones touch, and start / stop on multiples of 2^(n_levels-1) (dist.py: vct_brick_plan).  Pre-passes:
 *   vrb_sv_build_brick      levels 0..n_levels-1 of the window's mean/stddev pyramid (texels identical to the whole
 *                           volume's, PreProcessSuperVoxels preprocessingstages.cpp:35-145); largest deviation found
 *   vrb_sv_top_means_read   fp64 means of the brick's OWNED texels of the last level (gathered across bricks ...)
 *   vrb_sv_reduce_top       (... to finish the levels above on any one context:) largest deviation of the coarser levels
 *   vrb_preint_build        the pre-integration LUT for the deviation range of the WHOLE volume (:147-202)
 * Rendering: vrb_vct_render_brick in VRB_BRICK_ALPHA mode (opacity of the segment into vrb_brick_alpha_device_ptr),
 * then VRB_BRICK_EXACT with the front bricks' opacity buffers, then vrb_composite_sum -- or VRB_BRICK_SEGMENT and
 * vrb_composite_ordered.  The declarations follow vrb_vct_params below. */
enum { VRB_BRICK_SEGMENT = 0, VRB_BRICK_ALPHA = 1, VRB_BRICK_EXACT = 2 };
int  vrb_sv_build_brick(vrb_ctx* ctx, const vrb_brick* brick, int n_levels, double* local_max_stddev);
int  vrb_sv_top_means_read(vrb_ctx* ctx, const vrb_brick* brick, double* host_out, size_t cap_doubles, int dims_out[3], int origin_out[3]);
int  vrb_sv_reduce_top(vrb_ctx* ctx, const double* level_means, int w, int h, int d, double* max_stddev);
int  vrb_preint_build(vrb_ctx* ctx, const float* opc_by_density, int n_opc, double max_stddev);

/* ---- extinction-based shading (rc1pextbsd) ------------------------------------------------------------------- */
/* Replaces RC1PExtinctionBasedShading::GenerateExtinctionSAT3DTex (ebsrenderer.cpp:624-723) +
 * SummedAreaTable3D<double>::BuildSAT (libs/vis_utils/summedareatable.h:218-278): inclusive 3-D prefix sum of the
 * per-voxel extinction over the zero-bordered (W+2)(H+2)(D+2) grid, fp64 accumulate, fp32 store.
 * ext_lut[v] = tf->GetExtN(v / max) for every voxel value v (256 or 65536 floats), computed by the host. */
int  vrb_sat_build(vrb_ctx* ctx, const float* ext_lut, int n_lut);
/* Evaluation order of the fp64 accumulation (the sums are not exactly representable in fp64, so the order decides the
 * last bit of a few float texels, and the marcher's box queries are differences of ~1e7-1e8-sized texels):
 *   VRB_SAT_ORDER_REFERENCE (default) BuildSAT's own recurrence, S = v + S(x-1,y-1,z-1) + S(x,y,z-1) + ... left to right,
 *                           run as anti-diagonal wavefronts: the float SAT is bit-identical to the reference's;
 *   VRB_SAT_ORDER_SCAN      three separable fp64 scan passes (HBM-bound, several times faster): equal to the reference up to
 *                           one float ulp in ~1e-6 of the texels.
 * Also selectable with the environment variable VRB_SAT_ORDER=reference|scan at context creation. */
#define VRB_SAT_ORDER_REFERENCE 0
#define VRB_SAT_ORDER_SCAN      1
int  vrb_sat_set_order(vrb_ctx* ctx, int order);
int  vrb_sat_get_order(const vrb_ctx* ctx);
/* Host-side predicate (no device needed): 1 when every multiple k * step_size, k up to the sample count of a ray of length
 * longest_ray, is exactly representable in fp32.  The shaders' `s = s + h` then produces exactly those multiples, and a
 * sort-last brick may start its ray loop at s = k0 * step_size instead of adding k0 times (bit-identical).
 * True for the reference's default step 0.5 (rc1prenderer.cpp:62-63 with unit voxels). */
int  vrb_step_multiples_exact_f(float step_size, float longest_ray);
/* Sharded SAT build for sort-first runs (no reference counterpart: the reference builds the table on one CPU thread,
 * summedareatable.h:218-278).  Rank r owns the slices [z_lo, z_hi) of the bordered (D+2)-slice grid:
 *   vrb_sat_build_slab   three scan passes restricted to the slab (fp64), the context's float SAT is (re)allocated full size;
 *   vrb_sat_slab_plane   device pointer to the slab's last fp64 plane, (W+2)(H+2) doubles: all-gather these;
 *   vrb_sat_finish_slab  adds the caller's prefix plane (sum of the last planes of all slabs below; NULL for the first slab)
 *                        to every slice of the slab and stores the float texels into the context's SAT;
 *   vrb_sat_device_ptr   the full float SAT, x fastest: exchange the slabs so that every rank holds the whole table;
 *   vrb_sat_commit       builds the layout the marcher samples (gather atlas).
 * Results equal vrb_sat_build in VRB_SAT_ORDER_SCAN up to the fp64 association (bit-equal for integer-valued extinction). */
int  vrb_sat_build_slab(vrb_ctx* ctx, const float* ext_lut, int n_lut, int z_lo, int z_hi);
int  vrb_sat_slab_plane(vrb_ctx* ctx, void** dev_plane_fp64, size_t* count);
int  vrb_sat_finish_slab(vrb_ctx* ctx, const void* dev_prefix_plane_fp64);
int  vrb_sat_device_ptr(vrb_ctx* ctx, void** dev, int dims[3]);
int  vrb_sat_commit(vrb_ctx* ctx);
/* Integer mode (bit-exact): same scan over integer weights lut_u32[v]; result as u64, no border. */
int  vrb_sat_build_u64(vrb_ctx* ctx, const uint32_t* lut_u32, int n_lut, uint64_t* host_out);
/* Read back the float SAT ((W+2)*(H+2)*(D+2) floats, x fastest). */
int  vrb_sat_read(vrb_ctx* ctx, float* host_out);

typedef struct vrb_ebs_params {
  float step_size;
  int   apply_occlusion;            /* ApplyOcclusion */
  int   apply_shadow;               /* ApplyShadow */
  int   amb_occ_shells;             /* AmbOccShells (15) */
  float amb_occ_radius;             /* AmbOccRadius (1.0) */
  float sdw_cone_angle_rad;         /* DirSdwConeAngle, radians (ebsrenderer.cpp:166) */
  float sdw_sample_interval;        /* DirSdwSampleInterval (2) */
  float sdw_initial_step;           /* DirSdwInitialStep (2) */
  float sdw_ui_weight;              /* DirSdwUserInterfaceWeight (1) */
  float sdw_cone_max_distance;      /* DirSdwConeMaxDistance (0.75 * diagonal) */
  int   type_of_shadow;             /* TypeOfShadow: 0 point, 1 directional */
  int   count_samples;
} vrb_ebs_params;
int  vrb_ebs_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_ebs_params* p);

/* ---- directional occlusion + cone shadows by cone tracing (rc1pdosct) --------------------------------------------- */
/* Replaces ExtinctionCoefficientVolume::BuildMipMappedTexture (extcoefvolumegenerator.cpp:20-40,92-408) and the
 * glslextgen compute passes: Gaussian (7^3 taps, sigma0 * 2^level) mip pyramid of TF opacity, converted to extinction,
 * every level stored R16F.  (rw,rh,rd) = base resolution (128^3 by default, :10-15); 0 = same size as the volume.
 * Needs the RGBA (opacity) transfer function of vrb_tf_upload. */
int  vrb_extcoef_build(vrb_ctx* ctx, float sigma0, int rw, int rh, int rd);
int  vrb_extcoef_info(vrb_ctx* ctx, int* n_levels, int* dims_xyz, int cap_levels);
int  vrb_extcoef_read_level(vrb_ctx* ctx, int level, float* host_out);     /* w*h*d floats, x fastest */

/* Uniforms of one ConeGaussianSampler (conegaussiansampler.h:21-163) as BindConeOcclusionUniforms /
 * BindConeShadowUniforms upload them (dosrcrenderer.cpp:805-985).  sections = the GL_FLOAT client array of
 * GetConeSectionsInfoTex (n x [interval distance, mip level, d_integral, amplitude]); rounded to RGBA16F on upload. */
typedef struct vrb_cone_sampler {
  const float* sections;
  int   n_sections;
  int   integration_samples[3];     /* gaussian_samples_1 / _3 / _7 */
  float initial_step;
  float ray7_adj_weight;
  float ui_weight;
  float ray_axes[10][3];            /* Get3ConeRayID(0..2), Get7ConeRayID(0..6) */
} vrb_cone_sampler;
int  vrb_dos_set_cones(vrb_ctx* ctx, const vrb_cone_sampler* occlusion, const vrb_cone_sampler* shadow);

typedef struct vrb_dos_params {
  float step_size;
  int   apply_occlusion;            /* ApplyOcclusion (on by default) */
  int   apply_shadow;               /* ApplyShadow (off by default, dosrcrenderer.cpp:53) */
  int   type_of_shadow;             /* 0 point, 1 spot, 2 directional */
  float spot_cos;                   /* SpotLightMaxAngle uniform = cos(pi * angle / 180) (dosrcrenderer.cpp:159) */
  int   count_samples;
} vrb_dos_params;
int  vrb_dos_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_dos_params* p);

/* ---- object-space light cache (PreIlluminationStructuredVolume, cppvolrend/utils/preillumination.cpp:7-80) ---------- */
/* Secondary mode of the shaded renderers, inactive by default in the reference (dosrcrenderer.cpp:63-64, 32^3): every
 * Update recomputes an RG16F volume of (Iocc, Ishadow) and Redraw dispatches _common_shaders/obj_ray_marching.comp
 * (:210-333), which reads the shading from that volume instead of tracing cones per sample.
 * vrb_dos_light_cache_build replaces PreComputeLightCache (dosrcrenderer.cpp:555-657) = the dispatch of
 * rc1pdosct/lightcachecomputation.comp; eye = camera->GetEye(), eye_up = the v_up of Camera::GetCameraVectors
 * (the EyeCamUp uniform).  Needs the pyramid (vrb_extcoef_build) and the cone samplers (vrb_dos_set_cones). */
int  vrb_dos_light_cache_build(vrb_ctx* ctx, const float eye[3], const float eye_up[3], const vrb_lighting* light,
                               const vrb_dos_params* p, int res_w, int res_h, int res_d);
/* (vrb_ebs_light_cache_build / vrb_vct_light_cache_build: at the end of this header, after their parameter blocks.) */
/* Read the cache back: res_w*res_h*res_d pairs (Iocc, Ishadow), x fastest; host_out_rg may be NULL to query dims only. */
int  vrb_light_cache_read(vrb_ctx* ctx, float* host_out_rg, int dims_out[3]);
typedef struct vrb_obj_params {
  float step_size;
  int   apply_occlusion, apply_shadow;   /* a sample is composited only if one of them is on (obj_ray_marching.comp:312) */
  int   count_samples;
} vrb_obj_params;
/* Replaces the dispatch of obj_ray_marching.comp (:210-333); lighting supplies Kambient / Kdiffuse and, with
 * apply_phong = 1 (needs vrb_gradient_build), the gradient Blinn-Phong branch of its ShadeSample (:236-257). */
int  vrb_obj_march_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_obj_params* p);

/* ---- cone ground truth: many occlusion / shadow rays per sample (rc1pcrtgt) --------------------------------------- */
/* Ray-direction tables (n x RGB, GL_FLOAT client arrays of the two RGB16F 1-D textures of
 * RC1PConeLightGroundTruthSteps::Update, crtgtrenderer.cpp:131-187); rounded to fp16 on upload. */
int  vrb_gt_set_rays(vrb_ctx* ctx, const float* occ_rays, int n_occ, const float* sdw_rays, int n_sdw);
typedef struct vrb_gt_params {
  float step_size;                  /* StepSize */
  float light_ray_initial_gap;      /* LightRayInitialGap (1.0) */
  float light_ray_step_size;        /* LightRayStepSize (0.5) */
  int   apply_occlusion;            /* ApplyConeOcclusion */
  int   occ_num_rays;               /* OccNumberOfSampledRays */
  float occ_cone_distance;          /* OccConeDistanceEvaluation (0.5 * diagonal) */
  int   apply_shadow;               /* ApplyConeShadow */
  int   sdw_num_rays;               /* SdwNumberOfSampledRays */
  float sdw_cone_distance;          /* SdwConeDistanceEvaluation (0.75 * diagonal) */
  int   shadow_type;                /* SdwShadowType: 0 point, 1 spot, 2 directional */
  int   count_samples;
} vrb_gt_params;
/* One call renders the CONVERGED image of the reference's progressive loop (crtgtrenderer.cpp:272-325): every ray is
 * marched to the end with the per-dispatch rgba16f / rg16f state round trips applied.  light->light_forward must hold
 * this shader's LightCamForward uniform (= -GetBlinnPhongLightSourceCameraForward(), crtgtrenderer.cpp:226-236). */
int  vrb_gt_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_gt_params* p);
/* Replaces RC1PConeLightGroundTruthSteps::RedrawCube (crtgtrenderer.cpp:327-338, vol_intersection.comp:64-110): the
 * bounding-box placeholder the reference shows while "Show Generated Frame Texture" is off (its default, and after every
 * camera or parameter change): every ray's entry point on the box coloured by the face it lies on. */
int  vrb_gt_cube_render(vrb_ctx* ctx, const vrb_camera* cam);

/* ---- voxel-cone-traced shadows over a mean/stddev super-voxel pyramid (rc1pvctsg) -------------------------------- */
/* Replaces VCTPreProcessing::PreProcessSuperVoxels + PreProcessPreIntegrationTable (preprocessingstages.cpp:35-202),
 * both CPU loops in the reference.  opc_by_density[i] = tf->GetOpc(i, maxDensity) for i = 0..maxDensity
 * (256 or 65536 floats), computed by the host. */
int  vrb_vct_build(vrb_ctx* ctx, const float* opc_by_density, int n_opc);
int  vrb_vct_info(vrb_ctx* ctx, int* n_levels, int* dims_xyz, int cap_levels, int* lut_w, int* lut_h, float* max_stddev);
/* level >= 0: super-voxel level as w*h*d x (mean, stddev) floats; level == -1: the LUT as lut_w*lut_h floats */
int  vrb_vct_read(vrb_ctx* ctx, int level, float* host_out);
typedef struct vrb_vct_params {
  float step_size;
  int   apply_occlusion;            /* ApplyOcclusion: ambient term, constant 1 (vctrenderer.cpp:33) */
  int   apply_shadow;               /* ApplyShadow: voxel cone tracing */
  float tan_cone_apex_angle;        /* TanRadiusConeApexAngle = tan(angle * pi / 180) (vctrenderer.cpp:145) */
  float cone_step_size;             /* ConeStepSize (2) */
  float cone_step_increase_rate;    /* ConeStepIncreaseRate (1) */
  float cone_initial_step;          /* ConeInitialStep (2) */
  float opacity_correction_factor;  /* OpacityCorrectionFactor (2) */
  int   apply_opacity_correction;   /* ApplyOpacityCorrectionFactor */
  int   cone_number_of_samples;     /* ConeNumberOfSamples (50) */
  float volume_max_density;         /* VolumeMaxDensity = (float)GetMaxDensity() */
  float volume_max_stddev;          /* VolumeMaxStandardDeviation = (float)maximum_standard_deviation (vrb_vct_info) */
  int   count_samples;
} vrb_vct_params;
int  vrb_vct_render(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_vct_params* p);
/* One brick of a sort-last partition (see "Sort-last bricks of the voxel-cone-tracing renderer" above); mode is one of
 * VRB_BRICK_SEGMENT / VRB_BRICK_ALPHA / VRB_BRICK_EXACT, front_alphas only with VRB_BRICK_EXACT. */
int  vrb_vct_render_brick(vrb_ctx* ctx, const vrb_camera* cam, const vrb_lighting* light, const vrb_vct_params* p, const vrb_brick* brick,
                          int mode, const void* const* front_alphas_in_order, int n_front);

/* ---- light cache of the other two shaded renderers (see "object-space light cache" above) ----------------------- */
/* rc1pextbsd/lightcachecomputation.comp dispatched by ebsrenderer.cpp:441-555 (needs vrb_sat_build), and
 * rc1pvctsg/lightcachecomputation.comp dispatched by vctrenderer.cpp:393-515 (needs vrb_vct_build; its cone has no
 * leave-the-volume cut and Iocc is always 1).  Both feed vrb_obj_march_render. */
int  vrb_ebs_light_cache_build(vrb_ctx* ctx, const vrb_lighting* light, const vrb_ebs_params* p, int res_w, int res_h, int res_d);
int  vrb_vct_light_cache_build(vrb_ctx* ctx, const vrb_lighting* light, const vrb_vct_params* p, int res_w, int res_h, int res_d);

#ifdef __cplusplus
}
#endif
#endif /* VRB200_H */
