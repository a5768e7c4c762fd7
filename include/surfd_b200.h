/* surfd_b200.h -- C ABI of the B200-native Surf-D generation hot path.
 *
 * The reference (Yzmblog/SurfD) has no FFI registry; its hot path sits behind three Python call
 * boundaries.  Each entry point below names the reference interface it replaces.
 * Conventions: every pointer marked "dev" is CUDA device memory owned by the caller (torch tensors
 * in the shipped host code); the library never frees caller memory, never throws, and returns an
 * int status: 0 ok, >0 domain status (see SURFD_* below), <0 = -(cudaError_t).  `stream` is a
 * cudaStream_t passed as void*.  Work is stream-ordered; calls that must report a count to the host
 * (documented per function) synchronise that stream once.  One host thread per GPU.
 */
#ifndef SURFD_B200_H
#define SURFD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURFD_OK 0
#define SURFD_EMPTY_SURFACE 1   /* reference: RuntimeError("No surface found ...") _marching_cubes_lewiner.py:130-131 */
#define SURFD_CAPACITY 2        /* output buffers too small; *n_v / *n_f hold the required sizes */
#define SURFD_QUEUE_OVERFLOW 3  /* internal BFS queue bound exceeded (pathological field) */
#define SURFD_BAD_ARGUMENT 4    /* reference: ValueError on bad shapes _marching_cubes_lewiner.py:102-105 */
#define SURFD_ABORTED 5         /* the persistent sampler kernel gave up at a grid barrier (never expected) */
#define SURFD_RANGE 6           /* persistent sampler, precision mode 1: an activation left the fp16 range of the split products */

int surfd_version(void);
/* last CUDA / argument error text of the calling thread (static buffer) */
const char* surfd_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Decoder: CoordsEncoder + CbnDecoder + udf_func closure
 *   AutoEncoder/models/coordsenc.py:34-51, AutoEncoder/models/cbndec.py:16-134,
 *   sample/generate_uncond.py:96-101 (udf = (1 - sigmoid(logit)) * 0.1)
 * `packed` (host or device pointer, see `packed_on_device`) is the float32 blob produced by
 * surfd_b200.decoder.pack_decoder(): fc_p W[512][64] (63 padded), b[512]; 5x{fc_0 W[512][512], b,
 * fc_1 W, b}; fc_out w[512], b[4]; 11x CBN {Wg[512][L], bg[512], Wb[512][L], bb[512], mean[512], var[512]}.
 * ------------------------------------------------------------------------------------------- */
typedef struct surfd_decoder surfd_decoder;

int surfd_dec_create(const float* packed, size_t n_floats, int latent_dim, int packed_on_device,
                     int max_chunk_points, surfd_decoder** out);
void surfd_dec_destroy(surfd_decoder* d);
/* size of the packed blob in floats for a latent dimension */
size_t surfd_dec_packed_floats(int latent_dim);

/* Fold the conditional batch-norm of one shape: s = gamma(z)/sqrt(var+1e-5), t = beta(z) - s*mean
 * (cbndec.py:68-82).  lat_dev: [L] float32 dev.  Must precede queries for that shape. */
int surfd_dec_set_latent(surfd_decoder* d, const float* lat_dev, void* stream);

/* precision: 0 = fp32 FFMA (parity mode), 1 = TF32 tcgen05 tensor cores (fast mode) */
int surfd_dec_set_precision(surfd_decoder* d, int mode);
/* grid size of the persistent tcgen05 GEMM (0 = every SM): leave SMs to concurrently running marching-cubes replays */
int surfd_dec_set_sm_budget(surfd_decoder* d, int n_sms);
int surfd_dec_num_sms(surfd_decoder* d);
/* TF32 mode: 1 (default) = all 512x512 layers of a pass in one launch (each CTA runs every layer on its own 128-row panels), 0 = one launch
 * per layer; bit-identical results */
int surfd_dec_set_chain(surfd_decoder* d, int on);

/* measurement hooks (bench.py): points per internal chunk; average duration of the dominant kernel (one 512x512
 * layer GEMM over M <= chunk points) timed with CUDA events on `stream`, `iters` back-to-back launches. */
int surfd_dec_chunk_points(surfd_decoder* d);
int surfd_dec_time_layer(surfd_decoder* d, int M, int iters, float* ms_per_launch, void* stream);
/* Measurement of the same kernel inside the real layer chain (lattice / face filter calls): on = 1 starts (every layer GEMM
 * is then bracketed by its own CUDA event pair on its stream), on = 0 stops, synchronises the device and reports the
 * launch count, the point rows processed (2 * 512 * 512 FLOP each) and the summed launch durations. */
int surfd_dec_profile(surfd_decoder* d, int on, int64_t* launches, int64_t* points, double* total_ms);
/* test hook: one 512x512 layer (fc_0 of block `blk` + CBN/ReLU epilogue) over A_dev [M][512] with kernel `mode`
 * (0 fp32 FFMA, 1 tcgen05 TF32), both on the TF32-rounded weights; out_dev [M][512] */
int surfd_dec_debug_layer(surfd_decoder* d, const float* A_dev, int M, int blk, int mode, float* out_dev, void* stream);

/* udf (and optionally -normalize(d udf/dx), meshudf.py:231-251) at explicit points.
 * pts_dev [M][3]; udf_dev [M]; grad_dev [M][3] or NULL.  Replaces udf_func / sample_udf /
 * sample_grads (meshudf.py:209-251) for the (CbnDecoder, latent) closure. */
int surfd_udf_query(surfd_decoder* d, const float* pts_dev, int64_t M, float* udf_dev, float* grad_dev,
                    void* stream);

/* decoder logits at explicit points: CbnDecoder.forward(CoordsEncoder.encode(c), lat) (cbndec.py:127-134,
 * coordsenc.py:34-51) -- what a udf_func closure kept from the reference scripts calls. pts_dev [M][3]; logit_dev [M]. */
int surfd_dec_logits(surfd_decoder* d, const float* pts_dev, int64_t M, float* logit_dev, void* stream);

/* Whole lattice of one shape.  mode 0: dense get_udf_and_grads (meshudf.py:254-304);
 * mode 1: coarse-to-fine GridFiller.fill_grid (meshudf.py:36-206).  udf_dev [N^3], grad_dev [N^3][3]
 * (zero where not evaluated), both also clamped like meshudf.py:342.  counts (host, may be NULL):
 * [0] udf evaluations, [1] gradient evaluations.  Synchronises `stream` (level sizes are read back).
 * grad_dev NULL: udf only -- utils.GridFiller.fill_grid (utils/utils.py:252-339), the filler of the --watertight branch. */
int surfd_udf_lattice(surfd_decoder* d, int N, int mode, double max_dist, float* udf_dev, float* grad_dev,
                      int64_t* counts_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Marching cubes: meshudf/_marching_cubes_lewiner_cy.pyx marching_cubes_udf (:1115-1773) behind
 * udf_mc_lewiner (_marching_cubes_lewiner.py:87-154).  Output numbering is the reference's own
 * (traversal order).  Vertices are float32 pyx-level vertices (x,y,z)=(axis2,axis1,axis0) in index
 * units (the wrapper's fliplr/spacing is applied by the host shell exactly as numpy does); faces
 * int32 in raw pyx order.  surfd_mc_udf synchronises `stream` to read the counts.
 * ------------------------------------------------------------------------------------------- */
typedef struct surfd_mc surfd_mc;
int surfd_mc_create(surfd_mc** out);
void surfd_mc_destroy(surfd_mc* m);
/* Runs classification + ordered replay into buffers owned by the handle (sized from the candidate
 * count) and reports the vertex / face counts; surfd_mc_fetch() then copies them to caller memory. */
int surfd_mc_udf(surfd_mc* m, const float* udf_dev, const float* grad_dev, int N, int64_t* n_v, int64_t* n_f,
                 int64_t* stats_host /* [8] or NULL: n_cand, n_seed, n_accept, n_unsure, n_nontrivial */, void* stream);
/* The same in two halves, so several shapes can be in flight on different streams: _launch enqueues everything
 * without a host synchronisation (the candidate count stays on the device); _finish waits for that stream and
 * returns counts/status.  SURFD_CAPACITY from _finish means the per-handle buffers were grown: launch again. */
int surfd_mc_launch(surfd_mc* m, const float* udf_dev, const float* grad_dev, int N, void* stream);
int surfd_mc_finish(surfd_mc* m, int64_t* n_v, int64_t* n_f, int64_t* stats_host);
/* Diagnostics: cycle counters of the last finished replay (all zero unless the library was built with -DMC_PROFILE):
 * total, fetch, sign propagation, tiling selection, emission, visits, window refills, 0. */
int surfd_mc_profile(surfd_mc* m, int64_t* prof_host /* [8] */);
int surfd_mc_fetch(surfd_mc* m, float* verts_dev /* [n_v][3] */, int32_t* faces_dev /* [n_f][3] */, void* stream);
/* classification pass alone (HBM-bound scan, pyx:1157-1158,1215-1218): candidate bitmask words
 * [ceil(N^3/32)] and count; used by the benchmarks / tests. */
int surfd_mc_classify(surfd_mc* m, const float* udf_dev, int N, uint32_t* bits_dev_or_null,
                      int64_t* n_cand_host, void* stream);

/* measurement hook: average duration of the classification kernel alone (`iters` launches, CUDA events on `stream`) */
int surfd_mc_time_classify(surfd_mc* m, const float* udf_dev, int N, int iters, float* ms_per_launch, void* stream);

/* UDF face filter of get_mesh_from_udf (meshudf.py:356-379): the reference evaluates the decoder at both end
 * points and the midpoint of every directed face edge (9 points per face, float64 positions rounded to float32
 * like `torch.from_numpy(points).float()`), keep[f] = 0 if any udf > 1/N.  Same decisions here with every vertex
 * evaluated once (V + 3F decoder queries instead of 9F).
 * verts64_dev [n_v][3] float64 final vertex positions, faces_dev [n_f][3] int32. */
int surfd_face_filter(surfd_decoder* d, const double* verts64_dev, int64_t n_v, const int32_t* faces_dev, int64_t n_f,
                      int N, uint8_t* keep_dev /* [n_f] */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sampler: GaussianDiffusion.p_sample_loop over MDM/UNetModel
 *   diffusion/gaussian_diffusion.py:570-708 (loop), :471-520 (p_sample), :258-363 (p_mean_variance)
 *   diffusion/respace.py:63-132, models/mdm.py:91-110, models/openaimodel.py:710-749
 * ------------------------------------------------------------------------------------------- */
typedef struct surfd_unet surfd_unet;
/* packed: float32 weight blob and `program`: int64 op program, both from surfd_b200.unet.pack_unet() (the
 * architecture walk of models/openaimodel.py:443-692 happens on the host; the device side interprets
 * GN / token-GEMM / attention ops -- layout in DESIGN.md); L = 32 or 64.  Pointers may be host or device. */
int surfd_unet_create(const float* packed, size_t n_floats, const int64_t* program, size_t n_program, int L,
                      int max_batch, surfd_unet** out);
void surfd_unet_destroy(surfd_unet* u);
/* The denoiser is latency-bound (~170 dependent small kernels per step), so surfd_sample() splits a batch over `n_lanes`
 * concurrent streams (private activations, shared weights).  Default 1 (measured fastest).  Results do not depend on it. */
int surfd_unet_set_lanes(surfd_unet* u, int n_lanes);
/* surfd_sample engine.  mode 1 (default): one persistent cooperative kernel runs the whole reverse-diffusion loop on `n_sms`
 * CTAs (0 = one per SM), ops separated by grid barriers; token GEMMs are wide units (32 tokens x 128 outputs x one K slice,
 * every op a single round of the resident CTAs) on tcgen05 tensor cores with fp32-class split products.  mode 0: CUDA-graph
 * replay of the per-step kernel sequence (~170 nodes, one graph launch per DDPM step); also used when the device cannot launch
 * cooperatively and for batches above 64 samples per call (the persistent engine's verified range).  mode 2: the persistent
 * kernel with the graph path's units and K split -- samples are bit-identical to mode 0 and independent of n_sms.
 * In mode 1 the token GEMMs' weights arrive as packed 32 KB images through cp.async.bulk (a ring that one thread keeps full
 * across barriers and ops), and classifier-free guidance runs both forwards of a step as one pass over 2B rows.
 * Measured on B200 at batch 8: 1.26 (mode 1) / 2.28 (mode 0) / 1.98 (mode 2) ms per DDPM step.
 * All modes compute in the precision selected by surfd_unet_set_precision and agree to fp32 rounding. */
int surfd_unet_set_sampler(surfd_unet* u, int mode, int n_sms);
/* Diagnostics: out == NULL switches the persistent kernel's per-op-type cycle counters on/off; out != NULL reads the
 * last run's counters: out[(h * 8 + op) * 3 + {body cycles, barrier cycles, count}], h = 0 first CTA / 1 last CTA,
 * op = 1 emb1, 2 linear, 3 in-conv, 4 group norm, 5 token GEMM, 6 attention, 7 out-conv + DDPM update. */
int surfd_unet_profile(surfd_unet* u, int on, int64_t* out /* [64] or NULL */);
/* SURFD_ABORTED if the last persistent run reported a barrier time-out, SURFD_RANGE if one of its activations left the fp16
 * range the split-product token GEMMs need (|x| < 6e4, finite); valid after the run's stream was synchronised. */
int surfd_unet_status(surfd_unet* u);
/* token-GEMM arithmetic: 0 = fp32 FFMA, 1 = fp32-class split products (default: 3xTF32 on mma.sync in the per-op kernels, fp16
 * hi + lo/4096 two-term split on tcgen05 in the persistent engine; both 2^-22 relative), 2 = single-pass TF32 */
int surfd_unet_set_precision(surfd_unet* u, int mode);
size_t surfd_unet_packed_floats(void);
/* one model evaluation x0_hat = model(x_t, t, context/labels): teacher-forced parity entry.
 * x_dev [B][L]; t_dev [B] int64 (already mapped through timestep_map); context_dev [B][512] or NULL;
 * labels_dev [B] int64 or NULL; out_dev [B][L]. */
int surfd_unet_forward(surfd_unet* u, int B, const float* x_dev, const int64_t* t_dev,
                       const float* context_dev, const int64_t* labels_dev, float* out_dev, void* stream);
/* full reverse process.  coef_dev [3][n_steps] float32: posterior_mean_coef1, posterior_mean_coef2,
 * exp(0.5*posterior_log_variance_clipped) (index = spaced step); tmap_dev [n_steps] int64;
 * noise_dev [n_steps+1][B][L]: row 0 = x_T, row 1+k = the randn_like draw of loop iteration k
 * (iteration k handles step index n_steps-1-k); guidance: 1 = single forward, otherwise the
 * reference's two-forward CFG combination (models/cfg_sampler.py:19-26) is replayed. */
int surfd_sample(surfd_unet* u, int B, int n_steps, const int64_t* tmap_dev, const float* coef_dev,
                 const float* noise_dev, const float* context_dev, const int64_t* labels_dev,
                 float guidance, float* out_dev, void* stream);

/* kernel-launch counter (all kernels launched by this library since load / last reset) */
int64_t surfd_launch_count(int reset);

/* ---- output stage, host side (no device work): the Wavefront .obj text conversion of o3d.io.write_triangle_mesh
 * (utils/utils.py:79-121, sample/generate_uncond.py:113-116) and of pymeshlab's load_new_mesh / save_current_mesh
 * (generate_uncond.py:117-122).  The layouts are the callers' (surfd_b200/output.py). */
/* header + nv lines "v x y z" + mid + nf lines "f a b c" (indices written 1-based) + footer.  number_format 0 = "%f"
 * (MeshLab exporter), 1 = "%g" (C++ ostream default: open3d).  verts: HOST double [nv][3]; faces: HOST int64 [nf][3], 0-based.
 * header / mid / footer may be NULL. */
int surfd_obj_write(const char* path, const char* header, const double* verts, int64_t nv, int number_format,
                    const char* mid, const int64_t* faces, int64_t nf, const char* footer);
/* `v x y z` and `f a[/..] b[/..] c[/..]` records (1-based), everything else skipped.  Call with verts = faces = NULL for the
 * counts (*nv, *nf), then with HOST buffers of those sizes (double [nv][3], int64 [nf][3], 0-based on return). */
int surfd_obj_read(const char* path, int64_t* nv, int64_t* nf, double* verts, int64_t* faces);

#ifdef __cplusplus
}
#endif
#endif
