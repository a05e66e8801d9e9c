/* diffmusic-b200 C ABI  --  the drop-in boundary of the guidance hot path.
 *
 * The reference (jwliao1209/DiffMusic) is pure Python and has no FFI of its own; every entry point below replaces a
 * chain of torch / torchaudio library calls made by the reference file:line cited next to it (paths relative to
 * /root/reference).  Callers are the Python classes in diffmusic_b200/ that mirror diffmusic/inverse_problem and
 * diffmusic/schedulers (see INTEGRATION.md for the ctypes binding).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes and strides (in elements), fp32 scalars by value, a CUDA stream;
 *   - every function returns 0 on success, <0 on error (DM_ERR_*); dm_last_error() gives the thread-local message;
 *   - nothing allocates, nothing synchronises; work is enqueued on `stream`; workspaces are caller-owned;
 *   - batched semantics are per clip: every norm / loss is taken over one clip.
 */
#ifndef DM_ABI_H_
#define DM_ABI_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DM_OK 0
#define DM_ERR_INVALID (-1)
#define DM_ERR_CUDA (-2)
#define DM_ERR_UNSUPPORTED (-3)

typedef void* dm_stream_t; /* cudaStream_t */

int dm_version(void);
const char* dm_last_error(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
unsigned long long dm_launch_count(void);
void dm_reset_launch_count(void);
/* Process-wide kernel-selection knobs for A/B measurements and tests (results do not depend on them):
 *   DM_TUNE_STREAM_KERNELS  1 (default): persistent-grid variants of the scale-2 resampling forward (always) and adjoint
 *                           (from about 32 ten-second clips per launch; below that the plain kernel is faster);
 *                           16-byte aligned rows, <= 1024 clips.  2: both always.  0: one CTA per 2048 samples. */
#define DM_TUNE_STREAM_KERNELS 0
/*   DM_TUNE_PDL             1: the kernels of a fused guidance chain (A(x) -> STFT guidance -> adjoint) are launched with
 *                           programmatic stream serialization: the next kernel's launch and prologue (table staging,
 *                           filter taps) overlap the tail of the previous one, its first dependent access waits
 *                           (griddepcontrol.wait); 0: plain stream order. */
#define DM_TUNE_PDL 1
#define DM_TUNE_COUNT 2
int dm_set_tuning(int knob, int value);
int dm_get_tuning(int knob);

/* ------------------------------------------------------------------------------------------------------------------
 * Scheduler algebra on the latent  (SURVEY.md Appendix B).  n = total elements, n_clip = C*H*W of one clip.
 * Scalars are the fp32 values the reference computes on the host (scheduling_dps.py:157-162):
 *   sqrt_a = alpha_prod_t ** 0.5, sqrt_b = (1 - alpha_prod_t) ** 0.5, sqrt_p = alpha_prod_t_prev ** 0.5,
 *   dir_coef = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** 0.5, std = eta * variance ** 0.5.
 * `coef` (may be NULL): DEVICE array [sqrt_a, sqrt_b, sqrt_p, dir_coef (ddim: sqrt_1mp), std, r]; when given it
 * overrides the by-value scalars, so a CUDA graph captured once can be replayed for every timestep by updating 24
 * bytes of device memory (the guided loop is launch-bound: SURVEY.md 7.3 "tiny problem sizes").
 * ---------------------------------------------------------------------------------------------------------------- */

/* x0 = (x - sqrt_b * eps) / sqrt_a, optionally clamped to +-clip_range.
 * Replaces diffusers DDIMScheduler.step(...).pred_original_sample as called at scheduling_ddim.py:84-93,
 * scheduling_dps.py:166-175, scheduling_mpgd.py:164-173, scheduling_dsg.py:178-186, scheduling_diffmusic.py:180-188. */
int dm_sched_x0(const float* x, const float* eps, float* x0, long long n, float sqrt_a, float sqrt_b, int clip,
                float clip_range, const float* coef, dm_stream_t stream);

/* scheduling_ddim.py:95-96: e = (x - sqrt_a x0)/sqrt_b ; prev = sqrt_p x0 + sqrt_1mp e */
int dm_sched_ddim_update(const float* x, const float* x0, float* prev, long long n, float sqrt_a, float sqrt_b,
                         float sqrt_p, float sqrt_1mp, const float* coef, dm_stream_t stream);

/* scheduling_dps.py:177-213: prev = sqrt_p x0 + dir_coef (x - sqrt_a x0)/sqrt_b (+ std z) - rate * g0 / sqrt_a.
 * g0 = dLoss/dx0 (from torch autograd through vocoder + VAE); z may be NULL when eta == 0. */
int dm_sched_dps_update(const float* x, const float* x0, const float* g0, const float* z, float* prev, long long n,
                        float sqrt_a, float sqrt_b, float sqrt_p, float dir_coef, float std, float rate,
                        const float* coef, dm_stream_t stream);

/* scheduling_mpgd.py:199-218: x0' = x0 - rate g0 ; prev = sqrt_p x0' + dir_coef (x - sqrt_a x0')/sqrt_b (+ std z) */
int dm_sched_mpgd_update(const float* x, const float* x0, const float* g0, const float* z, float* prev,
                         float* x0_out, long long n, float sqrt_a, float sqrt_b, float sqrt_p, float dir_coef,
                         float std, float rate, const float* coef, dm_stream_t stream);

/* scheduling_dsg.py:189-224 with per-clip norms: g = grad_scale * g0 / sqrt_a ; mean = sqrt_p x0 + dir_coef eps ;
 * d* = -r g/(|g|+e) ; mix = std z + rate (d* - std z) ; prev = mean + r mix/(|mix|+e).  r = sqrt(n_clip) * std. */
int dm_sched_dsg_update(const float* x0, const float* eps, const float* g0, const float* z, float* prev, int n_clips,
                        long long n_clip, float sqrt_a, float sqrt_p, float dir_coef, float std, float rate, float r,
                        float grad_scale, float e, const float* coef, dm_stream_t stream);

/* scheduling_diffmusic.py:191-223 + slerp (59-68), branch resolved on the device: g as above ;
 * u = -g/(|g|+e) |z| ; c = <z/|z|, u/|u|> ; m = |c| > thr ? z + rate (u - z) : slerp ; prev = mean + std m. */
int dm_sched_diffmusic_update(const float* x0, const float* eps, const float* g0, const float* z, float* prev,
                              int n_clips, long long n_clip, float sqrt_a, float sqrt_p, float dir_coef, float std,
                              float rate, float grad_scale, float e, float threshold, const float* coef,
                              dm_stream_t stream);

/* ---- the same six entry points with a dtype for the LATENT-TYPED tensors ----
 * The reference pipelines run in fp16 (run.py:218): `sample`, `model_output` come in and `prev_sample`,
 * `pred_original_sample` go back in the pipeline's dtype.  With io_dtype = DM_IO_F16 / DM_IO_BF16 the kernels read x / eps
 * and write prev (and x0_pub, the caller's copy of x0) in that type directly -- no cast kernels around the step; the
 * step's own x0, the gradient g0, the noise z and all arithmetic stay fp32.  DM_IO_F32 = the functions above. */
#define DM_IO_F32 0
#define DM_IO_F16 1
#define DM_IO_BF16 2
int dm_sched_x0_io(const void* x, const void* eps, float* x0, void* x0_pub /* may be NULL */,
                   void* x0_leaf /* may be NULL */, float leaf_scale, long long n, float sqrt_a, float sqrt_b, int clip,
                   float clip_range, const float* coef, int io_dtype, dm_stream_t stream);
int dm_sched_ddim_update_io(const void* x, const float* x0, void* prev, long long n, float sqrt_a, float sqrt_b,
                            float sqrt_p, float sqrt_1mp, const float* coef, int io_dtype, dm_stream_t stream);
/* The reference differentiates through `1 / vae.config.scaling_factor * x0` (scheduling_dps.py:195-197): two torch
 * kernels (the scaling and its backward) around the networks.  Here dm_sched_x0_io also writes the scaled, latent-typed
 * decoder input x0_leaf = leaf_scale * x0, autograd runs from that leaf, and the update kernels take g0 = dLoss/d(leaf)
 * in the latent dtype together with leaf_scale: dLoss/dx0 = leaf_scale * g0, applied first, exactly where torch's
 * multiply-backward would have rounded it.
 * losses / loss_total (may be NULL): per-clip losses of the batch and where to put their 2-norm = the reference's 0-d
 * `loss` (torch.linalg.norm over the whole batch, scheduling_dps.py:211), written by one thread of the update kernel. */
int dm_sched_dps_update_io(const void* x, const float* x0, const void* g0, float leaf_scale, const float* z, void* prev,
                           long long n, float sqrt_a, float sqrt_b, float sqrt_p, float dir_coef, float std,
                           float rate, const float* coef, int io_dtype, const float* losses, int n_losses,
                           float* loss_total, dm_stream_t stream);
/* x0_out is latent-typed here (it is only handed back to the caller) */
int dm_sched_mpgd_update_io(const void* x, const float* x0, const void* g0, float leaf_scale, const float* z, void* prev,
                            void* x0_out, long long n, float sqrt_a, float sqrt_b, float sqrt_p, float dir_coef,
                            float std, float rate, const float* coef, int io_dtype, const float* losses, int n_losses,
                            float* loss_total, dm_stream_t stream);
int dm_sched_dsg_update_io(const float* x0, const void* eps, const void* g0, float leaf_scale, const float* z,
                           void* prev, int n_clips, long long n_clip, float sqrt_a, float sqrt_p, float dir_coef,
                           float std, float rate, float r, float grad_scale, float e, const float* coef, int io_dtype,
                           const float* losses /* n_clips */, float* loss_total, dm_stream_t stream);
int dm_sched_diffmusic_update_io(const float* x0, const void* eps, const void* g0, float leaf_scale, const float* z,
                                 void* prev, int n_clips, long long n_clip, float sqrt_a, float sqrt_p, float dir_coef,
                                 float std, float rate, float grad_scale, float e, float threshold, const float* coef,
                                 int io_dtype, const float* losses /* n_clips */, float* loss_total,
                                 dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * STFT / mel guidance  (operator.py:24-36,123-124 T_mel ; 153-154 phase mel ; 162-171 |STFT|)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct dm_stft_tables {
    const float* window;   /* [1024] hann (periodic) or ones (phase retrieval, operator.py:163-169 window=None) */
    const float* tw512;    /* [512][2] exp(-2 pi i m/512)  */
    const float* w1024;    /* [257][2] exp(-2 pi i k/1024) */
    const int* mel_kstart; /* [64] */
    const int* mel_klen;   /* [64] */
    const float* mel_w;    /* [mel_wstride][64] banded filterbank, transposed: mel_w[i*64+m] = fb[kstart[m]+i, m] */
    int mel_wstride;
    const int* bin_m0;     /* [513] */
    const float* bin_w0;   /* [513] */
    const float* bin_w1;   /* [513] */
    /* optional: shared-memory image of the warp-per-frame-pair kernel's tables (halved paired window, exchange
     * twiddles, per-lane bin-pair filterbank rows, per-bin filterbank), built on the host by
     * diffmusic_b200/tables.py warp_image() from the tables above; NULL selects the 64-thread frame-pair kernel. */
    const float* warp_image;
    int warp_image_floats; /* multiple of 4 */
    int warp_na, warp_nb;  /* bin-pair rows per lane for band l and band 63 - l */
} dm_stft_tables;

#define DM_STFT_MEL_DB 0    /* |X|^2 -> mel -> 10 log10(max(.,1e-10)) [-> clamp +-80]   (n_fft = win = 1024) */
#define DM_STFT_PHASE_MEL 1 /* |X|   -> mel [-> clamp +-80]                                                    */
#define DM_STFT_PHASE_WAV 2 /* |X|                                                                             */

/* Kernel selection for dm_stft_guidance (process-wide; for A/B measurements and tests).
 *   AUTO : the warp-per-frame-pair kernel (one warp = two frames as one 32 x 32 complex FFT, csrc/stft_warp.cu) when
 *          hop % 4 == 0 and frames_per_tile <= 16; else the 64-thread frame-pair kernel when hop is even; else the
 *          frame-at-a-time kernel.
 *   PAIR : the 64-thread frame-pair kernel (even hops).      FRAME: always the frame-at-a-time kernel (fp32 waveforms).
 * All three compute the same quantities (<= 1e-6 apart). */
#define DM_STFT_ENGINE_AUTO 0
#define DM_STFT_ENGINE_FRAME 1
#define DM_STFT_ENGINE_PAIR 2
int dm_stft_set_engine(int engine);

/* number of frame tiles per clip for a signal of Ly samples: ceil((1 + Ly/hop) / frames_per_tile) */
int dm_stft_num_tiles(long long Ly, int hop, int frames_per_tile);

/* One launch over B clips.  y: (B, Ly) with row stride y_bstride, optionally multiplied by mask[Ly] on load
 * (inpainting A(x) = x * mask, operator.py:132-133, fused into the frame load).
 *   transform mode : out != NULL, ref == NULL   -> out (B, R, T) = transform value, R = 64 (mel) or 513 (PHASE_WAV)
 *   guidance mode  : ref != NULL (row stride ref_bstride, 0 = shared by all clips), out == NULL
 *        partial[b * ntiles + tile] = sum over the tile of (ref - value)^2
 *        ypbar (B, Ly + 1024), may be NULL for loss only: accumulates (+=, caller zeroes) the UNSCALED cotangent of
 *        the reflect-padded signal, i.e. d(loss)/d(ypad) * loss ; the adjoint kernels below fold the padding
 *        and apply 1/loss per clip.
 *   noise (B, 513, T) / sigma: phase modes only, magnitude += sigma * noise (GaussianNoise on |STFT|,
 *        operator.py:171); NULL otherwise.
 * Replaces torch.stft + abs/pow + MelScale matmul + AmplitudeToDB + clamp + sub + linalg.norm and their autograd
 * replay (scheduling_dps.py:204-212). */
int dm_stft_guidance(const dm_stft_tables* tab, int mode, int clamp, int hop, const float* y, long long y_bstride,
                     long long Ly, const float* mask, int B, const float* ref, long long ref_bstride,
                     const float* noise, float sigma, float* out, float* ypbar, float* partial, int frames_per_tile,
                     dm_stream_t stream);

/* The mel-space guidance chain of SuperResolutionOperator at scale 2 with the resampling FUSED into the STFT kernel:
 * the same as dm_resample_fwd (orig 2, new 1, 28 taps, width 13: torchaudio Resample 16 kHz -> 8 kHz, operator.py:180,
 * 203-205 with run.py:188) followed by dm_stft_guidance(DM_STFT_MEL_DB, guidance mode) on y (B, Ly = ceil(L / 2)), but y
 * never exists in HBM: every CTA computes its tile's span of y from x (B, L) fp32 (row stride x_bstride, rows 16-byte
 * aligned) -- the first tile by all warps, the following ones by the warp that owns no frame pair, behind the
 * transforms of the current tile.  taps: [28] floats in HOST memory (row 0 of the resampling kernel; they travel as
 * kernel parameters and are read by the FMAs from the constant bank).
 * frames_per_tile <= 14, hop = 160.  ypbar (B, Ly + 1024) / partial (B, ntiles) as in dm_stft_guidance; the VJP
 * continues with dm_resample_adjoint.  Bit-identical to the unfused chain. */
int dm_stft_guidance_fir2(const dm_stft_tables* tab, int clamp, int hop, const float* x, long long x_bstride,
                          long long L, const float* taps, int B, const float* ref, long long ref_bstride, float* ypbar,
                          float* partial, int frames_per_tile, dm_stream_t stream);

/* out (B, 64, T) = clamp(mel filterbank applied to a materialised magnitude (B, 513, T), +-80)
 * (PhaseRetrievalOperator.transform, operator.py:153-154: MelScale matmul + clamp, no log). */
int dm_mel_project(const dm_stft_tables* tab, const float* mag, int B, long long T, int clamp, float* out,
                   dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Residual in the measurement ("wav_form") space and the adjoint / finalise family
 * ---------------------------------------------------------------------------------------------------------------- */
#define DM_RESID_CHUNK 4096
/* d = meas - y * (mask ? mask : 1); ybar = -d * (mask ? mask : 1) ... written UNSCALED to ybar (B, n);
 * partial[b * ntiles + c] = sum of d^2 over chunk c of DM_RESID_CHUNK elements.  (scheduling_dps.py:202-203,211) */
int dm_residual_wav(const float* y, long long y_bstride, long long n, int B, const float* mask, const float* meas,
                    long long meas_bstride, float* ybar, float* partial, dm_stream_t stream);

/* loss[b] = sqrt(sum partial[b, :]); dwav[b, j] = (1/loss[b]) * fold(ybar[b])[j] * (mask ? mask[j] : 1)
 * pad = 512: ybar is a reflect-padded cotangent (B, Ly + 1024) from dm_stft_guidance, folded back here
 *            (reflect adjoint); pad = 0: ybar is (B, Ly).
 * Serves identity / inpainting (operator.py:132-133 VJP) / phase retrieval.  dwav == NULL: loss only. */
int dm_fold_adjoint(const float* ybar, int pad, long long Ly, int B, const float* mask, const float* partial,
                    int ntiles, float* dwav, long long dwav_bstride, float* loss, dm_stream_t stream);

/* torchaudio sinc resampling as used by SuperResolutionOperator.forward (operator.py:180,203-205):
 * y[j*new + p] = sum_k xz[orig*j + k] * kernel[p, k], xz = zero-pad (width, width + orig), Ly = ceil(new*L/orig). */
int dm_resample_fwd(const float* x, long long x_bstride, long long L, int B, const float* kernel, int n_new,
                    int taps, int orig, int width, float* y, long long Ly, dm_stream_t stream);
/* VJP of the above fused with the reflect fold and the 1/loss scale. */
int dm_resample_adjoint(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                        const float* kernel, int n_new, int taps, int orig, int width, float* dwav,
                        long long dwav_bstride, long long L, float* loss, dm_stream_t stream);

/* Dereverberation (operator.py:244-250): y[i] = sum_k xz[i+k] ir[k], zero padding K/2, by overlap-save FFT.
 * tw4096: [4096][2] exp(-2 pi i m/4096); w8192: [2049][2] exp(-2 pi i k/8192).  K <= DM_RIR_MAX_TAPS.
 * spec: caller-owned (2 * 4097) floats, written by dm_rir_spectrum, read by the other two. */
#define DM_RIR_FFT 8192
#define DM_RIR_MAX_TAPS 6144
int dm_rir_spectrum(const float* ir, int K, const float* tw4096, const float* w8192, float* spec,
                    dm_stream_t stream);
int dm_rir_correlate(const float* x, long long x_bstride, long long L, int B, const float* spec, int K,
                     const float* tw4096, const float* w8192, float* y, long long Ly, dm_stream_t stream);
/* xbar[j] = (1/loss) * sum_k fold(ybar)[j + K/2 - k] ir[k]    */
int dm_rir_adjoint(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                   const float* spec, int K, const float* tw4096, const float* w8192, float* dwav,
                   long long dwav_bstride, long long L, float* loss, dm_stream_t stream);

/* y[b, j] = x[b, j] * mask[j]   (MusicInpaintingOperator.forward, operator.py:132-133) */
int dm_mask_apply(const float* x, long long x_bstride, long long L, int B, const float* mask, float* y,
                  dm_stream_t stream);

/* dst[i] = src[i] by a kernel; src may be PINNED HOST memory (read over PCIe through unified addressing).  For the
 * per-step impulse response of the dereverberation operator (operator.py:238-242, 20 KB drawn on the host): a copy-engine
 * transfer on the compute stream would queue behind the bulk latent uploads / downloads of a pipelined loop. */
int dm_copy_f32(float* dst, const float* src, long long n, dm_stream_t stream);

/* y += sigma * noise  (GaussianNoise.forward, noise.py:13-18, with the torch-drawn noise as an input) */
int dm_add_scaled(float* y, const float* noise, float sigma, long long n, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * FAD embedding statistics (fadtk/fad.py:41-47, fadtk/utils.py:13-46): raw moments of an (N, d) fp16 block,
 * accumulated (+=) into acc = [n | sum x (d) | sum x x^T (d*d)] in float64.  All-reducing acc over ranks (one NCCL
 * sum) and finalising mu = sx/n, cov = (sxx - n mu mu^T)/(n-1) equals the reference's Chan merge.
 * ---------------------------------------------------------------------------------------------------------------- */
int dm_fad_moments(const void* x_f16, long long N, int d, double* acc, dm_stream_t stream);
/* same with the engine for sum x x^T chosen explicitly: AUTO = tcgen05 + TMA when d % 8 == 0 and X is 16-byte aligned
 * (TMA requirements), else the SIMT tile kernel. */
#define DM_FAD_AUTO 0
#define DM_FAD_SIMT 1
#define DM_FAD_TCGEN05 2
#define DM_FAD_TCGEN05_PAIR 3 /* tcgen05.mma.cta_group::2: a CTA pair (two SMs) per 256 x 128 super-tile */
int dm_fad_moments_ex(const void* x_f16, long long N, int d, double* acc, int engine, dm_stream_t stream);
/* mu (d) and cov (d, d) in float64 from acc */
int dm_fad_finalize(const double* acc, int d, double* mu, double* cov, dm_stream_t stream);
/* the same from an accumulator whose sum x x^T holds (at least) the upper triangle -- what the exchanges below leave */
int dm_fad_finalize_sym(const double* acc, int d, double* mu, double* cov, dm_stream_t stream);

/* The one exchange of the path (fadtk/utils.py:36-40 across GPUs): sum acc over the ranks of a node.
 * Only [n | sum x | UPPER TRIANGLE of sum x x^T] crosses the links: dm_fad_packed_doubles(d) = 1 + d + d (d + 1) / 2.
 *
 * (a) over NVLink peer memory, no collective library: every rank's acc and a flag pad of dm_fad_flag_words() zeroed
 *     32-bit words live in memory the peers have mapped (e.g. CUDA IPC).  Per round (round = 1, 2, ... per call, the same
 *     on all ranks):  dm_fad_reset_shared (waits until every peer has read exchange round `done_round` -- the last
 *     one this accumulator took part in, 0 if none -- then clears acc) -> dm_fad_moments ... -> dm_fad_allreduce_peers (one kernel: flags READY to the peers, waits for theirs, sums
 *     every peer's acc in rank order -- bit-identical on all ranks -- into out_acc, flags DONE).  peer_acc /
 *     peer_flags: HOST arrays of `world` device pointers, entry r = rank r's buffers (entry `rank` = this rank's own).
 *     Nothing synchronises with the host.
 * (b) dm_fad_allreduce: through ncclAllReduce on a caller-owned communicator (ncclComm_t passed as void*; the symbol is
 *     resolved from the process at run time, the library does not link NCCL), in place on acc;
 *     packed_work: dm_fad_packed_doubles(d) doubles of scratch. */
int dm_enable_peer_access(int device, int peer_device); /* kernels on `device` may read / write `peer_device` memory */
/* Peer-visible buffers for the exchange above: zeroed device memory of this process (dm_peer_alloc), its 64-byte CUDA
 * IPC handle (dm_ipc_export; hand it to the other ranks of the node by any host channel), and the mapping of a peer's
 * handle into this process WITH `device` CURRENT (dm_ipc_open), so kernels on `device` can dereference the result. */
int dm_peer_alloc(int device, long long bytes, void** ptr);
int dm_peer_free(int device, void* ptr);
int dm_ipc_export(int device, const void* ptr, unsigned char* handle64);
int dm_ipc_open(int device, const unsigned char* handle64, void** ptr);
int dm_ipc_close(int device, void* ptr);
long long dm_fad_packed_doubles(int d);
int dm_fad_flag_words(void);
int dm_fad_reset_shared(double* acc, int d, const unsigned* my_flags, int world, unsigned done_round,
                        dm_stream_t stream);
int dm_fad_allreduce_peers(const double* const* peer_acc, unsigned* const* peer_flags, int world, int rank, int d,
                           unsigned round, double* out_acc, dm_stream_t stream);
/* (a') the default for W > 2: reduce-and-push ("two-shot").  Rank r reduces the triangle rows r, r + W, ... from every
 *     peer's acc and writes them into EVERY rank's sum buffer (peer_sum[r]: peer-visible, same layout as acc), so
 *     2 (W - 1) / W of the packed size crosses the links per GPU instead of (W - 1) times.  The sum is complete on a
 *     rank once every rank has raised DONE for the round: dm_fad_finalize_shared waits for that inside its kernel. */
int dm_fad_allreduce_push(const double* const* peer_acc, double* const* peer_sum, unsigned* const* peer_flags,
                          int world, int rank, int d, unsigned round, dm_stream_t stream);
int dm_fad_finalize_shared(const double* sum, int d, const unsigned* my_flags, int world, unsigned round, double* mu,
                           double* cov, dm_stream_t stream);
int dm_fad_allreduce(void* nccl_comm, double* acc, int d, double* packed_work, dm_stream_t stream);
int dm_fad_pack_tri(const double* acc, int d, double* packed, dm_stream_t stream);
int dm_fad_unpack_tri(const double* packed, int d, double* acc, dm_stream_t stream);

/* Workspace sizes in bytes (the library never allocates; callers size their buffers with this):
 *   DM_WS_STFT_COTANGENT (a = Ly, b = B), DM_WS_STFT_PARTIAL (a = Ly, b = B, c = frames per tile; hop 160),
 *   DM_WS_FAD_ACC / DM_WS_FAD_PACKED (c = d), DM_WS_FAD_FLAGS.  Returns -1 for invalid arguments. */
#define DM_WS_STFT_COTANGENT 0
#define DM_WS_STFT_PARTIAL 1
#define DM_WS_FAD_ACC 2
#define DM_WS_FAD_PACKED 3
#define DM_WS_FAD_FLAGS 4
long long dm_workspace_bytes(int kind, long long a, long long b, int c);


/* ------------------------------------------------------------------------------------------------------------------
 * 16-bit WAVEFORMS: in an fp16 / bf16 pipeline the vocoder output and the cotangent handed back
 * to autograd are 16-bit.  These variants read the waveform / write dLoss/dwav in that type (DM_IO_*), everything in
 * between stays fp32.  y_bstride / dwav_bstride are in ELEMENTS of the given type.
 * ---------------------------------------------------------------------------------------------------------------- */
int dm_stft_guidance_io(const dm_stft_tables* tab, int mode, int clamp, int hop, const void* y, int y_dtype,
                        long long y_bstride, long long Ly, const float* mask, int B, const float* ref,
                        long long ref_bstride, const float* noise, float sigma, float* out, float* ypbar,
                        float* partial, int frames_per_tile, dm_stream_t stream);
int dm_residual_wav_io(const void* y, int y_dtype, long long y_bstride, long long n, int B, const float* mask,
                       const float* meas, long long meas_bstride, float* ybar, float* partial, dm_stream_t stream);
int dm_fold_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* mask, const float* partial,
                       int ntiles, void* dwav, int dwav_dtype, long long dwav_bstride, float* loss, dm_stream_t stream);
int dm_resample_fwd_io(const void* x, int x_dtype, long long x_bstride, long long L, int B, const float* kernel,
                       int n_new, int taps, int orig, int width, float* y, long long Ly, dm_stream_t stream);
/* dm_resample_fwd_io that also sets fill[0 .. fill_count) = 0 (fill may be NULL): the padded cotangent buffer that the
 * dm_stft_guidance call of the same chain accumulates into -- folded into the resampling kernel where it is the
 * persistent scale-2 one (a memset node on the same stream otherwise), so the chain has no fill launch of its own. */
int dm_resample_fwd_fill_io(const void* x, int x_dtype, long long x_bstride, long long L, int B, const float* kernel,
                            int n_new, int taps, int orig, int width, float* y, long long Ly, float* fill,
                            long long fill_count, dm_stream_t stream);
int dm_resample_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                           const float* kernel, int n_new, int taps, int orig, int width, void* dwav, int dwav_dtype,
                           long long dwav_bstride, long long L, float* loss, dm_stream_t stream);
int dm_rir_correlate_io(const void* x, int x_dtype, long long x_bstride, long long L, int B, const float* spec, int K,
                        const float* tw4096, const float* w8192, float* y, long long Ly, dm_stream_t stream);
int dm_rir_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                      const float* spec, int K, const float* tw4096, const float* w8192, void* dwav, int dwav_dtype,
                      long long dwav_bstride, long long L, float* loss, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-clip step noise (diffmusic/torch_utils.py:31-76 `randn_tensor` with a LIST of generators: one torch.randn((1, ...),
 * generator=g_b) per clip).  One launch draws all clips with torch's own Philox4x32-10 / curand_normal4 index mapping, from
 * HOST arrays of each generator's (seed, offset) -- the values are bit-identical to the per-clip torch draws; the caller
 * advances every generator's offset by dm_randn_offset_increment(n_per_clip) afterwards.  out: (n_clips, n_per_clip) fp32;
 * round_dtype = DM_IO_F16 / DM_IO_BF16 rounds each value through that type first (what a 16-bit torch.randn returns).
 * ---------------------------------------------------------------------------------------------------------------- */
#define DM_RNG_MAX_CLIPS 128
long long dm_randn_offset_increment(long long n_per_clip);
int dm_randn_clips(const unsigned long long* seeds, const unsigned long long* offsets, int n_clips,
                   long long n_per_clip, int round_dtype, float* out, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Frechet distance and the FAD-inf bootstrap (fadtk/fad.py:50-119 `calc_frechet_distance`, :303-350 `score_inf`).
 * ---------------------------------------------------------------------------------------------------------------- */
/* out_f16[r, :] = x_f16[idx[r], :]  -- embeds[np.random.choice(N, n)] of score_inf (fad.py:331-332); idx is int64 on
 * the device (drawn by the caller with the reference's NumPy generator). */
int dm_fad_gather_rows(const void* x_f16, long long N, int d, const long long* idx, long long n, void* out_f16,
                       dm_stream_t stream);
/* Singular values (= eigenvalues for a symmetric PSD input) of the d x d row-major float64 matrix W by one-sided
 * Jacobi row rotations, in place (W ends with mutually orthogonal rows); eig[k] = |row k|.  state: 8 x uint64 of device
 * scratch; state[3] = sweeps executed.  All max_sweeps sweeps are enqueued (each one replay of a CUDA graph holding its d - 1
 * rounds; DM_JACOBI_GRAPH=0: one launch per round); they become no-ops once the largest normalised inner product of a
 * sweep is <= tol. */
int dm_sym_eig_jacobi(double* W, int d, int max_sweeps, double tol, double* eig, unsigned long long* state,
                      dm_stream_t stream);
/* doubles of device scratch for dm_frechet_distance */
long long dm_frechet_workspace_doubles(int d);
/* out[0] = |mu1-mu2|^2 + tr C1 + tr C2 - 2 tr sqrt(C1 C2), out[1] = tr sqrt(C1 C2), out[2..3] = Jacobi sweeps used.
 * mu*: (d), cov*: (d, d) row-major float64 on the device (what dm_fad_finalize writes). */
int dm_frechet_distance(const double* mu1, const double* cov1, const double* mu2, const double* cov2, int d,
                        int max_sweeps, double tol, double* work, double* out, dm_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Evaluation metrics (diffmusic/metrics/lsd.py:17-40, mse.py:9-29).
 * ---------------------------------------------------------------------------------------------------------------- */
/* out (B, T = 1 + L/hop): per-frame sqrt(mean over the 513 bins of (log10(|STFT ref| + eps) - log10(|STFT est| + eps))^2),
 * n_fft = win = 1024 with tab->window, centred frames, zero padding (librosa >= 0.10) or reflection.  `est` is always
 * sanitised like np.nan_to_num(nan=0, posinf=1, neginf=-1) (lsd.py:23), `ref` only if sanitize_ref.  hop even. */
int dm_lsd_frames(const dm_stft_tables* tab, const float* ref, long long ref_bstride, const float* est,
                  long long est_bstride, long long L, int B, int hop, int pad_reflect, int sanitize_ref, float eps,
                  float* out, dm_stream_t stream);
/* out[b] = mean_i (nan_to_num(ref[b, i]) - nan_to_num(est[b, i]))^2 over i < n; partial: B x dm_mse_num_chunks(n)
 * float64 of device scratch (fixed summation order: bit-reproducible). */
long long dm_mse_num_chunks(long long n);
int dm_mse(const float* ref, long long ref_bstride, const float* est, long long est_bstride, long long n, int B,
           double* partial, float* out, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * mel_spectrogram_to_waveform_with_phase (diffmusic/pipelines/pipeline_musicldm.py:263-301, plpeline_audioldm2.py:681):
 * torchaudio InverseMelScale(n_stft 513, n_mels 64, sr 16000) -- the minimum-norm least-squares solution, i.e. the fixed
 * matrix W = fb (fb^T fb)^-1, then relu -- times exp(i * phase), then torch.istft(n_fft 1024, hop, win 1024, window=None,
 * center=True), clipped / zero-padded to out_len.
 * mel: element (b, m, t) at mel[b * mel_bstride + m * mel_mstride + t * mel_tstride] (the pipeline's (B, 1, T, 64) tensor
 * is read in place: mstride 1, tstride 64); phase: (513, T) shared by the batch (phase_bstride 0) or (B, 513, T);
 * winv_t: (64, 513) = W^T; ola: dm_istft_workspace_floats(B, T, hop) floats of device scratch (zeroed by the call);
 * out: (B, out_len), out[b, j] = 0 for j >= hop (T - 1).  hop even. */
/* waveform_to_spectrogram (diffmusic/utils.py:11-20; run.py:305): torch.stft(n_fft 1024, hop, win 1024, window =
 * tab->window -- the reference passes none, i.e. rectangular --, center=True, reflect) -> mag = |X| and phase = angle(X),
 * each (B, 513, T = 1 + L/hop); either output may be null.  wav: (B, L) fp32 rows wav_bstride apart.  hop even. */
int dm_stft_spectrogram(const dm_stft_tables* tab, const float* wav, long long wav_bstride, long long L, int B, int hop,
                        float* mag, float* phase, dm_stream_t stream);
long long dm_istft_workspace_floats(int B, long long T, int hop);
int dm_istft_mel_phase(const dm_stft_tables* tab, const float* winv_t, const float* mel, long long mel_bstride,
                       long long mel_mstride, long long mel_tstride, const float* phase, long long phase_bstride, int B,
                       long long T, int hop, float* ola, float* out, long long out_len, dm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DM_ABI_H_ */
