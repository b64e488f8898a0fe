/*
 * fcd_b200.h -- C ABI of the B200-native fast-checkerboard-demodulation (FCD) path.
 *
 * The reference (Trapped-modes-2025-LTG/Trapped-Modes-LTG, /root/reference) is pure Python
 * and has no FFI of its own; its boundary for this path is the classmethod surface of
 * pyfcd.  Each entry point below names the reference interface it
 * replaces.  All pointers named *_dev are CUDA device pointers on the device that was
 * current when the plan was created; `stream` is a cudaStream_t (NULL = default stream).
 * No torch types, no exceptions: every function returns 0 on success or a negative code,
 * and fcd_last_error() returns the message of the last failure on the calling thread.
 * A plan is bound to one device and may be used from one stream at a time.
 * Shapes: rows and cols powers of two in [64, 4096] give a FUSED plan (the float32 pipeline, fcd_execute, the
 * structure mask).  Any other shape with 2 <= rows, cols <= 2048 gives a GENERIC plan: the reference takes any shape
 * (pyfcd/fcd.py:14), and for those the float64 stage-level entry points (carrier search, bind, carrier mask / ccsgn,
 * fcd_fft2_c128 -- through Bluestein's chirp convolution on padded power-of-two transforms --, fcd_unwrap_phase,
 * fcd_count_residues) work and the host layer composes the per-frame path from them.
 */
#ifndef FCD_B200_H
#define FCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fcd_plan fcd_plan;

#define FCD_OK 0
#define FCD_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define FCD_ERR_RUNTIME (-2)   /* CUDA failure */
#define FCD_ERR_STATE (-3)     /* call order (e.g. execute before bind) */
#define FCD_ERR_NOPEAKS (-4)   /* no carrier candidate (reference: ValueError from min([]), fourier.py:38) */

/* Create a plan for rows x cols frames.  `frames_per_launch` frames are processed per kernel
 * wave (workspace is sized for that many); any batch size may be passed to fcd_execute.
 * Replaces: nothing in the reference (it is stateless and recomputes everything per call,
 * pyfcd/fcd.py:27); this is the hoisting of the per-reference work. */
int fcd_plan_create(int rows, int cols, int frames_per_launch, fcd_plan** out);
int fcd_plan_destroy(fcd_plan* plan);
const char* fcd_last_error(void);

/* fftshift(|fft2(image - mean)|) with k^2 <= (4 pi / min(shape))^2 zeroed, and its maximum.
 * Replaces pyfcd/fourier.py:18-23,34 (front half of fourier.find_peaks) and
 * pyfcd/fcd.py:149-155 (fcd.fft_peaks).  spectrum_dev: rows*cols float64 or NULL. */
int fcd_highpass_spectrum(fcd_plan* plan, const void* image_dev, int image_is_f64,
                          double* spectrum_dev, double* max_out, void* stream);

/* Threshold, clear the four border lines, 8-connected blobs, per-blob first maximum,
 * stable ascending sort by that maximum, first max_peaks.  Replaces
 * pyfcd/fourier.py:140-168 (fourier.find_peak_locations).  rc_out: 2*max_peaks ints
 * (row, col) on the host. */
int fcd_peak_locations(fcd_plan* plan, const double* image_dev, double threshold, int max_peaks,
                       int* rc_out, int* count_out, void* stream);

/* (rightmost, perpendicular) carrier pixels in shifted coordinates:
 * peaks_out = {r0, c0, r1, c1}.  Replaces pyfcd/fourier.py:8-41 (fourier.find_peaks). */
int fcd_find_peaks(fcd_plan* plan, const void* image_dev, int image_is_f64, int peaks_out[4], void* stream);

/* Per-reference state: disk masks, ccsgn = conj(ifft2(fft2(ref) * mask)) for both carriers,
 * carrier wavevectors and the folded integration coefficients.  Replaces
 * pyfcd/carriers.py:10-24 (Carrier.__init__) for both carriers plus the per-call constants
 * of pyfcd/fcd.py:123-138 and pyfcd/fourier.py:128-132.  peaks = {r0,c0,r1,c1} shifted. */
int fcd_bind_reference(fcd_plan* plan, const void* reference_dev, int reference_is_f64,
                       const int peaks[4], double radius, double calibration_factor,
                       double height, void* stream);

/* The per-frame path: frames_dev[n][rows][cols] float32 -> height_dev[n][rows][cols] float32
 * (+ optional phases_dev[n][2][rows][cols] float32).  Replaces pyfcd/fcd.py:28-33:
 * fft2(displaced) -> compute_phases (fcd.py:104-120, incl. unwrap_phase when unwrap != 0:
 * 1 = row/column scan, exact where the wrapped phases have no residues and the fast path;
 * 2 = reliability-guided like scikit-image, see fcd_unwrap_phase;
 * 3 = "auto", what the drop-in's unwrap=True runs: the scan path, during which the demodulation kernel flags
 * every frame that has |phase| > pi/2 somewhere -- only those can hold a 2*pi jump, let alone a residue; the
 * flags are read back once per call (4 bytes per frame), the flagged frames get their wrapped phases
 * materialised and their residues counted, and the frames that hold residues are redone exactly as mode 2 would; frames without residues keep the scan result, where every
 * unwrapper yields the same integers.  fcd_last_auto reports what happened)
 * -> compute_displacement_field (fcd.py:123-138) -> -u/height -> integrate_in_fourier
 * (fourier.py:116-137).  Optional mask_dev (uint8, nonzero = masked): the frame is replaced
 * by the reference under the mask before the transform and the height map is zeroed under
 * it afterwards (pydata/analyze.py:229-234,254-255); mask_stride = elements between
 * consecutive frames' masks (0: one mask for all frames). */
int fcd_execute(fcd_plan* plan, const float* frames_dev, int n_frames, float* height_dev,
                float* phases_dev, const uint8_t* mask_dev, long long mask_stride, int unwrap,
                void* stream);

/* Same as fcd_execute for camera frames that are still integers: frame_dtype 0 = float32,
 * 1 = uint8, 2 = uint16.  Replaces the `.astype(np.float32)` of analyze.load_image
 * (pydata/analyze.py:40) in front of the path: the widening happens in the first kernel's loads. */
int fcd_execute_typed(fcd_plan* plan, const void* frames_dev, int frame_dtype, int n_frames, float* height_dev,
                      float* phases_dev, const uint8_t* mask_dev, long long mask_stride, int unwrap,
                      void* stream);

/* After fcd_execute(..., unwrap = 3): number of frames flagged for a second look, number of frames redone with the
 * reliability-guided unwrap and (up to `capacity` of) their indices within that call.  Outputs may be NULL. */
int fcd_last_auto(const fcd_plan* plan, long long* flagged_out, int* guided_count_out, int* guided_frames_out,
                  int capacity);

/* Change the effective height (pyfcd/fcd.py:16-25, :32) of a bound plan without redoing the
 * per-reference work. */
int fcd_set_height(fcd_plan* plan, double height);

/* Residue guard: for each of n_maps phase maps (rows*cols float32 each, e.g. the phases output
 * of fcd_execute with unwrap = 0) the number of 2x2 loops whose wrapped differences do not sum
 * to zero.  The unwrapped phases equal scikit-image's unwrap_phase (pyfcd/fcd.py:119) up to a
 * global 2*pi*k only where this is zero.  counts_out: n_maps ints on the host. */
int fcd_count_residues(fcd_plan* plan, const float* phases_dev, int n_maps, int* counts_out, void* stream);

/* Reliability-guided 2-D phase unwrapping of n_maps maps (rows*cols float32 each, values in
 * [-pi, pi]); may run in place.  Replaces skimage.restoration.unwrap_phase as called at
 * pyfcd/fcd.py:119 (Herraez, Burton, Lalor, Gdeisat, Appl. Opt. 41, 7437 (2002)): pixel
 * reliability from wrapped second differences, neighbour pairs merged in order of the summed
 * reliabilities.  The merge order makes the result the integral of the wrapped differences
 * along the minimum spanning tree of that edge weighting, which is built here with Boruvka
 * rounds; output = input + 2*pi*integer, pixel (0, 0) keeps its value (scikit-image's result
 * differs by one global 2*pi*k).  This is what fcd_execute runs when unwrap == 2. */
int fcd_unwrap_phase(fcd_plan* plan, const float* wrapped_dev, int n_maps, float* unwrapped_dev, void* stream);

/* Temporal harmonic analysis of a device-resident stack of height maps maps_dev[n][rows][cols]
 * float32, for all spatial blocks at once: a grid of block_rows x block_cols blocks of
 * block_size^2 pixels from the top-left corner (the reference: block_size = rows // blocks_per_row,
 * block_rows = block_cols = blocks_per_row, analyze.py:572-573; a row band of a frame-sharded stack
 * holds fewer block rows).  Replaces analyze.block_amplitude / analyze.block_split
 * (pydata/analyze.py:542-641, 365-417), which re-read every map file once per block.
 *
 * fcd_temporal_mean_spectrum: np.nanmean(|np.fft.fft(block stack, axis=-1)|, axis=(0,1)) at the
 *   non-negative frequencies (analyze.py:603-614), pixels that are zero in first_map_dev excluded
 *   (analyze.py:568,590) and `zero` subtracted (analyze.py:585).  mean_out: host,
 *   [blocks][npos] float64 with npos = n/2 (n even) or (n+1)/2; valid_out: host, [blocks] valid-pixel
 *   counts (0 -> NaN row).  n must satisfy fcd_temporal_frames_supported: a power of two in [64, 4096] or at
 *   most 2048 (one chirp convolution), or a product of two factors <= 2048 (two levels; e.g. 20,000 = 160 x 125).
 * fcd_temporal_accumulate: adds frames [t0, t0 + n_chunk) of an n_total-frame series to the
 *   DFT sums of n_bins bins per block (bins: host, [blocks][n_bins]); acc_dev is
 *   [n_bins][2][rows*cols] float64 (init != 0 overwrites).  Any n_total; additive over chunks
 *   and over frame shards (sum the acc arrays of all ranks).
 * fcd_temporal_finalize: amplitude (|X|/n for the first bin, 2|X|/n after) and phase planes in
 *   the reference's (rows, cols, n_bins + 1) float64 layout (analyze.py:629-638), NaN where
 *   first_map_dev is zero. */
int fcd_temporal_mean_spectrum(fcd_plan* plan, const float* maps_dev, int n_frames, int rows, int cols,
                               const float* first_map_dev, float zero, int block_size, int block_rows, int block_cols,
                               double* mean_out, int* valid_out, void* stream);
int fcd_temporal_frames_supported(int n_frames);
/* The same with the two-level factorisation n = n1 * (n / n1) forced (long series take this path on their own:
 * X[k1 + n1 k2] from length-n1 transforms over the frames n1' * n2 + n2', a twiddle, and length-n2 transforms,
 * through an [n][pixels] complex workspace per chunk of spatial blocks); for tests of that path on short series. */
int fcd_temporal_mean_spectrum_split(fcd_plan* plan, const float* maps_dev, int n_frames, int rows, int cols,
                                     const float* first_map_dev, float zero, int block_size, int block_rows, int block_cols,
                                     int n1, double* mean_out, int* valid_out, void* stream);
int fcd_temporal_accumulate(fcd_plan* plan, const float* maps_dev, int n_chunk, int t0, int n_total, int rows, int cols,
                            float zero, int block_size, int block_rows, int block_cols, const int* bins, int n_bins,
                            double* acc_dev, int init, void* stream);
int fcd_temporal_finalize(fcd_plan* plan, const double* acc_dev, int n_bins, int n_total, int rows, int cols,
                          const float* first_map_dev, double* amps_dev, double* phases_dev, void* stream);

/* Floating-structure mask of each frame: box filter of width `smoothed`, threshold at the mean of
 * the filtered image, largest 8-connected region below it.  Bit-exact replacement of
 * analyze.mask (pydata/analyze.py:43-100).  frames_dev [n][rows][cols] float32 ->
 * mask_dev [n][rows][cols] uint8 (1 inside the structure). */
int fcd_structure_mask(fcd_plan* plan, const float* frames_dev, int n_frames, int smoothed, uint8_t* mask_dev,
                       void* stream);

/* Centre of the structure's cavity: int(centroid) of the largest 8-connected region of ~mask
 * whose bounding box does not touch the border.  Replaces analyze.center
 * (pydata/analyze.py:104-140).  centers_out: 2*n ints on the host, (cy, cx) per frame, or
 * (-1, -1) where the reference would fail for lack of an enclosed region. */
int fcd_mask_center(fcd_plan* plan, const uint8_t* mask_dev, int n_frames, int* centers_out, void* stream);

/* Carrier attributes (pyfcd/carriers.py:14-15): boolean mask in unshifted layout and ccsgn
 * as complex128 (as_c128 != 0) or complex64. */
int fcd_get_carrier_mask(fcd_plan* plan, int carrier, uint8_t* mask_dev, void* stream);
int fcd_get_carrier_ccsgn(fcd_plan* plan, int carrier, void* ccsgn_dev, int as_c128, void* stream);

/* Stage-level float64 building blocks used by the drop-in classmethods that take
 * user-supplied intermediates (fcd.compute_phases with an arbitrary displaced_fft,
 * fourier.integrate_in_fourier): in-place-capable 2-D complex128 FFT, direction -1 forward
 * / +1 inverse (inverse scaled by 1/(rows*cols) like scipy.fft.ifft2). */
int fcd_fft2_c128(fcd_plan* plan, const void* in_dev, void* out_dev, int direction, void* stream);

/* Per-stage device timing for benchmarks: when enabled, fcd_execute brackets every stage
 * with CUDA events on the launch stream.  Stages: 0 row-forward, 1 column band-pass,
 * 2 row demodulation, 3 row linking, 4 phase fix-up, 5 column integration, 6 row inverse.
 * fcd_stage_times synchronises the device and returns accumulated milliseconds, launches
 * and frames per stage since profiling was last enabled. */
int fcd_set_profiling(fcd_plan* plan, int enable);
int fcd_stage_times(fcd_plan* plan, double ms_out[7], long long launches_out[7], long long frames_out[7]);

/* Introspection for benchmarks: number of kernel launches issued through this plan and the
 * padded number of band columns per carrier (workspace geometry). */
long long fcd_launch_count(const fcd_plan* plan);
int fcd_band_columns(const fcd_plan* plan);
/* 1 for a fused plan (power-of-two shape: fcd_execute available), 0 for a generic one. */
int fcd_plan_is_fused(const fcd_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* FCD_B200_H */
