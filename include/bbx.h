/*
 * bbx.h -- C ABI of libbbx.so: B200 (sm_100a) kernels for BlackBOX's per-frame CCD reduction
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer unless the name ends in _h (host);
 *   - images are row-major, C-contiguous; "raw" = frame with overscans (bbx_geom.H x W),
 *     "red" = reduced frame (ny*ysize_chan x nx*xsize_chan);
 *   - stream is a cudaStream_t passed as void*; calls enqueue work and return, they never
 *     synchronise the device unless documented;
 *   - return value 0 = ok, < 0 = error; bbx_last_error() describes the last failure of the
 *     calling thread.  Nothing here calls exit() or resets the device -- the Python shim
 *     raises RuntimeError so blackbox_reduce's per-step try/except keeps working
 *     (reference: blackbox.py:1476-1488, 1531-1594, 1750-1761, 1866-1878, 1897-1912);
 *   - no hidden device allocation: scratch space is passed in by the caller.
 *
 * Each entry point cites the reference code it replaces (file:line into the reference repo).
 */
#ifndef BBX_H
#define BBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBX_NCHAN 16
#define BBX_MAX_POLY_DEG 7

/* raw pixel types accepted by the kernels that read a raw frame */
#define BBX_RAW_U16 0
#define BBX_RAW_F32 1

/* telescope families: decides the horizontal-overscan masking rule (blackbox.py:6586-6643) */
#define BBX_TEL_ML 0
#define BBX_TEL_BG 1

/* Integer description of the channel layout; filled from define_sections
 * (blackbox.py:6334-6402) by blackbox_b200/geometry.py. */
typedef struct {
    int H, W;                     /* raw frame shape                                         */
    int ny, nx;                   /* channels in y (2) and x (8)                             */
    int dy, dx;                   /* channel tile incl. overscans                            */
    int ysize_chan, xsize_chan;   /* data section of one channel                             */
    int vos_x0, vos_w;            /* vertical-overscan strip: first column (tile-relative), width */
    int hos_rows;                 /* rows of the (cut) horizontal-overscan strip             */
    int data_y0_bot, data_y0_top; /* first raw row of the data section, bottom / top half    */
    int hos_y0_bot, hos_y0_top;   /* first raw row of the horizontal-overscan strip          */
} bbx_geom;

/* Internal marker bit of the uint8 mask: "this pixel was found saturated by the data >= level
 * test" (mask_sat of the reference, blackbox.py:4469-4498), as opposed to a 'saturated' bit a
 * bad-pixel mask may already carry.  Set by bbx_reduce_apply, consumed by
 * bbx_mask_sat_neighbours / bbx_count_objects, cleared by bbx_fill_sat_holes.  No mask value
 * may use it. */
#define BBX_TMP_SAT 0x80

/* mask bit values (set_zogy.mask_value; reference uses them at blackbox.py:4413-4596, 7171-7184) */
typedef struct {
    int bad, cosmic, saturated, satcon, sattrail, edge, crosstalk;
} bbx_maskbits;

const char *bbx_last_error(void);
int bbx_version(void);

/* ---------------------------------------------------------------------------------------
 * Overscan statistics and fits -- os_corr, blackbox.py:6407-6879
 * ------------------------------------------------------------------------------------- */

/* Row-wise sigma-clipped mean (sigma 3, <= maxiters, centre = mean, zeros masked) of the 16
 * vertical-overscan strips of a raw frame after the gain multiply.
 * Replaces blackbox.py:6480-6490 (+ gain_corr 7460 when raw is u16).
 * out_mean: float64 [16][dy], NaN for a fully masked row. */
int bbx_vos_rowstats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                     double sigma, int maxiters, double *out_mean, void *stream);

/* Per channel: 5-sigma clip of the row means, drop the rows overlapping the horizontal
 * overscan, least-squares polynomial of degree `deg` (<= 7) in the row index, evaluate on all
 * dy rows.  Replaces blackbox.py:6497-6556 (np.polyfit / np.polyval).
 * out_fit  float64 [16][dy]   value subtracted from every row of the channel tile
 * out_coef float64 [16][8]    ascending monomial coefficients (BIAS{i}A{n})
 * out_biasm float64 [16]      BIASM{i}
 * out_ok   int32 [16]         VFITOK{i} (0: fit not finite -> out_fit = nanmedian of the means)
 * smem budget: 3*dy doubles per block. */
int bbx_vos_fit(const double *mean_vos, const bbx_geom *g, int deg, double nsigma,
                double *out_fit, double *out_coef, double *out_biasm, int32_t *out_ok,
                void *stream);

/* BlackGEM: per channel and data column, count pixels >= 0.9*sat_e in the lim1 / lim2 data
 * rows nearest the horizontal overscan (values after gain and vertical-overscan subtraction).
 * Replaces blackbox.py:6624-6640.  out_cnt: int32 [16][2][xsize_chan], zeroed by the call. */
int bbx_hos_satcount(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                     const double *vos_fit, const double *sat_e_h, int lim1, int lim2,
                     int32_t *out_cnt, void *stream);

/* Horizontal-overscan strip of every channel: level offset dlevel (clipped mean of the last
 * 300 data columns), masking (ML: > data_limit with single-column un-masking and 5x5 growth;
 * BG: saturated columns from out_cnt), column-wise 2.5-sigma clipped mean / std(ddof 1) /
 * count.  Replaces blackbox.py:6565-6568, 6583-6662.
 * out_dlevel f64[16]; out_mean,out_std f32[16][xsize_chan]; out_n i32[16][xsize_chan];
 * out_satcol u8[16][xsize_chan] (BG; zeros for ML). */
int bbx_hos_stats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                  const double *vos_fit, int tel_kind, float data_limit,
                  const int32_t *satcnt, double *out_dlevel, float *out_mean, float *out_std,
                  int32_t *out_n, uint8_t *out_satcol, void *stream);

/* Sigma-clipped std (sigma 3, <= 5 iterations, zeros masked) of each vertical-overscan strip
 * after the vertical-overscan fit and dlevel were subtracted: RDN{i}.
 * Replaces blackbox.py:6572-6573.  out_std f64[16]. */
int bbx_vos_std(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                const double *vos_fit, const double *dlevel, double *out_std, void *stream);

/* Per channel: turn the column statistics into the overscan vector to subtract: errors,
 * 5-sigma pre-clean, 3x [polynomial fit of degree `deg` on columns >= 120, reject > 3 err],
 * BG2 channel 9 split fit, plain column mean where valid for the first 150 columns.
 * Replaces blackbox.py:6666-6678, 6727-6814 except the smoothing spline, which is only
 * needed for columns flagged in out_need_spline (host evaluates scipy's UnivariateSpline for
 * those; blackbox.py:6698-6723, 6795).
 * out_oscan f64[16][xsize_chan]; out_need_spline u8[16][xsize_chan]; out_status i32[16]
 * (0 ok, 1 = too few points for the fit). */
int bbx_hos_fit(const float *hos_mean, const float *hos_std, const int32_t *hos_n,
                const uint8_t *satcol, const bbx_geom *g, int tel_kind, int split_chan,
                int split_col, double *out_oscan, uint8_t *out_need_spline,
                int32_t *out_status, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused per-pixel pass -- gain_corr 7460, os_corr 6553/6556/6844-6847, master bias 1679,
 * mask_init 4408-4414 + 4494-4498 + 4538, master flat 1825
 *   v = f32(raw)*gain; v -= vos_fit[row]; v -= oscan[col]; crop; v -= mbias;
 *   non-finite -> 0 and 'bad' (if unmasked); saturated = v >= satlevel[chan] (float64
 *   compare) -> 'saturated'; v /= mflat.
 * Null mbias / mflat / bpm / out_mask / satlevel skip the corresponding step.
 * out_img f32 [red]; out_mask u8 [red].
 * seeds / seed_count (optional, device): the pixels found saturated are appended to seeds
 * (index into the reduced frame; bit 31 marks a pixel whose bad-pixel mask already carries a
 * saturated / saturated-connected bit); seed_count[0] is reset by the call and may end up
 * larger than seed_cap (overflow: the list is then incomplete).  Input of
 * bbx_mask_morph_sparse. */
int bbx_reduce_apply(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                     const double *vos_fit, const double *oscan, const float *mbias,
                     const float *mflat, const uint8_t *bpm, const double *satlevel,
                     const bbx_maskbits *bits, float *out_img, uint8_t *out_mask,
                     unsigned int *seeds, unsigned int *seed_count, unsigned int seed_cap,
                     void *stream);

/* bbx_reduce_apply AND the dense Laplacian scan of detect_cosmics' first iteration (the call at
 * blackbox.py:4323-4332 comes right after mask_init and the flat division) in ONE pass over the
 * frame: the reduced image is scanned for cosmic-ray candidates while it is being made, so
 * LACosmic does not read back the 446 MB the fused pass has just written.  Same arguments as
 * bbx_reduce_apply (out_mask required), plus those of bbx_lacosmic; the call also does what
 * bbx_lacosmic_begin does.  Afterwards: bbx_mask_morph_sparse_track (which corrects the background
 * statistics -- taken here against the seed mask -- for the pixels it masks), then
 * bbx_lacosmic_iteration(..., mode 3) for iter = 0 .. niter-1 with the SAME thresholds and work
 * buffer, bbx_lacosmic_finish(mode 0).  Bit-identical to the separate calls.  Requires the
 * 4-pixel-aligned layout (xsize_chan, dx, W multiples of 4; 16-byte aligned images). */
int bbx_reduce_apply_scan(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                          const double *vos_fit, const double *oscan, const float *mbias,
                          const float *mflat, const uint8_t *bpm, const double *satlevel,
                          const bbx_maskbits *bits, float *out_img, uint8_t *out_mask,
                          unsigned int *seeds, unsigned int *seed_count, unsigned int seed_cap,
                          uint8_t *crmask, float sigclip, float sigfrac, float objlim, float readnoise,
                          const double *readnoise_dev, int niter, void *lac_work, long long *lac_info,
                          void *stream);

/* bbx_reduce_apply that also takes the statistics of detect_cosmics' background level (count of
 * unmasked pixels, of those below a bracket, histogram inside it) while the values are in
 * registers, so that LACosmic's dense scan is the Laplacian alone (no mask read, no key
 * arithmetic).  Does what bbx_lacosmic_begin does.  Afterwards: bbx_mask_morph_sparse_track (corrects
 * the statistics for the pixels it masks), bbx_lacosmic_iteration(..., mode 4), bbx_lacosmic_finish
 * (mode 0).  Bit-identical to the separate calls.  Needs the 4-pixel-aligned layout.
 * niter < 0: ONLY the per-pixel kernel, with the bracket an earlier call left in lac_work (no
 * set-up kernels; the statistics keep accumulating) -- for timing that kernel on its own. */
int bbx_reduce_apply_stats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                           const double *vos_fit, const double *oscan, const float *mbias,
                           const float *mflat, const uint8_t *bpm, const double *satlevel,
                           const bbx_maskbits *bits, float *out_img, uint8_t *out_mask,
                           unsigned int *seeds, unsigned int *seed_count, unsigned int seed_cap,
                           int niter, void *lac_work, long long *lac_info, void *stream);

/* satlevel[i] = sat_e_h[i] - biasm[i] on the device (blackbox.py:4448-4454) */
int bbx_satlevels(const double *sat_e_h, const double *biasm, double *out_satlevel,
                  void *stream);

/* out[0] = BIASMEAN = nanmean(biasm[16]), out[1] = RDNOISE = nanmean(std_vos[16]), summed in
 * numpy's order (blackbox.py:6865-6868); device scalars so LACosmic can start without a host
 * round trip. */
int bbx_header_means(const double *biasm, const double *std_vos, double *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * Mask morphology -- mask_init blackbox.py:4473-4566, fill_sat_holes 4584-4596
 * ------------------------------------------------------------------------------------- */

/* crosstalk-victim bit from saturated pixels of the other 15 channels (y-mirrored across CCD
 * halves), then saturated-connected = 3x3 dilation of saturated minus saturated.
 * mask: u8 [red] in place. */
int bbx_mask_sat_neighbours(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                            const bbx_maskbits *bits, void *stream);

/* fill_sat_holes: closing (3x3, border 0) of saturated|saturated-connected, fill holes
 * (8-connected background), new pixels with mask==0 get 'saturated-connected'.
 * work: >= bbx_fill_holes_work_bytes(H, W) bytes.  `rounds` propagation rounds are enqueued;
 * *unconverged (device int32) is non-zero afterwards if more rounds are needed (call
 * bbx_fill_holes_more). */
size_t bbx_fill_holes_work_bytes(int H, int W);
int bbx_fill_sat_holes(uint8_t *mask, int H, int W, const bbx_maskbits *bits, void *work,
                       int rounds, int32_t *unconverged, void *stream);
int bbx_fill_holes_more(uint8_t *mask, int H, int W, const bbx_maskbits *bits, void *work,
                        int rounds, int32_t *unconverged, void *stream);

/* The whole of mask_init's morphology driven by the seed list of bbx_reduce_apply instead of
 * dense passes over the mask: crosstalk-victim and saturated-connected bits, NOBJ-SAT (8-
 * connected components of the saturated pixels), fill_sat_holes, clearing of BBX_TMP_SAT.
 * Bit-identical to bbx_mask_sat_neighbours + bbx_count_objects + bbx_fill_sat_holes.
 * work >= bbx_fill_holes_work_bytes(H, W); labels int32 [H*W] scratch (touched sparsely);
 * out_nobj device int32; status device int32[2]: status[0] bit 0 = seed list overflow, bit 1 =
 * hole propagation did not converge in `rounds` -- in both cases the mask is NOT final and the
 * dense entry points have to be run on a fresh mask; status[1] is scratch. */
int bbx_mask_morph_sparse(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                          const bbx_maskbits *bits, const unsigned int *seeds,
                          const unsigned int *seed_count, unsigned int seed_cap, void *work,
                          int32_t *labels, int32_t *out_nobj, int rounds, int32_t *status,
                          void *stream);

/* bbx_mask_morph_sparse with two optional extras.  After bbx_reduce_apply_scan: img = the reduced
 * image, lac_work = the LACosmic work buffer of that call (both may be null).  state_clean != 0: the
 * first H * W bytes of `work` (the hole-filling state image) are all ones -- set once by the
 * caller, left like that by every call of this function -- and the per-frame 111 MB memset is
 * skipped; bbx_fill_sat_holes / bbx_fill_holes_more do NOT leave them like that. */
int bbx_mask_morph_sparse_track(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                const bbx_maskbits *bits, const unsigned int *seeds,
                                const unsigned int *seed_count, unsigned int seed_cap, void *work,
                                int32_t *labels, int32_t *out_nobj, int rounds, int32_t *status,
                                const float *img, void *lac_work, int state_clean, void *stream);

/* number of 8-connected components of (mask & bit) != 0  (ndimage.label; blackbox.py:4354,
 * 4544).  labels: int32 [H*W] scratch; out_count device int32. */
int bbx_count_objects(const uint8_t *mask, int bit, int H, int W, int32_t *labels,
                      int32_t *out_count, void *stream);

/* per-bit pixel counts (mask_header, blackbox.py:4601-4620): out_counts int64 [8], bit k */
int bbx_mask_counts(const uint8_t *mask, size_t n, unsigned long long *out_counts,
                    void *stream);

/* ---------------------------------------------------------------------------------------
 * Crosstalk -- xtalk_corr blackbox.py:7138-7258
 * coeffs_h: float64 [16][16] indexed [source][victim].  img in place.
 * ------------------------------------------------------------------------------------- */
int bbx_xtalk(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
              const double *coeffs_h, const bbx_maskbits *bits, void *stream);
/* the same with the kernel chosen by the caller: variant 0 = bbx_xtalk's own choice (the tile kernel
 * when the layout allows), 3 = the tile kernel, 5 = the TMA-staged persistent kernel (measured
 * slower: the kernel is issue-bound), 1 / 2 / 4 = the generic register-only kernel with that many
 * pixels per thread (parity tests, benchmarks) */
int bbx_xtalk_variant(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                      const double *coeffs_h, const bbx_maskbits *bits, int variant, void *stream);
/* ... and with the per-bit pixel counts of the mask (mask_header, blackbox.py:4601-4620; the mask
 * is final when the crosstalk correction runs) taken on the way by the crosstalk kernel, which
 * sees every mask byte anyway: out_counts device uint64 [BBX_XTALK_COUNTS_LEN], [0:8] = the counts,
 * the rest scratch (spread partial counters); null: not wanted */
#define BBX_XTALK_COUNTS_LEN 136
int bbx_xtalk_counts(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                     const double *coeffs_h, const bbx_maskbits *bits, int variant,
                     unsigned long long *out_counts, void *stream);

/* ---------------------------------------------------------------------------------------
 * Master frames -- master_prep core blackbox.py:4908-4984 (+5063-5073)
 * frames_h: host array of N device pointers to float32 [npix]; scale_h: float32 [N] divisors
 * (0 = leave frame i unscaled; blackbox.py:4941).  np.median semantics: even N -> (a+b)/2 in
 * float32; any NaN -> NaN.  flat_fix != 0: pixels <= 0 or with bpm == edge become 1.
 * ------------------------------------------------------------------------------------- */
int bbx_stack_median(const float *const *frames_h, const float *scale_h, int N, size_t npix,
                     int flat_fix, const uint8_t *bpm, int edge_value, float *out,
                     void *stream);

/* The same with every master pixel stored to `ndst` (<= 8) buffers at once: outs_h = host array of
 * device pointers.  For the row-stripe sharded combine (north_star: "the master-frame combine shards
 * by row stripes ... NCCL only to assemble the stripes") the destinations are this GPU's slot of the
 * gather buffer and the same slot in every peer's buffer, mapped into this process over NVLink /
 * NVSwitch (torch symmetric memory): the combine kernel then IS the all-gather, and the transfer
 * runs under the N loads per pixel instead of after them.  multicast != 0: outs_h[0] (ndst = 1) is
 * an NVSwitch multicast address, written with multimem.st -- one store, replicated by the switch. */
int bbx_stack_median_multi(const float *const *frames_h, const float *scale_h, int N, size_t npix,
                           int flat_fix, const uint8_t *bpm, int edge_value, float *const *outs_h,
                           int ndst, int multicast, void *stream);

/* Optional sigma-clipped combine (BASELINE.json's wording; the reference's master_prep uses the
 * plain median above, so this is off by default in reduce.master_combine): per pixel
 * astropy.stats.sigma_clip(cube, sigma, maxiters, cenfunc='median', stdfunc='std', axis=0)
 * followed by np.ma.median -- (a+b)/2 in float32, NaN if nothing survives; non-finite inputs
 * never take part.  Same frame / scale / flat_fix arguments as bbx_stack_median. */
int bbx_stack_clipped_median(const float *const *frames_h, const float *scale_h, int N, size_t npix,
                             double sigma, int maxiters, int flat_fix, const uint8_t *bpm,
                             int edge_value, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * LACosmic -- astroscrappy.detect_cosmics 1.0.8 (sepmed=False, fsmode='median',
 * cleantype='medmask', gain=1, pssl=0, satlevel=inf), call site blackbox.py:4323-4332
 * img    f32 [H][W]  in: image, out: cleaned image
 * inmask u8  [H][W]  non-zero = excluded (may be null)
 * crmask u8  [H][W]  out: 0/1
 * work   >= bbx_lacosmic_work_bytes(H, W) bytes
 * readnoise_dev: optional device scalar (float64) used instead of `readnoise`
 * mode   0 = lazy: one dense Laplacian pass, then everything (medians, fine structure, growth,
 *            cleaning, the Laplacian of later iterations) only where it can matter
 *            (bit-identical to mode 1; requires sigclip >= 0 and sigfrac >= 0)
 *        1 = dense: every intermediate image is materialised
 *        2 = lazy, with the global background level (lower median of the unmasked input
 *            pixels) computed up front by three extra passes over image + mask
 *        3 = lazy, the dense pass already done by bbx_reduce_apply_scan (bbx_lacosmic_iteration
 *            only; bbx_lacosmic_begin is a no-op)
 *        4 = lazy, the background statistics already taken by bbx_reduce_apply_stats: the dense
 *            pass is the Laplacian alone (bbx_lacosmic_begin is a no-op)
 * out_info int64 [4 + niter] device: [0] iterations run, [1] internal, [2] status bits
 * (BBX_LAC_OVERFLOW: a work list overflowed -> result incomplete, repeat with mode 1;
 * BBX_LAC_NEED_BG (mode 0 only): a cosmic-ray pixel without any usable neighbour in its 5x5
 * box needs the background level -> result incomplete, repeat with mode 2),
 * [4+k] new CR pixels of iteration k.
 * The iteration loop runs on the device without host synchronisation; iterations after one
 * that found nothing are skipped (as the reference's `break`).
 * ------------------------------------------------------------------------------------- */
#define BBX_LAC_OVERFLOW 1
#define BBX_LAC_NEED_BG 2
size_t bbx_lacosmic_work_bytes(int H, int W);
int bbx_lacosmic(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                 float sigclip, float sigfrac, float objlim, float readnoise,
                 const double *readnoise_dev, int niter, int mode, void *work,
                 long long *out_info, void *stream);

/* The same in two parts, so one iteration can be enqueued (and timed) on its own:
 * _begin resets out_info and the work lists (mode 1 / 2: also computes the background level),
 * _iteration enqueues the kernels of iteration `iter` (a no-op on the device once an earlier
 * iteration found nothing).  crmask is valid after iteration 0: in modes 0 and 2 its dense scan
 * clears it (_begin does when niter == 0). */
int bbx_lacosmic_begin(const float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                       int niter, int mode, void *work, long long *out_info, void *stream);
int bbx_lacosmic_iteration(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                           float sigclip, float sigfrac, float objlim, float readnoise,
                           const double *readnoise_dev, int iter, int mode, void *work,
                           long long *out_info, void *stream);

/* After bbx_lacosmic (same mode, same work buffer): mask[crmask != 0] |= cosmic_bit
 * (blackbox.py:4349; mask may be null) and out_ncosmics = number of 8-connected cosmic-ray
 * objects (ndimage.label, blackbox.py:4354-4355).  mode 0 walks the cosmic-ray pixel list of
 * the lazy path (also for mode 2); mode 1 does dense passes.  labels: int32 [H*W] scratch. */
int bbx_lacosmic_finish(const uint8_t *crmask, uint8_t *mask, int cosmic_bit, int H, int W,
                        int mode, void *work, int32_t *labels, int32_t *out_ncosmics,
                        void *stream);

/* lower median a[(n-1)/2] of the pixels with inmask == 0 (astroscrappy's background level)
 * work >= bbx_select_work_bytes(); out device float32 */
size_t bbx_select_work_bytes(void);
int bbx_masked_lower_median(const float *img, const uint8_t *inmask, size_t n, void *work,
                            float *out, void *stream);

/* single LACosmic building blocks, exported for kernel-level parity tests */
int bbx_medfilt(const float *in, float *out, int H, int W, int ksize, void *stream);
int bbx_laplace_plus(const float *in, float *out, int H, int W, void *stream);

/* ---------------------------------------------------------------------------------------
 * Per-channel medians and the edge-pixel fill at the end of blackbox_reduce
 * (blackbox.py:1958-1974: for every channel, data[sec][mask_edge[sec]] = np.median(data[sec])).
 * bbx_channel_medians: out_med float32 [16] device = np.median of each channel of a reduced
 * frame (float32 mean of the two middle order statistics for an even pixel count; NaN if the
 * channel holds a NaN -- or, with ignore_nan != 0, np.nanmedian as get_flatstats uses it,
 * blackbox.py:3728-3733), by a three-pass radix select over all channels at once.
 * work >= bbx_chanmed_work_bytes().  bbx_fill_edge: img[(mask & edge_bit) != 0] = med[channel].
 * ------------------------------------------------------------------------------------- */
size_t bbx_chanmed_work_bytes(void);
int bbx_channel_medians(const float *img, int H, int W, int ysize_chan, int xsize_chan,
                        int ignore_nan, void *work, float *out_med, void *stream);
int bbx_fill_edge(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                  int edge_bit, const float *med, void *stream);

/* ---------------------------------------------------------------------------------------
 * Non-linearity correction -- nonlin_corr blackbox.py:7392-7437 (off in the reference's settings).
 * Per channel a FITPACK B-spline s (knots_h / coefs_h: float64 [16][max_knots], nknots_h, degree_h
 * <= 5; what scipy's UnivariateSpline._eval_args holds) gives the fractional deviation from
 * linearity as a function of the counts: img = f32(f64(img) / (frac + 1)) with
 * frac = (img / gain <= max_counts) ? s(img / gain) : 1 -- the reference's arithmetic, including
 * its division by 2 above the limit.  s is evaluated as FITPACK's splev does (bit-identical
 * float64 values).  img: reduced frame (2 x 8 channels), in place.
 * ------------------------------------------------------------------------------------- */
int bbx_nonlin_corr(float *img, int H, int W, int ysize_chan, int xsize_chan, const float *gain_h,
                    const double *knots_h, const double *coefs_h, const int *nknots_h,
                    const int *degree_h, int max_knots, float max_counts, void *stream);

/* ---------------------------------------------------------------------------------------
 * FITS data units on the device (the boundary either side of the path: raw frames come out of
 * read_hdulist, blackbox.py:1451 / 7653-7771, reduced images go into fits.writeto,
 * blackbox.py:1987-1990).  The host moves the file's bytes; these turn them into native arrays
 * and back.  n = number of pixels; buffers 16-byte aligned; in == out is allowed.
 * bbx_fits_decode: big-endian BITPIX 16 (unsigned16 != 0: BZERO 32768 -> uint16 counts, else
 * int16) or BITPIX -32 (float32) -> native.   bbx_fits_encode: the inverse.
 * ------------------------------------------------------------------------------------- */
int bbx_fits_decode(const void *be, int bitpix, int unsigned16, size_t n, void *out, void *stream);
int bbx_fits_encode(const void *in, int bitpix, int unsigned16, size_t n, void *out_be, void *stream);

/* Tile-compressed images (.fits.fz, ZCMPTYPE 'RICE_1', BLOCKSIZE 32, row tiles, BYTEPIX 1 / 2 / 4)
 * -- what read_hdulist (astropy / CFITSIO on the host in the reference) unpacks for every raw frame
 * (blackbox.py:1451), for the fpacked bad-pixel mask mask_init reads (blackbox.py:4386-4398,
 * Settings/set_blackbox.py:187-193) and for the fpacked reduced calibration frames master_prep
 * lists (blackbox.py:4698-4730).  heap: the binary table's heap (device, heap_bytes); offs / lens:
 * device arrays, one entry per tile = byte offset into the heap and compressed length (the
 * COMPRESSED_DATA column's descriptors); ntiles tiles of nx pixels, tile t = row t of out (uint8 /
 * int16 / int32 as BYTEPIX says).  flip_sign != 0: the sign bit of every pixel is inverted (BYTEPIX
 * 2 with BZERO 32768: stored int16 -> uint16 counts).  status: device int, zeroed by the call;
 * afterwards bit 0 = a tile ran past its bytes, bit 1 = a descriptor points outside the heap (the
 * affected rows are undefined); read it after synchronising.  No byte outside the heap is read. */
int bbx_rice_decode(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                    int ntiles, int nx, int blocksize, int bytepix, int flip_sign, void *out,
                    int *status, void *stream);
/* BYTEPIX 2 shorthand (raw frames) */
int bbx_rice_decode16(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                      int ntiles, int nx, int blocksize, int unsigned16, void *out, int *status,
                      void *stream);

/* Float images in a tile-compressed file are stored as scaled integers (ZQUANTIZ): this turns
 * the decoded int32 rows back into float32.  zscale / zzero: device float64 [ntiles] (the table's
 * ZSCALE / ZZERO columns); dither 0 = NO_DITHER: q * zscale + zzero; 1 / 2 = SUBTRACTIVE_DITHER_1 /
 * _2: (q - R[i] + 0.5) * zscale + zzero with R = rand10000 (device float32 [10000], the random
 * sequence the FITS standard defines for this purpose), restarted per tile from ZDITHER0;
 * dither 2: q == -2147483646 is an exact zero.  have_blank != 0: q == zblank -> NaN.
 * Reference side: read_hdulist(..., dtype='float32') of a `fpack -q` file (blackbox.py:826-840). */
int bbx_unquantize(const int32_t *q, int ntiles, int nx, const double *zscale, const double *zzero,
                   const float *rand10000, int dither, int zdither0, int zblank, int have_blank,
                   float *out, void *stream);

/* The other direction, for the product the reference fpacks losslessly (`fpack -D -Y` of the
 * uint8 mask, blackbox.py:826-827, 1990): img (device; ntiles rows of nx pixels of BYTEPIX 1 / 2 /
 * 4 bytes) -> out (device):
 *     [0:8)   int64  total heap bytes           [8:12) int32 ntiles
 *     [12:16) int32  status (bit 0: the heap did not fit into out_bytes; the sizes are still valid)
 *     [16 : 16 + 4 ntiles)  int32 compressed bytes of every tile (the descriptors of the
 *                            COMPRESSED_DATA column; offsets = their running sum), padded to 16
 *     then the heap: the tiles back to back, each the bytes fits_rcomp / _short / _byte produce.
 * work >= bbx_rice_encode_work_bytes(ntiles, nx, bytepix); bbx_rice_encode_out_bytes is the
 * out_bytes that always fits; a smaller out (the mask compresses 50-fold) is fine, status tells. */
size_t bbx_rice_encode_work_bytes(int ntiles, int nx, int bytepix);
size_t bbx_rice_encode_out_bytes(int ntiles, int nx, int bytepix);
int bbx_rice_encode(const void *img, int ntiles, int nx, int bytepix, void *work, size_t work_bytes,
                    void *out, size_t out_bytes, void *stream);

/* The reduced image itself leaves the reference as `fpack -q 16 -D -Y` (blackbox.py:826-836, the
 * float branch of fpack(), called from write_fits, blackbox.py:7677-7679): every row quantised to
 * integers of ZSCALE = (row noise) / q with subtractive dither, then Rice-coded (BYTEPIX 4).  This
 * does the same on the device -- CFITSIO's fits_quantize_float / FnNoise5_float restated in
 * oracle/rice.py (fpack_quantize_row) -- so that the compressed product is what crosses PCIe:
 * img (device float32, ntiles rows of nx <= 16384 pixels) -> out (device):
 *     [0:8)   int64 total heap bytes            [8:12) int32 ntiles
 *     [12:16) int32 status: bit 0 = the heap did not fit into out_bytes (sizes still valid);
 *             bits 8.. = number of rows NOT quantised (ZSCALE 0 and no Rice-coded bytes: zero noise,
 *             a span beyond 32 bits or a non-finite value; the caller stores those rows losslessly,
 *             fitsio.write_compressed gzips them into GZIP_COMPRESSED_DATA as CFITSIO would)
 *     int32 [ntiles] compressed bytes per tile, float64 [ntiles] ZSCALE, float64 [ntiles] ZZERO,
 *     each padded to a multiple of 16 bytes; then, at bbx_fpack_f32_heap_offset(ntiles), the heap.
 * qlevel: fpack's -q; zdither0: the ZDITHER0 keyword that goes with the file (1..10000; fpack
 * draws it from the clock); rand10000: as for bbx_unquantize.  work >= bbx_fpack_f32_work_bytes. */
size_t bbx_fpack_f32_work_bytes(int ntiles, int nx);
size_t bbx_fpack_f32_out_bytes(int ntiles, int nx);
size_t bbx_fpack_f32_heap_offset(int ntiles);
int bbx_fpack_f32(const float *img, int ntiles, int nx, float qlevel, int zdither0, const float *rand10000,
                  void *work, size_t work_bytes, void *out, size_t out_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * small elementwise helpers for the drop-in functions used one step at a time
 * ------------------------------------------------------------------------------------- */
/* gain_corr on a float32 raw frame in place (blackbox.py:7459-7460) */
int bbx_gain_corr(float *raw, const bbx_geom *g, const float *gain_h, void *stream);
/* a -= b (op 0) / a /= b (op 1), float32 (blackbox.py:1679, 1825) */
int bbx_binary_inplace(float *a, const float *b, size_t n, int op, void *stream);
/* mask[crmask != 0] |= bit  (blackbox.py:4349) */
int bbx_mask_or(uint8_t *mask, const uint8_t *flag, size_t n, int bit, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BBX_H */
