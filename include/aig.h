/*
 * aig.h - C ABI of libaig.so: the B200 (sm_100a) acoustic-image front end and
 * localisation scoring path of IIT-PAVIS/Acoustic-Image-Generation.
 *
 * The reference exposes this path as plain Python callables on NumPy arrays (the
 * only "plugin API" is tf.py_func(fn, [tensor], tf.float32),
 * dataloader/outdoor_data_mfcc.py:788).  Each entry point below replaces one of
 * those callables / inline blocks; the reference location is cited per function.
 * The Python drop-ins in acoustic_image_generation_b200/ bind these with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - Plain C: pointers and sizes only, no C++/torch types.
 *   - Every data pointer may be a DEVICE pointer (DLPack / __cuda_array_interface__ /
 *     torch.Tensor.data_ptr()) or a HOST pointer (NumPy; pinned memory makes the copies
 *     asynchronous).  The library classifies each pointer with
 *     cudaPointerGetAttributes and stages host buffers through its own device
 *     scratch; the caller owns every buffer, nothing allocated here crosses the ABI.
 *   - Every function returns 0 on success, a negative code on failure
 *     (AIG_ERR_* below, or -(1000 + cudaError_t) for CUDA runtime failures) and
 *     records a message retrievable with aig_last_error().  No C++ exception
 *     crosses the boundary.  There is no CPU fallback: without a CUDA device
 *     aig_create() fails.
 *   - A handle is one device + one stream.  Calls on one handle are serialised by
 *     the caller (the reference's tf.data map runs 4 worker threads,
 *     outdoor_data_mfcc.py:82: give each its own handle).  Calls return after
 *     the work has been enqueued when all buffers are device memory, and after
 *     completion when any buffer is host memory (NumPy semantics).
 */
#ifndef AIG_H_
#define AIG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIG_ABI_VERSION 2

#define AIG_OK 0
#define AIG_ERR_ARGUMENT (-1)    /* null / out-of-range / inconsistent argument        */
#define AIG_ERR_NO_DEVICE (-2)   /* no CUDA device, or device is not sm_100            */
#define AIG_ERR_TABLES (-3)      /* aig_set_tables not called, or unsupported geometry */
#define AIG_ERR_ALLOC (-4)       /* device / pinned allocation failed                  */
#define AIG_ERR_CUDA_BASE (-1000) /* -(1000 + cudaError_t)                             */

/* Frame geometry fixed by the reference (outdoor_data_mfcc.py:445 reshape [-1, 36, 48, 12];
 * iouenergythreshold.py:322 reshape (36, 48)). */
#define AIG_FRAME_H 36
#define AIG_FRAME_W 48
#define AIG_FRAME_PIXELS (AIG_FRAME_H * AIG_FRAME_W)
#define AIG_MFCC_NUM 12
#define AIG_FILTER_NUM 24

typedef struct aig_handle aig_handle;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Create a handle on CUDA device `device`.  `stream` is a cudaStream_t passed as an
 * integer (0 = the legacy default stream, which orders with torch's default stream;
 * pass torch.cuda.current_stream().cuda_stream to run on a torch stream).
 * Replaces: `with tf.device('/gpu:0')` / CUDA_VISIBLE_DEVICES selection
 * (iouenergythreshold.py:75, scripts/iou.bash:49). */
int aig_create(int device, uint64_t stream, aig_handle** out);
int aig_destroy(aig_handle* h);

/* Message of the last failure on this handle (or of the last failed aig_create when
 * h is NULL).  The pointer stays valid until the next call on the same handle. */
const char* aig_last_error(const aig_handle* h);

/* ABI version of the loaded library (== AIG_ABI_VERSION of the header it was built from). */
int aig_abi_version(void);

/* Hash of the sources and compiler flags the library was built from, stamped in by the build
 * (acoustic_image_generation_b200/_lib.py: source_build_id).  The Python loader refuses a library whose stamp differs
 * from the sources next to it, so a shipped binary is always the committed source. */
const char* aig_build_id(void);

/* Block until everything enqueued on the handle's stream(s) has finished. */
int aig_synchronize(aig_handle* h);

/* ---- tables --------------------------------------------------------------------------- */

/* Install the MFCC tables, all float64 host arrays, row-major:
 *   filter_mat [fft_len, filter_num]  from createfilters (outdoor_data_mfcc.py:826-849)
 *   dct_base   [filter_num, mfcc_num] (outdoor_data_mfcc.py:813-815)
 *   lifter     [mfcc_num]             (:816)
 *   mfnorm                            (:818)
 * The tables come from the caller so that parity never depends on re-deriving them
 * in C.  When they equal the reference configuration (512 bins, 24 filters, 12
 * coefficients, createfilters(512, 24, 0, 6400, 12800)) value for value, aig_mfcc
 * runs the fused banded kernel; any other tables run the generic float64 kernel
 * (fft_len <= 4096, filter_num <= 64, mfcc_num <= 32).
 * aig_tables_are_reference() reports which one is active (1 / 0, negative on error). */
int aig_set_tables(aig_handle* h, const double* filter_mat, int fft_len, int filter_num,
                   const double* dct_base, int mfcc_num, const double* lifter, double mfnorm);
int aig_tables_are_reference(const aig_handle* h);

/* ---- stage 1: spectra -> MFCC image ---------------------------------------------------- */

/* get_feats (outdoor_data_mfcc.py:851-876; copies at iouenergythreshold.py:325-350, ...)
 * followed by the np.float32 cast of its caller (:823):
 *   power    [n_rows, fft_len] float32  power spectra, one row per pixel (row-major)
 *   mfcc_out [n_rows, mfcc_num] float32
 * mel = power . filter_mat; floor at 0.001; ln; . dct_base; * mfnorm; * lifter; NaN/Inf -> 0.
 * flip180 != 0 additionally applies tf.image.flip_left_right + flip_up_down of
 * _parse_sequence (outdoor_data_mfcc.py:314-315) on store: rows are grouped in
 * frames of `frame_pixels` rows and row p of a frame is written to row
 * frame_pixels-1-p (n_rows must then be a multiple of frame_pixels). */
int aig_mfcc(aig_handle* h, const float* power, int64_t n_rows, float* mfcc_out,
             int flip180, int frame_pixels);

/* ---- stage 2: MFCC image -> energy map, mask, heat map --------------------------------- */

/* _normalize_acoustic_images_rescaled mapped over frames by _map_func_acoustic_images
 * (outdoor_data_mfcc.py:657-679): per frame, float32, x -= min(x); x /= max(x) over all
 * 36*48*12 values.  images / out: [n_frames, 36, 48, 12] float32 (may alias). */
int aig_normalize_images(aig_handle* h, const float* images, int64_t n_frames, float* out);

/* find_logen (iouenergythreshold.py:294-323) + the mean threshold mask
 * (iouenergythreshold.py:217-219), batched over frames.
 *   images          [n_frames, 36, 48, 12] float32 (12-channel acoustic image)
 *   normalize_first != 0: apply aig_normalize_images' arithmetic first (the dataloader
 *                   feeds find_logen with the normalised image)
 *   scaled_out      nullable [n_frames, 36, 48, 12] float32: the in-place side effect of
 *                   find_logen on its argument (x / lifter, then * mfnorm, each computed
 *                   in float64 and rounded to float32); may alias `images`
 *   energy_out      nullable [n_frames, 36, 48] float64: 1 / sum_j exp(sum_m x_m dct[j][m])
 *   mask_out        nullable [n_frames, 36, 48] uint8: energy > mean(energy), mean in
 *                   float64 with NumPy's pairwise summation order
 *   mean_out        nullable [n_frames] float64
 * Aliasing contract: scaled_out may be the very same buffer as images (find_logen scales its argument in place); no
 * other pair of buffers may overlap.  The kernels read images through ordinary (coherent) loads and load a pixel's
 * values before storing anything of that pixel.  Device image buffers must be 16-byte aligned.
 * Fewer frames than the device has SMs run one frame per thread-block cluster of 8 CTAs (see "small_batch_frames");
 * the results are bit-identical either way.  NaN in a frame propagates through the min / max of normalize_first like
 * tf.reduce_min / reduce_max: the whole frame's energies are NaN and its mask is empty. */
int aig_energy(aig_handle* h, const float* images, int64_t n_frames, int normalize_first,
               float* scaled_out, double* energy_out, uint8_t* mask_out, double* mean_out);

/* Heat-map rendering arithmetic: cv2.resize(map, (out_w, out_h)) (bilinear, half-pixel
 * centres, border clamp; showimages.py:147, showvideo.py:227) followed by the implicit
 * matplotlib Normalize of imshow (showimages.py:148): (x - min) / (max - min) with min/max
 * over the up-sampled image.
 *   energy   [n_frames, 36, 48] float64
 *   heat_out [n_frames, out_h, out_w] float32
 * A constant map gives NaN (0/0) like the reference.  See the "heatmap_exact" option for the two arithmetic modes. */
int aig_heatmap(aig_handle* h, const double* energy, int64_t n_frames, int out_h, int out_w,
                float* heat_out);

/* The per-frame body of showvideo.py:226-228 / showimages.py:146-148 (find_logen -> cv2.resize -> imshow
 * normalisation) for a batch, in ONE kernel launch: the float64 energy map goes from the find_logen warps to the
 * up-sampling and normalisation passes through shared memory and the heat map leaves the SM as bulk asynchronous
 * copies.  (Shapes with odd out_w or out_h * out_w not a multiple of 4, "heatmap_exact" mode and batches smaller than
 * the SM count run aig_energy's and aig_heatmap's kernels back to back instead - same arithmetic.)
 * energy_out / mask_out are nullable; heat_out [n_frames, out_h, out_w] float32 is required. */
int aig_energy_heatmap(aig_handle* h, const float* images, int64_t n_frames, int normalize_first, double* energy_out,
                       uint8_t* mask_out, float* heat_out, int out_h, int out_w);

/* 1.0 * (cv2.resize(mask * 1.0, (out_w, out_h)) > 0.5) (showimages_bb.py:303-304), decided in
 * exact integer arithmetic.  mask [n_frames, 36, 48] uint8 -> mask_up [n_frames, out_h, out_w] uint8. */
int aig_resize_mask(aig_handle* h, const uint8_t* mask, int64_t n_frames, int out_h, int out_w,
                    uint8_t* mask_up);

/* Stages 1 + 2 chained on the device (the north-star "MFCC + energy" pass):
 * power [n_frames, 36, 48, 512] float32 -> mfcc_out [n_frames, 36, 48, 12] float32 (flip180 as in
 * aig_mfcc), then aig_energy(normalize_first) on it.  mfcc_out must be non-null (device scratch is
 * NOT substituted: the MFCC image is a product of the pass); energy_out / mask_out / mean_out as in
 * aig_energy.  Requires the reference tables. */
int aig_mfcc_energy(aig_handle* h, const float* power, int64_t n_frames, int flip180,
                    int normalize_first, float* mfcc_out, double* energy_out, uint8_t* mask_out,
                    double* mean_out);

/* Stages 1 + 2 INCLUDING the heat map in the one persistent kernel (opt-in; the north star's "fused energy, upsample and
 * normalise" applied to the streaming pass): as aig_mfcc_energy, and the energy warps of every SM go on to up-sample and
 * normalise their frame's energy map (the arithmetic of aig_heatmap's default mode) and ship the rows with bulk
 * asynchronous copies while the next frame's spectra stream in.  heat_out [n_frames, out_h, out_w] float32, required.
 * Shapes the streaming heat-map code cannot take (odd out_w, out_h * out_w not a multiple of 4, > ~100 KiB of rows),
 * "heatmap_exact" mode and batches smaller than the SM count run the separate kernels instead - same arithmetic.
 * Host buffers are accepted but copied whole (no chunked overlap): the call is meant for device-resident streams. */
int aig_mfcc_energy_heatmap(aig_handle* h, const float* power, int64_t n_frames, int flip180, int normalize_first,
                            float* mfcc_out, double* energy_out, uint8_t* mask_out, double* mean_out, float* heat_out,
                            int out_h, int out_w);

/* ---- stage 3: scoring ------------------------------------------------------------------ */

/* The reference's whole ACIVW / AVIA evaluation step (iouenergythreshold.py:213-229) for a batch in one kernel launch:
 * real and reconstructed 12-channel image -> find_logen of each (on the np.stack copy, :216: the caller's arrays are not
 * scaled) -> mean masks -> I = sum(m & m2), U = sum(m | m2) -> iou -> pos[j] += iou > thr[j], num += 1 per frame.
 * Both energy maps and both masks stay in shared memory; they are written out only where a pointer is given.
 *   real, reconstructed  [n, 36, 48, 12] float32 (device buffers 16-byte aligned)
 *   normalize_first      as in aig_energy (0 for the reference's call)
 *   thr [k], inter_out / union_out nullable [n], pos_inout [k], num_inout [1]: as in aig_iou_sweep (accumulating)
 *   energy_*_out nullable [n, 36, 48] float64, mask_*_out nullable [n, 36, 48] uint8
 * Batches smaller than the SM count (the reference's are 2-16 frames) split every frame pair over a cluster of 8 CTAs. */
int aig_acivw_batch(aig_handle* h, const float* real, const float* reconstructed, int64_t n_frames, int normalize_first,
                    const double* thr, int k, int64_t* inter_out, int64_t* union_out, int64_t* pos_inout,
                    int64_t* num_inout, double* energy_real_out, double* energy_recon_out, uint8_t* mask_real_out,
                    uint8_t* mask_recon_out);

/* ACIVW / AVIA IoU and success counts (iouenergythreshold.py:224-229), all thresholds in one pass
 * (the reference re-runs the evaluation once per threshold, scripts/iou.bash:47-53).
 *   mask_a, mask_b [n, 36*48] uint8 (non-zero = set)
 *   thr            [k] float64 host or device
 *   inter_out, union_out nullable [n] int64: sum(a & b), sum(a | b)
 *   pos_inout      [k] int64: pos[j] += #frames with (double)I/(double)U > thr[j] (strict; U == 0
 *                  gives NaN and never counts).  ACCUMULATES, so shards/batches can be chained;
 *                  zero it before the first call.
 *   num_inout      [1] int64: += n */
int aig_iou_sweep(aig_handle* h, const uint8_t* mask_a, const uint8_t* mask_b, int64_t n,
                  const double* thr, int k, int64_t* inter_out, int64_t* union_out,
                  int64_t* pos_inout, int64_t* num_inout);

/* The same for a stream of fixed-length clips (BASELINE configs[3]: 10 s clips, 120 frames at the reference's 12 fps,
 * showvideo.py:139, or 300 at 30 fps): frame f belongs to clip f / frames_per_clip; pos_inout is [n_clips, k] row-major
 * (n_clips = ceil(n / frames_per_clip)) and accumulates per clip, so every clip's success curve and AUC come out of one
 * launch.  The per-clip frame count is frames_per_clip (the last clip may be shorter). */
int aig_iou_sweep_clips(aig_handle* h, const uint8_t* mask_a, const uint8_t* mask_b, int64_t n, int64_t frames_per_clip,
                        const double* thr, int k, int64_t* inter_out, int64_t* union_out, int64_t* pos_inout);

/* FlickrSoundNet consensus IoU and success counts (showimages_bb.py:288-321).
 *   mask  [n, 36*48] uint8 predicted mask at acoustic resolution (up-sampled here exactly as
 *         aig_resize_mask does)
 *   xmin, xmax, ymin, ymax [n, 3] int32 annotator boxes in out_w x out_h pixel coordinates,
 *         inclusive corners, clipped; a box is present iff xmax != 0 (dataloader/frames.py:290-299)
 *   inter2_out, union2_out nullable [n] int64: TWICE the reference's weighted intersection / union
 *         (the weights are multiples of 0.5, so the doubled sums are exact integers)
 *   pos_inout, num_inout as in aig_iou_sweep, with iou = (double)I2/(double)U2. */
int aig_ciou_sweep(aig_handle* h, const uint8_t* mask, const int32_t* xmin, const int32_t* xmax,
                   const int32_t* ymin, const int32_t* ymax, int64_t n, int out_h, int out_w,
                   const double* thr, int k, int64_t* inter2_out, int64_t* union2_out,
                   int64_t* pos_inout, int64_t* num_inout);

/* ---- callers either side of the path ("next" rows N1, N2 of SURVEY.md section 8(f)) ----------------
 *
 * aig_power_spectrum: front half of _build_spectrograms_function (outdoor_data_mfcc.py:796-805; the variant
 * without a window is dataloader/frames.py:659-667): rows of 1024 audio samples (int32 when audio_is_int32,
 * else float32) are multiplied by `window` (float64[1024], host or device; NULL = no window), transformed with
 * a 1024-point real FFT in float64, the Nyquist bin is dropped and the squared magnitude is written as
 * float32 power[n_rows, 512] - the input of aig_mfcc.
 *
 * aig_filtfilt: butter_lowpass_filter (outdoor_data_mfcc.py:565-575) = scipy.signal.filtfilt(b, a, x) along rows
 * (padtype 'odd', padlen 3*ntaps, lfilter_zi initial state), float64 arithmetic in scipy's operation order,
 * result cast to float32.  b, a: float64[ntaps] (ntaps <= 16), zi = scipy.signal.lfilter_zi(b, a): float64[ntaps-1],
 * all host arrays; x: [n_rows, length] int32 or float32 with length > 3*ntaps.
 *
 * aig_normalize_mfcc: _normalize_mfcc mapped by _map_func_mfcc (outdoor_data_mfcc.py:681-703): per 12-vector
 * float32 (x - min) / max(x - min).
 *
 * aig_tile_mfcc: mfccmap = tile(reshape(mfcc, (-1,1,12)), (1, 36*48, 1)) -> [n, 36, 48, 12]
 * (trainer/mfcctrainer.py:38-40, iouenergythreshold.py:99-101), optionally normalising each vector first. */
int aig_power_spectrum(aig_handle* h, const void* audio, int audio_is_int32, int64_t n_rows,
                       const double* window, float* power_out);
/* The whole _build_spectrograms_function (outdoor_data_mfcc.py:796-824) in one call: aig_power_spectrum followed by
 * aig_mfcc with the handle's tables, the [n_rows, 512] power spectra staying on the device.  mfcc_out: float32
 * [n_rows, mfcc_num]. */
int aig_audio_mfcc(aig_handle* h, const void* audio, int audio_is_int32, int64_t n_rows, const double* window,
                   float* mfcc_out);
int aig_filtfilt(aig_handle* h, const void* x, int x_is_int32, int64_t n_rows, int length, const double* b,
                 const double* a, const double* zi, int ntaps, float* y_out);
int aig_normalize_mfcc(aig_handle* h, const float* mfcc, int64_t n, float* out);
int aig_tile_mfcc(aig_handle* h, const float* mfcc, int64_t n, int normalize, float* map_out);

/* The channel-triplet slices of the trainers (trainer/mfcctrainer.py:105-112):
 * aig_split_triplets writes tf.slice(images, [0,0,0,3t], [-1,36,48,3]) for t = 0..3 as four contiguous
 * [n_frames, 36, 48, 3] float32 tensors, back to back in triplets_out (same total size as the input).
 * aig_triplet_mse reads a target and a generated image once and returns the five losses of
 * trainer/mfcctrainer.py:103,114-117: mse_out[0] = tf.losses.mean_squared_error over the whole image,
 * mse_out[1 + t] = over triplet t; squared differences in float32, summed in float64 in a fixed order. */
int aig_split_triplets(aig_handle* h, const float* images, int64_t n_frames, float* triplets_out);
int aig_triplet_mse(aig_handle* h, const float* a, const float* b, int64_t n_frames, double* mse_out);

/* Heat-map overlay rendering ("next" row N4; showvideo.py:217-233, showimages.py:144-150):
 *   heat     [n, out_h, out_w] float32 in [0, 1] (aig_heatmap's output)
 *   bgr      nullable [n, out_h, out_w, 3] uint8 video frames in OpenCV's BGR order
 *   jet_lut  [256, 3] uint8 host array: matplotlib's jet table (tables.jet_lut())
 *   rgb_out  [n, out_h, out_w, 3] uint8
 * gray = cv2.cvtColor(frame, COLOR_BGR2GRAY), min/max normalised and quantised to 256 levels (imshow(cmap=gray));
 * colour = jet_lut[min(int(heat * 256), 255)]; out = round(alpha * colour + (1 - alpha) * gray) (imshow(cmap=jet,
 * alpha=0.7)).  matplotlib's figure-resolution resampling and Agg's 8-bit compositing are not reproduced: the
 * overlay is produced at the heat map's own resolution. */
int aig_overlay(aig_handle* h, const float* heat, const uint8_t* bgr, int64_t n_frames, int out_h, int out_w,
                float alpha, const uint8_t* jet_lut, uint8_t* rgb_out);

/* ---- on-disk format ("next" row N3) --------------------------------------------------------------
 * Reader for the reference's data files: GZIP (or plain) TFRecord files of tf.train.SequenceExample records
 * (writer convert_data.py:247-279; parsers dataloader/outdoor_data_mfcc.py:260-344, dataloader/frames.py:246-341).
 * Host-side C++ (zlib + CRC-32C + a minimal protobuf wire walk), no TensorFlow.  Replaces
 * tf.data.TFRecordDataset(compression_type='GZIP') + tf.parse_single_sequence_example + tf.decode_raw.
 *   aig_records_open           inflate the file, verify every record's masked CRC-32C, index the records
 *   aig_records_count          number of records (-1 for a null reader)
 *   aig_record_context_int64   int64 context feature (e.g. "classes", "audio_image/height", "xmin"): up to
 *                              `capacity` values are written, *count_out receives how many the feature holds
 *   aig_record_sequence_size   feature list of byte strings (e.g. "audio/image"): steps and total payload bytes
 *   aig_record_sequence_read   the steps' byte strings concatenated into dst (host memory) - what
 *                              tf.decode_raw + reshape([-1, H, W, D]) sees
 *   aig_records_last_error     message of the calling thread's last reader failure
 *   aig_crc32c                 CRC-32C (Castagnoli) of a buffer, the checksum of the TFRecord framing (unmasked); uses the
 *                              SSE4.2 instruction when the CPU has it, force_table = 1 selects the portable table loop */
typedef struct aig_record_reader aig_record_reader;
int aig_records_open(const char* path, aig_record_reader** out);
int aig_records_close(aig_record_reader* r);
int64_t aig_records_count(const aig_record_reader* r);
int aig_record_context_int64(const aig_record_reader* r, int record, const char* key, int64_t* values_out,
                             int capacity, int* count_out);
int aig_record_sequence_size(const aig_record_reader* r, int record, const char* key, int64_t* steps_out,
                             int64_t* bytes_out);
int aig_record_sequence_read(const aig_record_reader* r, int record, const char* key, void* dst, int64_t dst_bytes);
const char* aig_records_last_error(void);
uint32_t aig_crc32c(const void* data, size_t n, int force_table);

/* ---- multi-GPU: the path's only exchange ------------------------------------------------------
 * Frames shard across GPUs with no data-path collective; at the end of an evaluation the int64[K+1]
 * count vector (pos[0..K-1], num) is summed over ranks.  The reference has no counterpart (single GPU,
 * one whole run per threshold, scripts/iou.bash:47-53).  NCCL is resolved at run time with
 * dlopen("libnccl.so.2") - in a torch process that is torch's bundled NCCL - so libaig has no link-time
 * dependency on it; without NCCL these calls return AIG_ERR_NO_DEVICE.
 *   aig_comm_unique_id   rank 0 creates the 128-byte NCCL unique id; the host program ships it to the other
 *                        ranks (torch.distributed broadcast, a file, MPI ...)
 *   aig_comm_init        every rank: join the communicator on the handle's device
 *   aig_allreduce_counts in-place sum of n int64 values over all ranks, enqueued on the handle's stream
 *                        (device buffer: asynchronous, ordered after the sweeps that filled it; host buffer:
 *                        staged and synchronous).  With no communicator (single rank) it is the identity.
 *                        A one-rank communicator (world == 1) is legal and still goes through ncclAllReduce.
 *   aig_comm_destroy     leave the communicator (also done by aig_destroy)
 * One process driving several GPUs - the shape of the reference's callers, which are single processes
 * (iouenergythreshold.py:140-236, scripts/iou.bash:47-53):
 *   aig_comm_init_all           one communicator over n handles of this process, one per distinct device
 *                               (ncclCommInitAll); handle i becomes rank i
 *   aig_group_allreduce_counts  the all-reduce for all n handles in one grouped NCCL call (ncclGroupStart / End) from one
 *                               host thread; counts[i] is handle i's DEVICE buffer of n int64, summed in place on
 *                               handle i's stream */
#define AIG_COMM_ID_BYTES 128
int aig_comm_unique_id(uint8_t* id_out);
int aig_comm_init(aig_handle* h, const uint8_t* id, int rank, int world);
int aig_allreduce_counts(aig_handle* h, int64_t* counts, int n);
int aig_comm_destroy(aig_handle* h);
int aig_comm_init_all(aig_handle** handles, int n);
int aig_group_allreduce_counts(aig_handle** handles, int64_t* const* counts, int n_handles, int n);

/* areaundercurve.py:32-37: sklearn.metrics.auc on the reversed (descending) threshold / success-rate
 * arrays == direction * trapezoid.  Host-side, float64.  thr / value are host arrays of length k. */
int aig_auc(const double* thr, const double* value, int k, double* auc_out);

/* ---- instrumentation and tuning --------------------------------------------------------- */

/* Number of kernels this handle has launched since creation (for bench.py's gpu_launches). */
int64_t aig_launch_count(const aig_handle* h);

/* Integer options, by name:
 *   "mfcc_variant"       ring geometry of the fused MFCC kernel (see aig_api.cu kVariants); -1 = default
 *   "chain_mode"         how aig_mfcc_energy runs its two stages: 2 (default) one fused persistent kernel
 *                        (fused_kernel.cuh); 1 two kernels overlapped on two streams; 0 two kernels in sequence
 *   "fused_variant"      ring geometry of the fused kernel (0..2)
 *   "chain_energy_ctas_per_sm" footprint of the overlapped energy kernel in mode 1 (default 3)
 *   "chain_chunk_frames" frames per launch pair in modes 0/1 (default 512: the chunk's MFCC
 *                        images, 42 MB, stay in the 126 MB L2 between the two kernels)
 *   "chain_overlap"      1 (default): the energy kernel of chunk i runs on a second stream while the
 *                        MFCC kernel of chunk i+1 streams from HBM; 0: both on the handle's stream
 *   "launch_row_limit"   spectra per kernel launch (default 2^31 - 1024: TMA tile coordinates are int32, larger batches
 *                        are split into several launches); lowering it exercises that split on small inputs
 *   "heatmap_exact"      aig_heatmap: 0 (default) float32 bilinear on the frame-normalised map (error ~1e-7, HBM-write
 *                        bound); 1 the float64 replica of cv2.resize + Normalize (bit-equal to the NumPy oracle)
 *   "keep_mfcc_in_l2"    fused kernel: 1 (default) L2 evict-last hint on the MFCC stores, so the energy warps'
 *                        read-back one frame later still hits L2; 0: plain stores
 *   "l2_evict_first"     1: L2 evict-first cache hint on the TMA spectrum loads; 0 (default): normal policy
 *   "host_copy_threads"  copies of 8 MiB (uploads: staged_min_bytes) and more from / to ordinary (pageable) host arrays are staged through
 *                        a ring of pinned 4 MiB slots by this many host threads (host_staging.h; 4-5x the driver's own
 *                        pageable path): -1 (default) min(6, hardware threads / 2); 0 leaves them to cudaMemcpyAsync
 *   "staged_min_bytes"   the size from which pageable uploads take that ring (default 8 MiB, from 65536).  Lowering it
 *                        to 1 MiB makes the reference's evaluation step from NumPy arrays (two arrays of 1.3 MB for a
 *                        batch of 16) 188 us instead of 274 us in the median, but bursts of 5-8 ms calls were measured
 *                        when the copy threads had been asleep (tools/add_batch_outlier_probe.py), so it is not the default
 *   "staged_small_piece_bytes"  piece size of staged uploads under 4 MiB (default 512 KiB; 64 KiB .. 4 MiB)
 *   "staged_solo_bytes"  uploads under this size are staged by the calling thread alone, the copy threads stay asleep
 *                        (default 0: none; measured equal to the driver's own pageable path, 266 vs 259 us)
 *   "host_copy_streaming" how the staging threads fill the pinned slots: 1 non-temporal stores (host_copy.cpp; needs AVX2:
 *                        a slot only the copy engine will read no longer evicts the caller's array from the host caches),
 *                        0 memcpy, -1 (default) non-temporal stores for uploads up to 128 MiB (measured: 20 % faster
 *                        there, 7 % slower on 900 MB jobs where the ring itself lives in the cache)
 *   "small_host_bytes"   host buffers up to this size (default 131072) are not copied with cudaMemcpy at all: they pass
 *                        through a pinned, device-mapped 1 MiB arena of the handle that the kernels read and write
 *                        directly over PCIe (zero-copy), which halves the latency of one-frame calls such as the
 *                        find_logen drop-in; 0 switches the path off
 *   "heat_bulk_store"    1 (default): heat maps are staged in shared memory and written with bulk asynchronous copies
 *                        (heat_stream_kernel); 0: the round-1 kernel with per-thread stores, for comparison runs
 *   "energy_wide"        1 (default): aig_energy and aig_acivw_batch on batches of at least small_batch_frames frames run eight
 *                        frames (four pairs) per CTA of 512 threads around one conflict-free (eight-copy) exponential table,
 *                        one CTA per SM; 0: one CTA of 64 (128) threads per frame (pair), eight (four) per SM, each with a
 *                        one-copy table (the kernel the chained modes overlap with the MFCC kernel)
 *   "acivw_wide_pairs"   aig_acivw_batch takes that form from this many pairs up (default 8192; below, four 128-thread CTAs per
 *                        SM balance a partial last round better)
 *   "overlay_luma"       1 (default): aig_overlay with a frame of up to 81 920 pixels (a multiple of four, aligned buffers) keeps
 *                        the frame's luma plane in shared memory between its min/max pass and its blend pass, so the BGR
 *                        frame is read once; 0: the BGR frame is read (and its luma computed) in both passes, as for larger frames
 *   "energy_heat_ws"     1 (default): aig_energy_heatmap runs the warp-specialised kernel (float64 warps and heat-map warps
 *                        of a CTA working on consecutive frames, two CTAs per SM); 0: the same warps do both phases in sequence
 *   "mask_packed"        1 (default): aig_resize_mask and aig_ciou_sweep at the reference's output sizes (224 x 298, 224 x 224)
 *                        run the packed kernels (two pixels per 32-bit multiply-add, compile-time taps, bulk-copy stores /
 *                        rectangle counts); 0: the generic kernels that serve every other size, for comparison runs
 *   "norm_bulk_copy"     1 (default): aig_normalize_images keeps each frame in shared memory between one bulk asynchronous
 *                        load and one bulk store (normalize_bulk_kernel); 0: the two-pass per-thread kernel, for comparison
 *   "small_batch_frames" batches with fewer frames than this spread each frame over a cluster of 8 CTAs (aig_energy,
 *                        aig_acivw_batch) and run aig_mfcc_energy as tiled MFCC kernel + cluster energy kernel instead of
 *                        the one-CTA-per-frame persistent kernel; 0 (default) = the device's SM count
 *   "debug_jitter"       non-zero seed: aig_mfcc_energy runs the jittered build of the persistent kernel, in which the TMA
 *                        producer, the MFCC consumers and the energy warps spin for pseudo-random times before every
 *                        barrier wait / arrive (race stress test standing in for compute-sanitizer), and aig_energy_heatmap
 *                        the jittered build of the warp-specialised kernel (float64 warps / heat-map warps); 0 (default): off
 *   "profile"            1: bracket every MFCC / energy kernel launch with CUDA events on the stream it
 *                        is launched on (read back with aig_profile_read); 0 (default): off
 * Unknown names or out-of-range values return AIG_ERR_ARGUMENT. */
int aig_set_option(aig_handle* h, const char* name, int64_t value);

/* Device self-tests of the energy stage's two arithmetic shortcuts (energy_kernel.cuh):
 *   which = 0  the Markstein division by the lifter constants against IEEE division for ALL 2^32 float32 inputs:
 *              out[0] = mismatching (input, lifter) pairs, out[1] = mismatches surviving the float32 store
 *   which = 1  the table-driven exp (1024-entry table, arguments in units of ln2 / 1024) against CUDA's exp() on
 *              2 * 2^26 points of [-700, 700] and [-12, 12]; the reference value is exp(u_hi) * (1 + u_lo) for the
 *              double-double natural argument, so two <= 1 ulp functions are compared (<= 2 ulp expected):
 *              out[0] = points differing, out[1] = largest difference in ulps, out[2] = points compared
 *   which = 2  the hoisted-reciprocal float32 division of the per-frame min-max normalisation against __fdiv_rn on
 *              2^32 (value, range) pairs: out[0] = pairs differing, out[1] = pairs compared, out[2] = pairs that took
 *              the reciprocal path (the rest fall back to IEEE division)
 * `out` is a host array of 4 uint64. */
int aig_selftest(aig_handle* h, int which, uint64_t* out);

/* Synchronise, then report and reset the event timings gathered while "profile" was on:
 * ms_out[3] / launches_out[3] = summed device time and launch count of
 * {0: MFCC kernels, 1: energy kernels, 2: every other kernel}. */
int aig_profile_read(aig_handle* h, double* ms_out, int64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* AIG_H_ */
