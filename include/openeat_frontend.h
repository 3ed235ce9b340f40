/* openeat_frontend.h -- C ABI of the B200-native OpenEAT acoustic front-end.
 *
 * The reference (TongtongSong/OpenEAT) has no native code and no FFI: its front-end is
 * Python calling torchaudio on the CPU, one utterance at a time.  This header is the
 * boundary a native replacement exposes instead; every entry point names the reference
 * code it replaces (paths relative to the reference tree, `kaldi.py` =
 * torchaudio/compliance/kaldi.py 2.11.0, the third-party function the reference calls).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an int status (OE_OK == 0); nothing throws or aborts across
 *     the ABI; oe_last_error() returns a thread-local message for the last failure.
 *   - the CALLER owns every buffer.  `d_` pointers are device memory on the handle's
 *     device; all others are host memory.  The library owns only the immutable tables
 *     inside the opaque handle (window, twiddles, sparse mel matrix, resampler taps).
 *   - all work is enqueued on the given cudaStream_t; no device-wide synchronisation and
 *     no allocation happens inside the batch calls (the caller passes a workspace whose
 *     size it queries first).  Host metadata arrays may be reused as soon as a call
 *     returns (they are copied to the workspace with a stream-ordered memcpy).
 *   - batches are ragged: one call handles B utterances of different lengths with ONE
 *     launch sequence (device-side work list), never one launch per utterance.
 */
#ifndef OPENEAT_FRONTEND_H_
#define OPENEAT_FRONTEND_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OE_ABI_VERSION 3    /* 2: oe_batch gained feature_dither / dither_seed / wav_dither (appended)
                             * 3: prepared batches (oe_batch_prepare / oe_fbank_run), ingest (oe_ingest_*) */

enum { OE_OK = 0, OE_ERR_INVALID = 1, OE_ERR_UNSUPPORTED = 2, OE_ERR_CUDA = 3, OE_ERR_WORKSPACE = 4 };
/* OE_WAV_I16: PCM int16.  OE_WAV_F32: fp32 on the int16 scale (dataset.py:75).
 * OE_FEATS_F32: the input already is log-mel features (rows x num_mel_bins fp32, data_type != 'wav',
 * dataset.py:190-191): wav_offsets / wav_lens then count ROWS, and only the post-fbank chain runs. */
enum { OE_WAV_I16 = 0, OE_WAV_F32 = 1, OE_FEATS_F32 = 2 };
enum { OE_NORM_NONE = 0, OE_NORM_PER_UTT = 1 };

typedef struct oe_frontend oe_frontend; /* opaque */
typedef void* oe_stream;                /* cudaStream_t */

/* Front-end parameters: the keyword arguments of the kaldi.fbank call at
 * openeat/dataset/dataset.py:93-100 (everything else is torchaudio's default). */
typedef struct {
    int32_t sample_rate;   /* sample_frequency                      (16000) */
    int32_t frame_length;  /* samples: sr*25 ms                     (400)   */
    int32_t frame_shift;   /* samples: sr*10 ms                     (160)   */
    int32_t fft_size;      /* round_to_power_of_two(frame_length)   (512)   */
    int32_t num_mel_bins;  /* feature_extraction_conf['mel_bins']   (80)    */
    float preemph;         /* preemphasis_coefficient               (0.97)  */
    float low_freq;        /* low_freq                              (20)    */
    float high_freq;       /* high_freq, <= 0: offset from Nyquist  (0)     */
    float log_floor;       /* torch.finfo(float32).eps, kaldi.py:31-37,631-633 */
} oe_config;

/* One ragged batch.  Replaces the per-utterance loop of _extract_feature
 * (dataset.py:53-111) plus the numeric part of audio_collate_func.__call__
 * (dataset.py:195-218) and GlobalCMVN.forward (openeat/modules/cmvn.py:35-46). */
typedef struct {
    int32_t batch;               /* B */
    int32_t wav_dtype;           /* OE_WAV_* of d_wav */
    const int64_t* wav_offsets;  /* [B] sample offset of each utterance in d_wav; multiple of 8 */
    const int32_t* wav_lens;     /* [B] samples.  < frame_length -> 0 frames: the reference drops
                                    such utterances (kaldi.py:142 raises, dataset.py:108-111) */
    const int64_t* out_rows;     /* [B] first output row of each utterance */
    const int32_t* out_nrows;    /* [B] rows to write (>= frames; the excess is padding, the zeros of
                                    pad_sequence at dataset.py:217-218), or NULL = frames */
    int64_t out_pitch;           /* floats between rows; 0 = num_mel_bins */
    int32_t norm_mode;           /* OE_NORM_PER_UTT = _normalization, feature_processor.py:5-8 */
    int32_t n_tmask;             /* time masks per utterance  (_spec_augmentation, :31-35) */
    int32_t n_fmask;             /* freq masks per utterance  (:37-41) */
    const int32_t* tmask;        /* [B][n_tmask][2] half-open frame ranges, or NULL */
    const int32_t* fmask;        /* [B][n_fmask][2] half-open bin ranges, or NULL */
    const int32_t* frame_map;    /* composed _spec_substitute index map (:44-64): output frame t of
                                    utterance b is source frame frame_map[frame_map_offsets[b]+t];
                                    NULL = identity.  Maps that only reach back (frame_map[t] <= t: what
                                    spec_sub composes to) are applied in place, any other gather goes
                                    through a raw scratch in the workspace */
    const int64_t* frame_map_offsets; /* [B] */
    const float* d_cmvn_mean;    /* DEVICE [num_mel_bins] or NULL (GlobalCMVN buffers) */
    const float* d_cmvn_istd;    /* DEVICE [num_mel_bins] or NULL (norm_var=False) */
    int32_t cmvn_on_padding;     /* 1: padding rows become (0-mean)*istd, what GlobalCMVN does to the
                                    zero-padded batch inside the encoder (encoder.py:221-222) */
    double* d_stats;             /* DEVICE [2*num_mel_bins+1] or NULL: += sum, sum of squares and
                                    frame count of the RAW log-mel frames (compute_cmvn_stats; the
                                    reference only ships the consumer, openeat/utils/cmvn.py:30-35) */
    int32_t* out_frames;         /* [B] out, or NULL: frames per utterance, 1+(N-400)//160 (kaldi.py:67) */
    const int32_t* resample_ids; /* [B] or NULL: fused speed perturb (_speed_perturb, audio_processor.py:19-35) --
                                    table id from oe_add_resampler (9:10 or 11:10, i.e. speed 0.9 / 1.1) or -1.
                                    The utterance is resampled while it is staged: no intermediate waveform;
                                    wav_lens stay INPUT lengths, frames follow ceil(new*N/orig).  int16 input only */
    float feature_dither;        /* a of dataset.py:199-201: x + (U[0,1) - 0.5) * a on the (normalised) features, before
                                    spec_sub / spec_aug; 0 = off.  The caller draws a = random.uniform(0, feature_dither)
                                    on the host (same `random` call as the reference); the per-cell uniforms come from
                                    Philox-4x32-10 keyed by dither_seed instead of numpy's global generator, so there is
                                    no value parity for this option */
    uint64_t dither_seed;        /* Philox key; vary it per batch */
    float wav_dither;            /* dither of kaldi.fbank (kaldi.py:179-181): every frame element gets + wav_dither * N(0,1)
                                    before DC removal; 0 = off.  Normals from Philox-4x32-10 + Box-Muller keyed by
                                    (dither_seed, utterance, frame, sample) instead of torch's global generator: no value
                                    parity.  Not combinable with resample_ids (resample first) */
} oe_batch;

/* One ragged resampling batch (speed perturb).  Replaces _speed_perturb
 * (openeat/dataset/audio_processor.py:19-35): `speed s` + `rate sr` == resample from
 * int(s*sr) to sr.  table_ids come from oe_add_resampler; -1 copies the utterance. */
typedef struct {
    int32_t batch;
    int32_t wav_dtype;           /* dtype of d_in; the output is always fp32 */
    const int64_t* in_offsets;   /* [B] */
    const int32_t* in_lens;      /* [B] */
    const int32_t* table_ids;    /* [B] */
    const int64_t* out_offsets;  /* [B] sample offsets in d_out (keep them multiples of 8) */
    int32_t* out_lens;           /* [B] out, or NULL: ceil(new*N/orig) */
    const int32_t* orig_rates;   /* [B] or NULL: rates of the utterances whose table_id is OE_RS_DIRECT -- the built-in */
    const int32_t* new_rates;    /* hann sinc evaluated on the fly, for ratios too long to tabulate (441:160, continuous speeds) */
} oe_resample_batch;
#define OE_RS_DIRECT (-2)

const char* oe_last_error(void);
int oe_abi_version(void);

int oe_config_default(oe_config* cfg);

/* window: [frame_length] or NULL; mel: [num_mel_bins][fft_size/2] row-major or NULL.
 * NULL tables are computed in double precision and rounded to fp32; pass the tables built
 * with torch's own fp32 expressions (kaldi.py:98-100, 436-511) for bit-identical constants.
 * `device` is the CUDA ordinal; -1 = current device. */
int oe_frontend_create(const oe_config* cfg, const float* window, const float* mel, int device,
                       oe_frontend** out);
int oe_frontend_destroy(oe_frontend* fe);
/* copies the tables in use back to the host (window[frame_length], mel[bins][fft/2]) */
int oe_frontend_get_tables(const oe_frontend* fe, float* window, float* mel);

/* 1 + (n - frame_length) / frame_shift, 0 when n < frame_length  (kaldi.py:63-67) */
int32_t oe_num_frames(const oe_frontend* fe, int64_t num_samples);

/* workspace bytes needed by oe_fbank_batch for this batch (depends on lengths and options) */
int oe_fbank_workspace_bytes(const oe_frontend* fe, const oe_batch* batch, size_t* bytes);

/* waveform -> [per-utt norm] -> [spec_sub] -> [spec_aug] -> [global CMVN] -> padded/ragged rows.
 * d_out may be NULL when only d_stats is wanted. */
int oe_fbank_batch(oe_frontend* fe, const oe_batch* batch, const void* d_wav, float* d_out,
                   void* d_workspace, size_t workspace_bytes, oe_stream stream);

/* Prepared batches.  oe_batch_prepare validates a batch description and packs its metadata ONCE (the library keeps
 * its own copy: every host array of `batch` may be freed afterwards; the DEVICE pointers d_cmvn_mean / d_cmvn_istd /
 * d_stats are kept as given); oe_fbank_run then costs one stream-ordered metadata copy plus the kernel launches -- the
 * hot loop of a trainer that plans batch i+1 on a worker thread while batch i runs (the reference does the same work
 * inside DataLoader workers, openeat/bin/train.py:110-116).  oe_fbank_batch == prepare + run + destroy.
 * A prepared batch may be run any number of times, with different d_wav / d_out / workspace / stream. */
typedef struct oe_prepared oe_prepared; /* opaque */
int oe_batch_prepare(oe_frontend* fe, const oe_batch* batch, oe_prepared** out);
int oe_prepared_destroy(oe_prepared* p);
size_t oe_prepared_workspace_bytes(const oe_prepared* p);
const int32_t* oe_prepared_frames(const oe_prepared* p);   /* [B] frames per utterance, valid until destroy */
int oe_fbank_run(oe_frontend* fe, const oe_prepared* p, const void* d_wav, float* d_out, void* d_workspace,
                 size_t workspace_bytes, oe_stream stream);

/* Small host -> device transfer that does not use the copy engine: `bytes` of host memory land at d_dst (16-byte
 * aligned, capacity rounded up to 16 bytes) through the handle's mapped pinned ring and a copy kernel, stream-ordered.
 * The host buffer may be reused as soon as the call returns.  For the int32 vectors of a batch (features_length,
 * targets, targets_length, dataset.py:221-231): a cudaMemcpy of those queues behind the next batch's PCM on the DMA
 * engine and stalls the collate pipeline by a whole bulk copy. */
int oe_upload_small(oe_frontend* fe, const void* host, size_t bytes, void* d_dst, oe_stream stream);

/* GlobalCMVN.forward (openeat/modules/cmvn.py:43-46): y = (x - mean) [* istd], rows x dim fp32. */
/* Number of kernels this handle has launched so far (oe_fbank_batch and oe_resample; every launch site of the
 * library counts itself).  bench.py reports the difference over its timed region as `gpu_launches`. */
int64_t oe_frontend_launch_count(const oe_frontend* fe);

/* Measurement hook: with timing on, oe_fbank_batch brackets its fbank kernel launch with two CUDA events on the
 * caller's stream (and launches that kernel and its successor without programmatic overlap, so the bracket holds
 * exactly that kernel).  oe_frontend_fbank_kernel_ms waits for the end event of the most recent call and returns the
 * kernel's duration.  bench.py uses it to report the dominant kernel's launch time INSIDE the timed step. */
int oe_frontend_set_kernel_timing(oe_frontend* fe, int32_t on);
int oe_frontend_fbank_kernel_ms(oe_frontend* fe, float* ms);
/* same hook: duration of the whole launch sequence of the most recent call (metadata copy .. last kernel) */
int oe_frontend_step_ms(oe_frontend* fe, float* ms);

int oe_cmvn_apply(const float* d_x, float* d_y, int64_t rows, int32_t dim, const float* d_mean,
                  const float* d_istd, oe_stream stream);

/* GlobalCMVN fused into the first layer of Conv2dSubsampling4 (forward only): replaces
 *   xs = global_cmvn(xs)                                  openeat/modules/encoder.py:221-222, cmvn.py:43-46
 *   Conv2d(1, odim, 3, 2) + ReLU on xs.unsqueeze(1)       openeat/modules/subsampling.py:76-78, 110-111
 * d_x (B, T, F) fp32 with row pitch `pitch` (0 = F) -- the padded tensor audio_collate_func returns; d_mean / d_istd [F]
 * or NULL (d_istd NULL = norm_var False); d_w [odim][3][3] (the Conv2d weight (odim, 1, 3, 3), contiguous), d_bias [odim]
 * or NULL; d_y (B, odim, (T-3)/2+1, (F-3)/2+1) fp32.  The normalised batch is never written: GlobalCMVN alone costs a
 * read and a write of the whole batch. */
int oe_cmvn_conv_subsample(const float* d_x, int64_t pitch, int32_t B, int32_t T, int32_t F, const float* d_mean,
                           const float* d_istd, const float* d_w, const float* d_bias, int32_t odim, float* d_y,
                           oe_stream stream);

/* Registers a polyphase table: kernel[new_rate][taps], taps = 2*width + orig_rate for any width >= 0
 * (torchaudio functional.py:1343-1398 layout; a long Kaiser design gives sox-quality resampling), or NULL =
 * hann-windowed sinc, lowpass width 6, rolloff 0.99 computed in double (OE_ERR_UNSUPPORTED when that table would
 * exceed 65 536 coefficients: use OE_RS_DIRECT).  Tables are kept for the life of the handle; storage grows on
 * demand (a setup-time call: it may synchronise the device).  Returns the table id in *table_id. */
int oe_add_resampler(oe_frontend* fe, int32_t orig_rate, int32_t new_rate, const float* kernel,
                     int32_t taps, int32_t* table_id);
/* 1 when the table can be resampled inside the fbank kernel's staging (oe_batch.resample_ids): the 9:10 / 11:10
 * tables whose bits equal the baked torchaudio tables.  oe_mel_is_baked: 1 when the handle runs the kernel that carries
 * torchaudio's 80-bin mel matrix as immediates (needed by wav_dither). */
int oe_resampler_fusable(const oe_frontend* fe, int32_t table_id);
int oe_mel_is_baked(const oe_frontend* fe);
int64_t oe_resample_out_len(int64_t n, int32_t orig_rate, int32_t new_rate);
int oe_resample_workspace_bytes(const oe_frontend* fe, const oe_resample_batch* batch, size_t* bytes);
int oe_resample(oe_frontend* fe, const oe_resample_batch* batch, const void* d_in, float* d_out,
                void* d_workspace, size_t workspace_bytes, oe_stream stream);

/* ---- native PCM ingest (host threads, no CUDA) ---------------------------------------------------------------
 * Replaces the per-utterance sox_io_backend.info + torchaudio.load of _extract_feature (dataset.py:55-75), which the
 * reference runs inside DataLoader worker processes (train.py:110-116): one call parses the RIFF/WAVE headers of a
 * whole batch (probe), the caller lays the utterances out (offsets: multiples of 8 samples) in ONE packed buffer --
 * normally the pinned staging buffer the H2D copy starts from -- and a second call reads every utterance's PCM
 * straight into its place with a pool of reader threads (pread, no intermediate copy for mono files; channel 0 of
 * multi-channel files, like torchaudio.load(...)[0]).
 *   paths[i]            file name; starts / ends: seconds of a segmented entry ("path,start,end", dataset.py:56-70:
 *                       frame_offset = int(start * sr), num_frames = int(end * sr) - frame_offset), starts[i] < 0
 *                       (or NULL arrays) = the whole file
 *   status[i]           OE_OK or an error code; oe_ingest_error(g, i) names the reason ("...: 24-bit FLAC
 *                       ...", "...: 24-bit samples ...", "...: No such file or directory").  Per-utterance failures do
 *                       not fail the call: the caller prints the message and drops the utterance, the reference's
 *                       convention (dataset.py:108-111).
 * 16-bit integer PCM RIFF/WAVE (incl. WAVE_FORMAT_EXTENSIBLE) is copied, 16-bit FLAC (the LibriSpeech corpus) is decoded
 * by the reader thread that owns the file (oe_flac_* below); anything libsox reads beyond that (24-bit, float, MP3) is
 * reported, not silently dropped.  OE_FLAC_VERIFY_MD5=1 additionally checks every stream's MD5 signature. */
typedef struct oe_ingest oe_ingest; /* opaque */
int oe_ingest_create(int32_t threads, oe_ingest** out);      /* threads <= 0: one per hardware thread */
int oe_ingest_destroy(oe_ingest* g);
int oe_ingest_probe(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                    int32_t* n_samples, int32_t* sample_rates, int32_t* status);
int oe_ingest_read(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                   int16_t* dst, const int64_t* offsets, const int32_t* n_samples, int32_t* status);
const char* oe_ingest_error(const oe_ingest* g, int32_t index);
/* The same asynchronously: oe_ingest_submit copies the request and returns at once; a driver thread of the handle
 * takes the jobs in order (probe -> 8-sample aligned layout -> read into dst, which must hold dst_capacity samples and
 * stay valid until the wait); oe_ingest_wait blocks until the job is done and hands out its result arrays (owned by
 * the job, valid until oe_ingest_job_release).  A destination that is too small fails the entries with
 * OE_ERR_WORKSPACE and reports the samples needed in *total.  Do not mix with oe_ingest_probe / oe_ingest_read on the
 * same handle while jobs are pending. */
typedef struct oe_ingest_job oe_ingest_job; /* opaque */
int oe_ingest_submit(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                     int16_t* dst, int64_t dst_capacity, oe_ingest_job** job);
int oe_ingest_wait(oe_ingest_job* job, const int64_t** offsets, const int32_t** n_samples, const int32_t** sample_rates,
                   const int32_t** status, int64_t* total);
const char* oe_ingest_job_error(const oe_ingest_job* job, int32_t index);
int oe_ingest_job_release(oe_ingest_job* job);

/* Host half of the CPU-tensor boundary (dataset.py:214-218: pad_sequence of the per-utterance feature matrices, then
 * GlobalCMVN on the padded tensor): scatters the ragged rows `src` (utterance b = frames[b] rows of F floats, one after
 * the other) into the padded tensor dst[B][tmax][F] and fills the rows behind every utterance with `pad_row` (F floats;
 * NULL = zeros), on the handle's reader threads with non-temporal stores.  Lets a pipeline bring back only the real rows
 * over PCIe.  Blocking; do not call while ingest jobs are pending on the same handle (use a handle of its own). */
int oe_host_pad_rows(oe_ingest* g, const float* src, const int32_t* frames, int32_t B, int32_t tmax, int32_t F,
                     const float* pad_row, float* dst);

/* ---- FLAC streams (host code, no CUDA) ---------------------------------------------------------------------------
 * Replaces torchaudio.load's libsox / libFLAC decode for .flac entries (openeat/dataset/dataset.py:62-72; the
 * LibriSpeech recipe's corpus).  Decoder written from RFC 9639; the frame header CRC-8 and the frame CRC-16 are always
 * checked, the STREAMINFO MD5 signature when verify_md5 != 0.  The native ingest above accepts 16-bit FLAC files next to
 * 16-bit PCM wav (one reader thread decodes one file); these two calls serve in-memory streams (shard-tar members) and
 * other sample sizes.  `data` is the whole .flac file.
 *   oe_flac_info    sample rate, channel count, bits per sample, samples per channel (counted by a decode pass when
 *                   STREAMINFO does not announce it)
 *   oe_flac_decode  samples [first, first + count) of `channel` as int32 (the stream's own integer scale), *decoded =
 *                   samples per channel in the stream */
int oe_flac_info(const void* data, int64_t size, int32_t* sample_rate, int32_t* channels, int32_t* bits, int64_t* total_samples);
int oe_flac_decode(const void* data, int64_t size, int32_t channel, int64_t first, int64_t count, int32_t* out,
                   int32_t verify_md5, int64_t* decoded);

/* ---- FLAC decoded on the GPU -----------------------------------------------------------------------------------------
 * For FLAC lists the compressed files cross PCIe (about half the bytes of their PCM for speech) and the GPU decodes
 * them: one thread per audio frame, straight into the packed int16 buffer oe_fbank_batch reads (oe_flac_gpu.cuh).
 *   oe_flac_pack          host: reads n files into `comp` (pinned; each file 16-byte aligned), walks every stream from
 *                         frame header to frame header and writes one oe_flac_frame per frame that overlaps the entry's
 *                         segment (starts / ends as for oe_ingest_probe), lays the utterances out at 8-sample aligned
 *                         pcm_offsets.  status[i] != OE_OK: oe_ingest_error(g, i) says why; OE_ERR_UNSUPPORTED marks
 *                         streams the GPU decoder does not take (more than one channel, other than 16 bits per sample, no announced
 *                         length) -- send those through oe_ingest_read / oe_flac_decode.  Returns OE_ERR_WORKSPACE with
 *                         *comp_bytes / *n_frames set to what is needed when a buffer is too small (comp needs 16 spare
 *                         bytes behind *comp_bytes).  Uses the handle's reader pool and per-batch state: one call at a
 *                         time per handle, and not while jobs of oe_flac_submit / oe_ingest_submit are in flight on it.
 *   oe_flac_decode_batch  device: decodes n_frames frames of d_comp into d_pcm; d_errors[utt] (caller-zeroed int32 per
 *                         entry) collects OE_FLAC_ERR_* bits: the frame must end where the host found the next header and
 *                         (verify_crc) match its CRC-16.  Stream-ordered, one launch.
 *   oe_flac_encode        host: 16-bit mono PCM -> a FLAC stream (fixed predictors, partitioned Rice coding, MD5
 *                         signature): writes the fixtures and the bench's corpus; returns OE_ERR_WORKSPACE with *bytes set
 *                         when `capacity` is too small. */
typedef struct oe_flac_frame {
    int64_t comp_off;      /* byte offset of the frame's sync code in the compressed buffer */
    int64_t out_off;       /* sample index in the PCM buffer of the first sample this frame contributes */
    int32_t frame_bytes;   /* header + subframe + CRC-16; <= 0: unknown (last frame), -frame_bytes = bytes available */
    int32_t hdr_bytes;     /* frame header incl. its CRC-8 */
    int32_t block;         /* samples in the frame */
    int32_t bps;           /* bits per sample */
    int32_t skip, take;    /* samples [skip, skip + take) of the frame belong to the entry's segment */
    int32_t utt;           /* entry index: errors are collected per entry */
    int32_t reserved;
} oe_flac_frame;
#define OE_FLAC_ERR_END 1      /* the frame did not end where the next one starts */
#define OE_FLAC_ERR_CRC 2      /* CRC-16 mismatch */
#define OE_FLAC_ERR_HOST 4     /* legal FLAC outside the GPU decoder's range (predictor order > 12): use the host decoder */
#define OE_FLAC_ERR_FORMAT 8   /* reserved codes / forbidden values */
#define OE_FLAC_ERR_OVERRUN 16 /* the bit stream ran past the end of the buffer */
int oe_flac_pack(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                 void* comp, int64_t comp_capacity, oe_flac_frame* frames, int64_t frames_capacity,
                 int64_t* comp_offsets, int64_t* pcm_offsets, int32_t* n_samples, int32_t* sample_rates, int32_t* status,
                 int64_t* comp_bytes, int64_t* n_frames, int64_t* total_samples);
/* oe_flac_pack on the handle's driver thread (like oe_ingest_submit / oe_ingest_wait: the caller -- a Python thread holding
 * the GIL -- only submits and, later, waits); buffers must stay valid until the wait returns.  oe_flac_wait returns what
 * oe_flac_pack returned (OE_ERR_WORKSPACE with *comp_bytes / *n_frames set: submit again with larger buffers); per-entry
 * messages through oe_ingest_job_error, release with oe_ingest_job_release. */
int oe_flac_submit(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends, void* comp,
                   int64_t comp_capacity, oe_flac_frame* frames, int64_t frames_capacity, oe_ingest_job** out);
int oe_flac_wait(oe_ingest_job* job, const int64_t** pcm_offsets, const int32_t** n_samples, const int32_t** sample_rates,
                 const int32_t** status, int64_t* comp_bytes, int64_t* n_frames, int64_t* total_samples);
int oe_flac_decode_batch(const void* d_comp, int64_t comp_bytes, const oe_flac_frame* d_frames, int64_t n_frames,
                         int16_t* d_pcm, int32_t* d_errors, int32_t verify_crc, oe_stream stream);
int oe_flac_encode(const int16_t* pcm, int64_t n, int32_t sample_rate, int32_t block, int32_t partition_order,
                   void* out, int64_t capacity, int64_t* bytes);

/* ---- host-side planning (no CUDA): the reference's random decisions, in its call order ----------
 * The reference draws every augmentation index from Python's global `random` module (Mersenne Twister).
 * These helpers continue that very generator natively: `mt_state` is `random.getstate()[1]` (624 state
 * words + position, 625 uint32), advanced in place exactly as CPython's random.random() /
 * _randbelow_with_getrandbits() would, so the indices equal the reference's for the same seed and the
 * caller can `random.setstate` the result back.
 *
 * oe_plan_speeds: per utterance, in input order (dataset.py:87-89):
 *     speed = item_speeds[i]; if random.random() < perturb_rate: speed = _speed_generator(speeds)
 *   with _speed_generator of audio_processor.py:5-18 (`speeds_cfg` = [start, end, step] or [fixed];
 *   n_speeds_cfg == 0 means None -> [0.9, 1.1, 0.1]).  active[i] == 0 skips the utterance (failed load:
 *   the reference raised before drawing).
 * oe_plan_augment: for all utterances in the given (length-sorted) order first every _spec_substitute
 *   draw (feature_processor.py:57-63, composed into frame_map), then every _spec_augmentation draw
 *   (:31-41) -> half-open, clipped ranges.  Pass do_sub / do_aug = 0 to skip either. */
int oe_plan_speeds(uint32_t* mt_state, int32_t n, double perturb_rate, const double* speeds_cfg,
                   int32_t n_speeds_cfg, const double* item_speeds, const uint8_t* active,
                   double* out_speeds);
int oe_plan_augment(uint32_t* mt_state, int32_t n, const int32_t* frames, int32_t num_freq,
                    int32_t do_sub, int32_t sub_max_t, int32_t sub_num_t_sub,
                    int32_t do_aug, int32_t num_t_mask, int32_t num_f_mask, int32_t max_t, int32_t max_f,
                    int32_t* frame_map, int32_t* tmask, int32_t* fmask);

#ifdef __cplusplus
}
#endif
#endif /* OPENEAT_FRONTEND_H_ */
