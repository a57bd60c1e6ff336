/* libvcpenc — C ABI of the B200-native H.264 encoder that replaces the `ffmpeg` child
 * process of the VCP consumer.
 *
 * Reference interfaces replaced (all paths relative to /root/reference):
 *   vcpenc_transcode      <- runFFmpegWithTimeout(ctx,input,output,ffmpegArgs,timeout)
 *                            cmd/consumer.go:370-394 (argv built at :376-380)
 *   vcpenc_verify         <- verifyOutputFile(path)           cmd/consumer.go:396-419
 *   vcpenc_parse_args     <- the option grammar of the preset strings
 *                            internal/config/config.go:44-52 after strings.Fields (:378)
 *   vcpenc_device_count   <- CUDA_VISIBLE_DEVICES convention  install.sh:279-297
 *   vcpenc_encode_frames, vcpenc_session_...: the in-memory core underneath (no reference
 *                            analogue; the reference hides this inside the ffmpeg process)
 *
 * Rules: plain pointers and sizes only; thread-safe and re-entrant; never aborts the
 * process; every failure is a non-zero return plus a message in `err`.  There is no CPU
 * fallback: with no usable CUDA device every encode entry point fails with
 * VCPENC_E_NODEVICE.
 */
#ifndef VCPENC_H
#define VCPENC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error classes (return values) */
#define VCPENC_OK 0
#define VCPENC_E_ARGS 1        /* malformed / unsupported option tokens            */
#define VCPENC_E_IO 2          /* cannot read input / write output                 */
#define VCPENC_E_FORMAT 3      /* input container or pixel format not understood   */
#define VCPENC_E_NODEVICE 4    /* no CUDA device (there is no CPU fallback)        */
#define VCPENC_E_CUDA 5        /* CUDA runtime error                               */
#define VCPENC_E_CANCELLED 6   /* *cancel became non-zero  ("任务被取消")           */
#define VCPENC_E_TIMEOUT 7     /* timeout_ms elapsed       ("编码超时")             */
#define VCPENC_E_NOTENCODE 8   /* preset is not a video encode (-c copy, -vn): the
                                  caller should hand the task to a stock ffmpeg    */
#define VCPENC_E_AUDIO 9       /* the input's audio cannot be carried into the output by the loaded FFmpeg libraries
                                  (no decoder / no aac encoder): the caller may hand the task to a stock ffmpeg */
#define VCPENC_E_VERIFY 10     /* verify: no valid video stream ("无有效视频流")     */
#define VCPENC_E_OVERFLOW 11   /* output buffer too small                          */
#define VCPENC_E_INTERNAL 12
#define VCPENC_E_UNSUPPORTED 13 /* a video encode this library does not implement (reserved; the HEVC
                                  presets run since k6_hevc.cu): like NOTENCODE the caller may hand the
                                  task to a stock ffmpeg                                           */

/* codecs */
#define VCPENC_CODEC_H264 0
#define VCPENC_CODEC_HEVC 1

/* rate-control modes */
#define VCPENC_RC_CQP 0        /* constant QP (also what -crf maps to)             */
#define VCPENC_RC_ABR 1        /* -b:v target, per-GOP budget                      */

/* input pixel formats accepted by the colour-convert kernel (K1) */
#define VCPENC_FMT_YUV420P 0
#define VCPENC_FMT_NV12 1
#define VCPENC_FMT_RGB24 2
#define VCPENC_FMT_YUV444P 3
#define VCPENC_FMT_YUV422P 4
#define VCPENC_FMT_BGR24 5

typedef struct vcpenc_params {
    int32_t width, height;     /* luma size of the OUTPUT picture (display size)    */
    int32_t fps_num, fps_den;
    int32_t codec;             /* VCPENC_CODEC_*                                    */
    int32_t gop;               /* closed-GOP length in frames (-g), IDR period      */
    int32_t rc_mode;           /* VCPENC_RC_*                                       */
    int32_t qp_i, qp_p;        /* CQP values / ABR starting point                   */
    int32_t bitrate;           /* bits per second for ABR (-b:v)                    */
    int32_t maxrate, bufsize;  /* -maxrate / -bufsize (bits/s, bits): both > 0 switch
                                  the per-GOP VBV model on (vcp_algo.h); -maxrate
                                  alone only caps the -b:v target                   */
    int32_t slices;            /* slices per picture (-slices)                      */
    int32_t deblock_idc;       /* disable_deblocking_filter_idc: 0 on, 1 off, 2 on
                                  but not across slice edges                        */
    int32_t entropy;           /* 0 CAVLC (-coder 0), 1 CABAC (-coder 1)            */
    int32_t in_fmt;            /* VCPENC_FMT_* of the frames handed in              */
    int32_t in_width, in_height; /* size of the frames handed in (0 = same as out)  */
    int32_t faststart;         /* -movflags +faststart: moov before mdat            */
    int32_t effort;            /* from -preset: 0 fast tiers (ultrafast..fast, p1..p3): the motion refine stops at half
                                  samples and inter macroblocks are not decimated; 1 medium (medium, p4, p5; the default);
                                  2 slow (slow.., p6, p7): as medium so far */
    int32_t debug;             /* 1: keep every reconstructed picture and per-MB
                                  decisions resident for the parity taps            */
    int32_t first_gop;         /* index of the first GOP handed in (sharded encodes):
                                  keeps idr_pic_id alternating across shards        */
    int32_t drop_audio;        /* -an: the input's audio is dropped (else it becomes an AAC track of the output:
                                  stream copy of AAC-LC, otherwise decode + libavcodec `aac` at -b:a)            */
    int32_t transform8x8;      /* 1: High profile, transform_8x8_mode_flag: inter macroblocks use the
                                  8x8 integer transform (-profile:v high, the libx264 default)  */
    int32_t hevc_subpel;       /* HEVC: 1 = half-sample luma motion (8-tap interpolation planes per picture), 2 = quarter
                                  samples as well: exact 7/8-tap motion compensation, the eight quarter-sample candidates
                                  ranked by averages of the half-sample planes; 3 = ranked by their exact prediction
                                  (2.7x the refine time for the same bits, measured).  The argument parser sets 2 for
                                  libx265 / hevc_nvenc (1 in the fast -preset tiers; -x265-params subme=0|1|2..4|5..
                                  selects 0|1|2|3)                                                               */
    int32_t hevc_sao;          /* HEVC: 1 = sample adaptive offset on luma (edge offsets, one decision per coding tree
                                  block, taken on the deblocked picture)                                         */
    int32_t hevc_intra_modes;  /* HEVC: 1 = intra CUs choose among planar / DC / horizontal / vertical prediction (else DC only).
                                  Implemented and pinned in the oracle; the device path does not have it yet:
                                  VCPENC_E_UNSUPPORTED                                                            */
    int32_t audio_bitrate;     /* -b:a (bits per second) for container inputs whose audio is re-encoded to AAC; 0: 128k */
    int32_t reserved[4];
} vcpenc_params;

/* per coded picture, filled by the encode calls */
typedef struct vcpenc_frame_info {
    uint64_t offset;           /* byte offset of the access unit in the output      */
    uint32_t size;             /* bytes (Annex-B, incl. start codes, SPS/PPS on IDR) */
    uint8_t  is_idr;
    uint8_t  qp;
    uint8_t  pad[2];
} vcpenc_frame_info;

/* per-kernel device timing, accumulated while profiling is enabled on a session */
#define VCPENC_K_CSC 0         /* K1 colour convert / pad / half-res pyramid        */
#define VCPENC_K_ME_PRE 1      /* K2a motion search pre-pass on originals           */
#define VCPENC_K_ME_REFINE 2   /* K2b full/half/quarter-pel refine on recon         */
#define VCPENC_K_P_RECON 3     /* K3 inter predict + T/Q/IQ/IT + recon              */
#define VCPENC_K_I_RECON 4     /* K3 intra predict + T/Q/IQ/IT + recon (wavefront)  */
#define VCPENC_K_MBINFO 5      /* mv prediction / skip / boundary strengths         */
#define VCPENC_K_DEBLOCK 6     /* K4 in-loop deblocking (wavefront)                 */
#define VCPENC_K_PAD 7         /* border extension of the reconstructed picture     */
#define VCPENC_K_CAVLC_COUNT 8 /* K5 pass 1: bits per macroblock                    */
#define VCPENC_K_CAVLC_SCAN 9  /* K5 pass 2: per-slice prefix sums + slice headers  */
#define VCPENC_K_CAVLC_WRITE 10/* K5 pass 3: bit-exact placement into the RBSP      */
#define VCPENC_K_RC 11         /* rate-control update                               */
#define VCPENC_K_HPEL 12       /* K2c half-sample planes of the reconstruction      */
#define VCPENC_K_CABAC_BINS 13 /* K5 CABAC: binarisation + context selection        */
#define VCPENC_K_CABAC_CODE 14 /* K5 CABAC: arithmetic coder (lane per slice) + NAL */
#define VCPENC_K_COUNT 15

typedef struct vcpenc_kernel_stat {
    double   ms;               /* summed CUDA-event time                            */
    uint64_t launches;
} vcpenc_kernel_stat;

/* ---- the two calls the consumer makes -------------------------------------------- */

/* Transcode `input` to `output` as the argv `ffmpeg -hide_banner -loglevel warning -y -i
 * input <argv...> output` would.  argv == strings.Fields(task.FFmpegArgs).  `cancel` is
 * polled at least once per GOP batch and replaces the SIGKILL that exec.CommandContext
 * delivers; timeout_ms <= 0 means none.  The callee creates/truncates `output`; on
 * failure the caller removes it (cmd/consumer.go:264). */
int vcpenc_transcode(const char* input, const char* output, int argc, const char* const* argv,
                     int timeout_ms, volatile int* cancel, char* err, size_t errlen);

/* ffprobe-equivalent acceptance check: file exists, size > 0, container parses and has a
 * video stream.  0 = ok. */
int vcpenc_verify(const char* path, char* err, size_t errlen);

/* ---- supporting entry points ---------------------------------------------------------- */

int vcpenc_device_count(void);           /* honours CUDA_VISIBLE_DEVICES; 0 if none    */
/* Device used by vcpenc_transcode calls made from the CALLING THREAD (default 0).  The reference
 * runs one consumer process per GPU under CUDA_VISIBLE_DEVICES (install.sh:279-297), where 0 is
 * right; a host that drives several GPUs from one process binds each worker thread with this. */
int vcpenc_set_thread_device(int device);
/* vcpenc_transcode reuses encoder sessions (all device buffers of one geometry / preset; a process-wide pool,
 * least recently used idle sessions are destroyed when memory is needed) and the two page-locked chunk buffers of
 * the calling thread: creating and freeing them costs more than encoding a short clip.  A host that retires a worker
 * thread releases its buffers (and every idle session) with this call; VCPENC_NO_CACHE=1 turns the reuse off. */
void vcpenc_thread_release(void);
const char* vcpenc_version(void);
void vcpenc_default_params(vcpenc_params* p);

/* Parse preset tokens into params (codec, rc, gop, ...).  Returns VCPENC_E_NOTENCODE for
 * `-c copy` / `-vn`, VCPENC_E_ARGS for unknown tokens. Host-only, needs no GPU. */
int vcpenc_parse_args(int argc, const char* const* argv, vcpenc_params* p, char* err,
                      size_t errlen);

/* Encode `nframes` frames held in HOST memory (tightly packed, p->in_fmt) on `device`;
 * GOPs are independent, so a caller may shard a clip by calling this once per GOP range.
 * Output: Annex-B byte stream in `out` (SPS/PPS in front of every IDR), one
 * vcpenc_frame_info per frame.  `recon` (optional, may be NULL) receives the encoder's
 * reconstructed pictures, yuv420p, display size — what a conformant decoder must output. */
int vcpenc_encode_frames(const vcpenc_params* p, int device, const uint8_t* frames,
                         int nframes, uint8_t* out, size_t out_cap, size_t* out_len,
                         vcpenc_frame_info* info, uint8_t* recon, volatile int* cancel,
                         char* err, size_t errlen);

/* Session API: keeps frames resident in HBM so the encode can be timed on the device. */
typedef struct vcpenc_session vcpenc_session;
int vcpenc_session_create(const vcpenc_params* p, int device, int max_frames,
                          vcpenc_session** out, char* err, size_t errlen);
/* copy host frames to the device and run K1 (convert, pad, pyramid) */
int vcpenc_session_upload(vcpenc_session* s, const uint8_t* frames, int nframes, char* err,
                          size_t errlen);
/* same, but returns once the copies are queued: the next vcpenc_session_encode starts each group of GOPs as
 * soon as ITS frames have landed, so the encode overlaps the rest of the transfer.  `frames` (pinned host
 * memory, vcpenc_host_alloc, for a truly asynchronous copy) must stay valid until that encode returns. */
int vcpenc_session_upload_async(vcpenc_session* s, const uint8_t* frames, int nframes, char* err,
                                size_t errlen);
/* vcpenc_session_upload_async for a buffer that is STILL BEING FILLED (a reader thread delivering a file): the copy of a
 * group of GOPs is queued as soon as *ready (pictures complete in `frames`, written by the producer with release
 * semantics) covers it, so the host-to-device transfer hides inside the read.  Returns when every copy is queued, or
 * VCPENC_E_CANCELLED when *finished becomes non-zero before `nframes` pictures exist (upload again with what there is). */
int vcpenc_session_upload_gated(vcpenc_session* s, const uint8_t* frames, int nframes, const volatile long* ready,
                                const volatile int* finished, char* err, size_t errlen);
/* vcpenc_session_upload_gated and vcpenc_session_encode in one call: the encode of a group of GOPs is queued right behind
 * its copy, so the first groups are being encoded while the producer still delivers the last ones.  Returns when the
 * bitstream is complete on the device (download next), or VCPENC_E_CANCELLED as above (nothing was encoded). */
int vcpenc_session_encode_gated(vcpenc_session* s, const uint8_t* frames, int nframes, const volatile long* ready,
                                const volatile int* finished, char* err, size_t errlen);
/* same, but the raw frames are already in device memory (`dframes` is a device pointer):
 * runs K1 only.  `ms` (optional) receives the CUDA-event time on the launching stream. */
int vcpenc_session_upload_device(vcpenc_session* s, const uint8_t* dframes, int nframes, float* ms,
                                 char* err, size_t errlen);
/* run K2..K5 over the resident frames; bitstream stays on the device.  If `ms` is given
 * it receives the CUDA-event time of the whole pass (events recorded on the launching
 * stream). */
int vcpenc_session_encode(vcpenc_session* s, float* ms, char* err, size_t errlen);
/* fetch the bitstream (and optionally recon) of the last encode */
int vcpenc_session_download(vcpenc_session* s, uint8_t* out, size_t out_cap, size_t* out_len,
                            vcpenc_frame_info* info, uint8_t* recon, char* err, size_t errlen);
/* per-kernel CUDA-event timing: enable, then encode, then read VCPENC_K_COUNT stats */
int vcpenc_session_profile(vcpenc_session* s, int enable);
int vcpenc_session_kernel_stats(vcpenc_session* s, vcpenc_kernel_stat* stats /*[K_COUNT]*/);
/* kernels launched by this session so far (always counted) */
uint64_t vcpenc_session_launch_count(vcpenc_session* s);
/* debug/parity taps: motion vectors (int16 x,y per MB per frame, quarter-pel) and MB types */
int vcpenc_session_debug_mbs(vcpenc_session* s, int16_t* mv_prepass, int16_t* mv_final,
                             uint8_t* mb_type, uint8_t* cbp);
/* GOP index (in the whole clip) of the first frame of the next upload; keeps idr_pic_id
 * alternating when a clip is fed in several chunks */
int vcpenc_session_set_first_gop(vcpenc_session* s, int first_gop);
void vcpenc_session_destroy(vcpenc_session* s);

/* Page-locked host memory for frame buffers handed to the upload calls (H2D at full PCIe
 * rate).  NULL on failure. */
void* vcpenc_host_alloc(size_t bytes);
void vcpenc_host_free(void* p);

/* What vcpenc_transcode sees of a CONTAINER input (.mp4 .mkv .avi .mov .webm ...; demux + decode by
 * libavformat/libavcodec loaded at run time, $VCPENC_FFMPEG_LIBDIR or the system's): geometry,
 * frame rate, VCPENC_FMT_* of the decoded pictures, and optionally up to max_frames decoded
 * pictures (tight layout) in `frames`.  Host-only. */
int vcpenc_probe_input(const char* path, int* width, int* height, int* fps_num, int* fps_den, int* fmt,
                       uint8_t* frames, size_t frames_cap, int max_frames, int* nframes, char* err,
                       size_t errlen);

/* The audio side of a container input as vcpenc_transcode handles it (every encode preset carries `-c:a aac -b:a Nk`,
 * internal/config/config.go:45-50): AAC-LC input is stream-copied, anything else is decoded and encoded with
 * libavcodec's `aac` encoder at audio_bitrate (0: 128k).  Returns the raw AAC access units back to back in `data`
 * with their sizes, the AudioSpecificConfig, and the encoder delay in samples.  Host-only. */
int vcpenc_probe_audio(const char* path, int audio_bitrate, int* sample_rate, int* channels, int* copied, int* priming,
                       uint8_t* asc, int asc_cap, int* asc_len, uint8_t* data, size_t data_cap, size_t* data_len,
                       uint32_t* sizes, int sizes_cap, int* nframes, char* err, size_t errlen);

/* Wrap an Annex-B stream produced above into an MP4 file (avc1/avcC, moov-first when
 * faststart).  Host-only. */
int vcpenc_mux_mp4(const vcpenc_params* p, const uint8_t* annexb, size_t len,
                   const vcpenc_frame_info* info, int nframes, const char* path, char* err,
                   size_t errlen);

/* Same with an AAC track: `aac` = raw access units back to back (aac_sizes[i] bytes each, 1024 samples per unit),
 * asc = AudioSpecificConfig, priming = encoder delay in samples (edit list).  Host-only. */
int vcpenc_mux_mp4_audio(const vcpenc_params* p, const uint8_t* annexb, size_t len, const vcpenc_frame_info* info,
                         int nframes, const uint8_t* aac, const uint32_t* aac_sizes, int aac_frames, int sample_rate,
                         int channels, int priming, const uint8_t* asc, int asc_len, const char* path, char* err,
                         size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* VCPENC_H */
