// Encoder core: session management, HBM allocation, GOP-batched stepping of the kernel chain.
//
// Execution model (DESIGN.md "Schedule"): closed GOPs are independent, so all GOPs resident on
// the device advance in lock-step — step t runs each kernel once over frame t of every GOP.
// The motion-search pre-pass runs once over all frames before the chain starts.  The only
// serial dimension is t (a P-frame needs the previous reconstruction); everything else is
// batch parallelism that fills the 148 SMs.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vcpenc.h"
#include "host_bits.h"
#include "vcp_dev.cuh"
#include "vcp_tma.cuh"

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_err(err, errlen, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return VCPENC_E_CUDA;                                                         \
        }                                                                                 \
    } while (0)

namespace {

// The encoder drives up to ~45 CUDA streams (GOP-group chains, entropy side streams, CABAC batches,
// copies).  The driver multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues
// (default 8); aliased streams serialise behind each other's long kernels (measured: CABAC 2 970 ->
// 8 330 fps with 32 queues).  The variable is read when the context is created, so it is set when
// the library is loaded, unless the operator chose a value.
__attribute__((constructor)) void vcp_more_hw_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

void set_err(char* err, size_t errlen, const char* fmt, ...) {
    if (!err || !errlen) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

size_t frame_bytes_of(const vcpenc_params& p) {
    const size_t w = p.width, h = p.height;
    return w * h + 2 * ((w + 1) / 2) * ((h + 1) / 2);
}

struct EventPair { cudaEvent_t a, b; int kind; };
constexpr int kCabacBatch = 8;   // pictures per GOP handed to the arithmetic coder at once

}  // namespace

struct vcpenc_session {
    vcpenc_params p{};
    VcpGeom g{};
    VcpBufs b{};
    VcpTmaps tm{};                              // tensor maps of the planes the motion search reads through TMA
    int device = 0;
    int max_frames = 0, nframes = 0, ngop_max = 0, ring = 2;
    int gop_base = 0;  // clip-level index of the first resident GOP
    cudaStream_t st = nullptr, st_copy = nullptr;
    cudaStream_t st_pre = nullptr;  // motion-search pre-pass, picture by picture, LOW priority: pure throughput work
                                    // that fills the latency gaps of the reconstruction chains
    std::vector<cudaEvent_t> ev_pre_t;   // [group][t]: pre-pass of picture t of the group's GOPs complete
    int pre_T = 0;                       // pictures per GOP the events cover
    bool streamed = false;               // the last upload was asynchronous: encode waits per group
    cudaStream_t st_up = nullptr;   // K1 of an upload: high priority, so that it is not queued behind the
                                    // thousands of CTAs of another session's encode on the same GPU
    static constexpr int kMaxGroups = 8;
    cudaEvent_t ev_piece[kMaxGroups] = {};      // streamed upload: the frames of GOP group k are in their planes
    cudaStream_t gst[kMaxGroups] = {};          // one stream per GOP group (the recon chain)
    cudaStream_t est[kMaxGroups] = {};          // entropy coding of the group, off the chain
    static constexpr int kCabacStreams = 8;
    int ncst = kCabacStreams;                    // streams per group actually used (the device has 32 hardware queues)
    cudaStream_t cst[kMaxGroups][kCabacStreams] = {};   // CABAC arithmetic coder batches: long-running, few
                                                        // warps each, independent -> they overlap each other
    cudaEvent_t ev_bins[kMaxGroups] = {};       // bins of the batch are complete
    cudaEvent_t gev[kMaxGroups] = {};
    cudaEvent_t ev_rec[kMaxGroups][2] = {};     // records of parity p are complete (after mbinfo)
    cudaEvent_t ev_ent[kMaxGroups][2] = {};     // entropy coding finished reading records of parity p
    VcpBufs bpar[2]{};                          // per-parity views of the double-buffered MB records
    cudaEvent_t ev_pre = nullptr;
    int ngroups = 4;
    int bins_per_mb = 128;                      // CABAC bin arena sizing (VCPENC_BINS_PER_MB)
    std::vector<void*> allocs;
    uint8_t* raw_dev = nullptr;                 // the raw frames of an upload, whole batch: H2D copies never wait for a kernel
    // K1 front stages (other pixel formats, scaling): scratch pictures of staging_frames each
    int in_w = 0, in_h = 0; size_t in_fb = 0;
    bool need_conv = false, need_scale = false;
    uint8_t *norm_a = nullptr, *norm_b = nullptr;
    cudaEvent_t staging_ready[2] = {nullptr, nullptr};   // H2D piece k landed in raw_dev
    int staging_frames = 0;                              // frames per pass of the K1 front stages' scratch
    // debug taps
    short2* dbg_mv = nullptr; uint8_t* dbg_type = nullptr; uint8_t* dbg_cbp = nullptr;
    // host
    std::vector<uint8_t> vps, sps, pps;   // vps: HEVC only
    std::vector<uint8_t> h_qp;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;   // pinned
    bool encoded = false;
    // profiling
    bool profile = false;
    std::vector<EventPair> events; size_t events_used = 0;
    vcpenc_kernel_stat stats[VCPENC_K_COUNT]{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
};

namespace {

template <typename T>
int dev_alloc(vcpenc_session* s, T** p, size_t count, char* err, size_t errlen) {
    void* q = nullptr;
    const size_t bytes = count * sizeof(T) + 256;  // slack: unaligned vector reads may touch a few bytes past the end
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {
        set_err(err, errlen, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return VCPENC_E_CUDA;
    }
    s->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return VCPENC_OK;
}

int check_params(const vcpenc_params& p, char* err, size_t errlen) {
    if (p.codec != VCPENC_CODEC_H264 && p.codec != VCPENC_CODEC_HEVC) { set_err(err, errlen, "unknown codec %d", p.codec); return VCPENC_E_ARGS; }
    if (p.hevc_subpel < 0 || p.hevc_subpel > 3) { set_err(err, errlen, "bad hevc_subpel %d", p.hevc_subpel); return VCPENC_E_ARGS; }
    if (p.hevc_intra_modes) { set_err(err, errlen, "HEVC intra modes beyond DC are implemented in the oracle only (device path: next round)"); return VCPENC_E_UNSUPPORTED; }
    if (p.hevc_sao < 0 || p.hevc_sao > 1) { set_err(err, errlen, "bad hevc_sao %d", p.hevc_sao); return VCPENC_E_ARGS; }
    if (p.entropy < 0 || p.entropy > 1) { set_err(err, errlen, "bad entropy coder %d", p.entropy); return VCPENC_E_ARGS; }
    if (p.effort < 0 || p.effort > 2) { set_err(err, errlen, "bad effort tier %d", p.effort); return VCPENC_E_ARGS; }
    if (p.width < 16 || p.height < 16 || (p.width & 1) || (p.height & 1)) { set_err(err, errlen, "unsupported picture size %dx%d", p.width, p.height); return VCPENC_E_ARGS; }
    if (p.gop < 1 || p.slices < 0 || p.slices > (p.height + 15) / 16) { set_err(err, errlen, "bad gop/slices"); return VCPENC_E_ARGS; }
    if (p.qp_i < 0 || p.qp_i > 51 || p.qp_p < 0 || p.qp_p > 51) { set_err(err, errlen, "qp out of range"); return VCPENC_E_ARGS; }
    if (p.in_fmt < VCPENC_FMT_YUV420P || p.in_fmt > VCPENC_FMT_BGR24) { set_err(err, errlen, "input pixel format %d not implemented", p.in_fmt); return VCPENC_E_FORMAT; }
    if (p.in_width < 0 || p.in_height < 0 || (p.in_width > 0) != (p.in_height > 0)) { set_err(err, errlen, "bad input size %dx%d", p.in_width, p.in_height); return VCPENC_E_ARGS; }
    if (p.deblock_idc < 0 || p.deblock_idc > 2) { set_err(err, errlen, "bad deblock_idc"); return VCPENC_E_ARGS; }
    if (p.rc_mode == VCPENC_RC_ABR && (p.bitrate <= 0 || p.fps_num <= 0 || p.fps_den <= 0)) { set_err(err, errlen, "bitrate mode needs -b:v and a frame rate"); return VCPENC_E_ARGS; }
    if (p.maxrate < 0 || p.bufsize < 0) { set_err(err, errlen, "bad -maxrate / -bufsize"); return VCPENC_E_ARGS; }
    return VCPENC_OK;
}

uint8_t initial_qp(const vcpenc_session* s, int n) {
    const bool idr = (n % s->p.gop) == 0;
    if (s->g.rc_abr) return (uint8_t)std::max(0, idr ? s->g.rc_qp0 - VCP_RC_QP_I_OFFSET : s->g.rc_qp0);
    return (uint8_t)(idr ? s->p.qp_i : s->p.qp_p);
}

struct Prof {
    vcpenc_session* s; int kind; size_t idx; bool on;
    cudaStream_t stream;
    Prof(vcpenc_session* s_, int kind_, int nlaunch = 1, cudaStream_t st_ = nullptr)
        : s(s_), kind(kind_), idx(0), on(s_->profile), stream(st_ ? st_ : s_->st) {
        s->launches += (uint64_t)nlaunch;
        if (!on) return;
        if (s->events_used == s->events.size()) {
            EventPair ep{};
            cudaEventCreate(&ep.a); cudaEventCreate(&ep.b);
            s->events.push_back(ep);
        }
        idx = s->events_used++;
        s->events[idx].kind = kind;
        cudaEventRecord(s->events[idx].a, stream);
    }
    ~Prof() { if (on) cudaEventRecord(s->events[idx].b, stream); }
};

void collect_profile(vcpenc_session* s) {
    for (size_t i = 0; i < s->events_used; i++) {
        float ms = 0;
        cudaEventSynchronize(s->events[i].b);
        cudaEventElapsedTime(&ms, s->events[i].a, s->events[i].b);
        s->stats[s->events[i].kind].ms += ms;
        s->stats[s->events[i].kind].launches++;
    }
    s->events_used = 0;
}

// CABAC arenas: bins_per_mb bins per macroblock on average over the whole session (128 ~ 25 Mb/s at
// 1080p30).  A picture may use more as long as the total fits; if it does not, the encode call
// grows the arenas and runs again (vcpenc_session_encode).
int alloc_cabac_arenas(vcpenc_session* s, char* err, size_t errlen) {
    VcpBufs& b = s->b;
    if (b.bins) { cudaFree(b.bins); s->allocs.erase(std::find(s->allocs.begin(), s->allocs.end(), (void*)b.bins)); b.bins = nullptr; }
    if (b.crbsp) { cudaFree(b.crbsp); s->allocs.erase(std::find(s->allocs.begin(), s->allocs.end(), (void*)b.crbsp)); b.crbsp = nullptr; }
    if (b.sbins) { cudaFree(b.sbins); s->allocs.erase(std::find(s->allocs.begin(), s->allocs.end(), (void*)b.sbins)); b.sbins = nullptr; }
    const size_t N = s->max_frames;
    b.bins_cap = std::max<size_t>(N * s->g.nmb * (size_t)s->bins_per_mb, (size_t)1 << 18);
    // slice streams: every bin once more, each stream rounded up to a 16-byte window plus one window of read-ahead
    b.sbins_cap = b.bins_cap + N * s->g.slices * 16;
    // slice RBSPs: a bin renormalises by at most 6 bits, so one byte per bin (+ header and flush) bounds a slice exactly
    b.crbsp_cap = b.bins_cap + N * s->g.slices * 160;
    int rc = dev_alloc(s, &b.bins, b.bins_cap, err, errlen);
    if (!rc) rc = dev_alloc(s, &b.sbins, b.sbins_cap, err, errlen);
    if (!rc) rc = dev_alloc(s, &b.crbsp, b.crbsp_cap, err, errlen);
    for (int q = 0; q < 2; q++) {
        s->bpar[q].bins = b.bins; s->bpar[q].bins_cap = b.bins_cap;
        s->bpar[q].sbins = b.sbins; s->bpar[q].sbins_cap = b.sbins_cap;
        s->bpar[q].crbsp = b.crbsp; s->bpar[q].crbsp_cap = b.crbsp_cap;
    }
    return rc;
}

// K1 over `cnt` raw frames already in device memory at `din` -> padded planes of frames n0..
// Other pixel formats / sizes go through the scratch pictures, staging_frames at a time.
void k1_chain(vcpenc_session* s, const uint8_t* din, int n0, int cnt, cudaStream_t st) {
    const size_t fb = frame_bytes_of(s->p);
    if (!s->need_conv) {
        Prof pr(s, VCPENC_K_CSC, 1, st);
        vcp_launch_k1_yuv420p(din, fb, n0, cnt, s->g, s->b, st);
        return;
    }
    const size_t afb = (size_t)vcp_in_frame_bytes(VCPENC_FMT_YUV420P, s->in_w, s->in_h);
    for (int i = 0; i < cnt; i += s->staging_frames) {
        const int k = std::min(s->staging_frames, cnt - i);
        Prof pr(s, VCPENC_K_CSC, s->need_scale ? 3 : 2, st);
        vcp_launch_k1_to_yuv420p(din + (size_t)i * s->in_fb, s->in_fb, s->p.in_fmt, s->in_w, s->in_h, s->norm_a, afb, k, st);
        const uint8_t* tight = s->norm_a;
        if (s->need_scale) {
            vcp_launch_k1_scale(s->norm_a, afb, s->in_w, s->in_h, s->norm_b, fb, s->p.width, s->p.height, k, st);
            tight = s->norm_b;
        }
        vcp_launch_k1_yuv420p(tight, fb, n0 + i, k, s->g, s->b, st);
    }
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point: libvcpenc.so does not link libcuda, so it still
// loads (for parse / verify / mux) on a host without a driver.
int vcp_make_tmap(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (EncodeFn)p;
    }();
    if (!fn) return -1;
    cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, base, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// free / total bytes of a device (transcode.cpp: session pool accounting); 0 on failure
size_t vcp_device_free_bytes(int device, size_t* total) {
    size_t f = 0, t = 0;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) { cudaGetLastError(); f = t = 0; }
    if (cur >= 0) cudaSetDevice(cur);
    if (total) *total = t;
    return f;
}

extern "C" {

int vcpenc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* vcpenc_version(void) { return "vcpenc 0.3 (sm_100a, H.264 Baseline/Main/High + HEVC Main, I/P)"; }

void vcpenc_default_params(vcpenc_params* p) {
    memset(p, 0, sizeof *p);
    p->fps_num = 30; p->fps_den = 1;
    p->codec = VCPENC_CODEC_H264;
    p->gop = 60;
    p->rc_mode = VCPENC_RC_CQP;
    p->qp_i = 24; p->qp_p = 26;
    p->slices = 1;
    p->deblock_idc = 0;
    p->entropy = 0;
    p->in_fmt = VCPENC_FMT_YUV420P;
    p->effort = 1;
}

void vcpenc_session_destroy(vcpenc_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->st) cudaStreamSynchronize(s->st);
    for (void* q : s->allocs) cudaFree(q);
    for (auto& e : s->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (int i = 0; i < 2; i++) {
        if (s->staging_ready[i]) cudaEventDestroy(s->staging_ready[i]);
    }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->h_out) cudaFreeHost(s->h_out);
    for (int i = 0; i < vcpenc_session::kMaxGroups; i++) {
        if (s->gst[i]) cudaStreamDestroy(s->gst[i]);
        if (s->est[i]) cudaStreamDestroy(s->est[i]);
        for (int q = 0; q < vcpenc_session::kCabacStreams; q++) if (s->cst[i][q]) cudaStreamDestroy(s->cst[i][q]);
        if (s->ev_bins[i]) cudaEventDestroy(s->ev_bins[i]);
        if (s->gev[i]) cudaEventDestroy(s->gev[i]);
        for (int q = 0; q < 2; q++) {
            if (s->ev_rec[i][q]) cudaEventDestroy(s->ev_rec[i][q]);
            if (s->ev_ent[i][q]) cudaEventDestroy(s->ev_ent[i][q]);
        }
    }
    if (s->ev_pre) cudaEventDestroy(s->ev_pre);
    if (s->st) cudaStreamDestroy(s->st);
    if (s->st_copy) cudaStreamDestroy(s->st_copy);
    if (s->st_up) cudaStreamDestroy(s->st_up);
    if (s->st_pre) cudaStreamDestroy(s->st_pre);
    for (auto e : s->ev_pre_t) cudaEventDestroy(e);
    for (auto e : s->ev_piece) if (e) cudaEventDestroy(e);
    delete s;
}

int vcpenc_session_create(const vcpenc_params* pp, int device, int max_frames, vcpenc_session** out,
                          char* err, size_t errlen) {
    if (!pp || !out || max_frames < 1) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    int rc = check_params(*pp, err, errlen);
    if (rc) return rc;
    const int ndev = vcpenc_device_count();
    if (ndev <= 0) { set_err(err, errlen, "no CUDA device available (libvcpenc has no CPU fallback)"); return VCPENC_E_NODEVICE; }
    if (device < 0 || device >= ndev) { set_err(err, errlen, "device %d out of range (%d visible)", device, ndev); return VCPENC_E_NODEVICE; }
    CK(cudaSetDevice(device));
    vcpenc_session* s = new vcpenc_session();
    s->p = *pp; s->device = device; s->max_frames = max_frames; s->gop_base = pp->first_gop;
    if (s->p.codec == VCPENC_CODEC_HEVC) { s->p.entropy = 1; s->p.transform8x8 = 0; }   // HEVC: CABAC only
    if (s->p.slices == 0) s->p.slices = vcp_auto_slices((pp->height + 15) / 16, s->p.entropy);   // encoder's choice
    pp = &s->p;
    VcpGeom& g = s->g;
    g.w = pp->width; g.h = pp->height;
    g.mbw = (g.w + 15) / 16; g.mbh = (g.h + 15) / 16; g.nmb = g.mbw * g.mbh;
    g.cw = 16 * g.mbw; g.ch = 16 * g.mbh;
    g.ys = (g.cw + 2 * VCP_PAD + 127) & ~127;
    g.cs = g.ys / 2; g.hs = g.ys / 2;
    g.ysize = (size_t)g.ys * (g.ch + 2 * VCP_PAD);
    g.csize = (size_t)g.cs * (g.ch / 2 + 2 * VCP_PADC);
    g.hsize = (size_t)g.hs * (g.ch / 2 + 2 * VCP_PAD1);
    g.yoff = VCP_PAD * g.ys + VCP_PAD;
    g.coff = VCP_PADC * g.cs + VCP_PADC;
    g.hoff = VCP_PAD1 * g.hs + VCP_PAD1;
    g.slices = pp->slices; g.deblock_idc = pp->deblock_idc; g.cabac = pp->entropy; g.t8x8 = pp->transform8x8 ? 1 : 0;
    g.hevc = pp->codec == VCPENC_CODEC_HEVC;
    g.hevc_subpel = g.hevc ? pp->hevc_subpel : 0;
    g.hevc_sao = g.hevc && pp->hevc_sao;
    g.effort = pp->effort;
    g.rc_abr = pp->rc_mode == VCPENC_RC_ABR;
    g.rc_bitrate = pp->bitrate; g.fps_num = pp->fps_num; g.fps_den = pp->fps_den;
    g.rc_qp0 = g.rc_abr ? vcp_rc_initial_qp(vcp_rc_eff_bitrate(pp->bitrate, pp->maxrate), pp->fps_num, pp->fps_den, pp->width, pp->height) : 0;
    {
        const bool vbv = vcp_rc_has_vbv(pp->maxrate, pp->bufsize, pp->fps_num, pp->fps_den) != 0;
        g.rc_fb = g.rc_abr || vbv;
        g.rc_maxrate = pp->maxrate; g.rc_qp_nom = pp->qp_p;
        g.rc_vbv_rate = vbv ? vcp_vbv_rate(pp->maxrate, pp->fps_num, pp->fps_den) : 0;
        g.rc_vbv_buf = vbv ? pp->bufsize : 0;
    }
    s->ngop_max = (max_frames + pp->gop - 1) / pp->gop;
    s->ring = pp->debug ? std::min(pp->gop, max_frames) : std::min(2, std::min(pp->gop, max_frames));
    if (s->ring < 1) s->ring = 1;

#define TRY(x) do { rc = (x); if (rc) { vcpenc_session_destroy(s); return rc; } } while (0)
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(err, errlen, "%s failed: %s", #call, cudaGetErrorString(e_)); vcpenc_session_destroy(s); return VCPENC_E_CUDA; } } while (0)
    CKS(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CKS(cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking));
    // Stream priorities (numerically lower = more urgent; plain streams sit at the least urgent level):
    // the reconstruction chains are latency-critical, the entropy side streams less so, and K1 of an
    // upload and the motion-search pre-pass are throughput work that should only fill gaps -- also
    // across sessions sharing the GPU (another consumer thread's upload must not slow this encode).
    int lo = 0, hi = 0;
    CKS(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int prio_chain = hi, prio_entropy = hi < lo ? hi + (lo - hi + 1) / 2 : lo;
    {
        // lowest priority for both: K1 of an upload and the pre-pass are throughput work; the chains of a
        // running encode (this session's or another one's on the same GPU) are latency-critical
        CKS(cudaStreamCreateWithPriority(&s->st_up, cudaStreamNonBlocking, lo));
        CKS(cudaStreamCreateWithPriority(&s->st_pre, cudaStreamNonBlocking, lo));
        s->pre_T = std::min(pp->gop, max_frames);
        s->ev_pre_t.resize((size_t)s->pre_T * 8);   // kMaxGroups
        for (auto& e : s->ev_pre_t) CKS(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : s->ev_piece) CKS(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    CKS(cudaEventCreate(&s->ev0)); CKS(cudaEventCreate(&s->ev1));
    CKS(cudaEventCreateWithFlags(&s->ev_pre, cudaEventDisableTiming));
    {
        const char* e = getenv("VCPENC_STREAMS");
        // H.264 hides two latency-bound wavefronts (deblocking, IDR intra) behind other groups' work: 4 groups.
        // The HEVC chain has no wavefront after the IDR picture, so fewer, larger launches win (measured 14.6k /
        // 13.8k / 12.0k fps with 2 / 4 / 8 groups).
        // H.264 with CABAC: the coder batches already fill the gaps; measured 1080p High 11.9k / 11.9k / 11.5k fps with
        // 2 / 3 / 4 groups (e2e 9.9k / 9.6k / 9.2k), 4K High 3.18k / 3.20k with 2 / 4.
        int ng = e ? atoi(e) : (pp->codec == VCPENC_CODEC_HEVC || pp->entropy ? 2 : 4);
        s->ngroups = std::max(1, std::min(ng, (int)vcpenc_session::kMaxGroups));
        const char* e2 = getenv("VCPENC_BINS_PER_MB");
        if (e2 && atoi(e2) > 0) s->bins_per_mb = std::min(atoi(e2), 65536);
    }
    {
        // Streams are multiplexed onto at most 32 hardware queues (CUDA_DEVICE_MAX_CONNECTIONS); streams that
        // share a queue serialise behind each other's long coder batches.  4 fixed + 2 per group + the coder
        // streams of all groups should fit: 8 per group with 2 groups, 5 with 4.
        const char* e3 = getenv("VCPENC_CABAC_STREAMS");
        const int fit = (32 - 4 - 2 * s->ngroups) / std::max(1, s->ngroups);
        s->ncst = std::max(1, std::min((int)vcpenc_session::kCabacStreams, e3 && atoi(e3) > 0 ? atoi(e3) : fit));
    }
    for (int i = 0; i < s->ngroups; i++) {
        CKS(cudaStreamCreateWithPriority(&s->gst[i], cudaStreamNonBlocking, prio_chain));
        CKS(cudaStreamCreateWithPriority(&s->est[i], cudaStreamNonBlocking, prio_entropy));
        if (pp->entropy) for (int q = 0; q < s->ncst; q++) CKS(cudaStreamCreateWithPriority(&s->cst[i][q], cudaStreamNonBlocking, prio_entropy));
        CKS(cudaEventCreateWithFlags(&s->ev_bins[i], cudaEventDisableTiming));
        CKS(cudaEventCreateWithFlags(&s->gev[i], cudaEventDisableTiming));
        for (int q = 0; q < 2; q++) {
            CKS(cudaEventCreateWithFlags(&s->ev_rec[i][q], cudaEventDisableTiming));
            CKS(cudaEventCreateWithFlags(&s->ev_ent[i][q], cudaEventDisableTiming));
        }
    }
    VcpBufs& b = s->b;
    const size_t N = max_frames, G = s->ngop_max, nmb = g.nmb;
    TRY(dev_alloc(s, &b.src_y, N * g.ysize, err, errlen));
    TRY(dev_alloc(s, &b.src_u, N * g.csize, err, errlen));
    TRY(dev_alloc(s, &b.src_v, N * g.csize, err, errlen));
    TRY(dev_alloc(s, &b.src_h, N * g.hsize, err, errlen));
    TRY(dev_alloc(s, &b.rec_y, G * s->ring * VCP_REC_PLANES * g.ysize, err, errlen));
    TRY(dev_alloc(s, &b.rec_u, G * s->ring * g.csize, err, errlen));
    TRY(dev_alloc(s, &b.rec_v, G * s->ring * g.csize, err, errlen));
    TRY(dev_alloc(s, &b.mvfp, N * nmb, err, errlen));
    TRY(dev_alloc(s, &b.mv, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.mvd, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.mbtype, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.cbp, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.modes, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.nnz, 2 * G * nmb * 24, err, errlen));
    TRY(dev_alloc(s, &b.levels, 2 * G * nmb * VCP_LV_STRIDE, err, errlen));
    TRY(dev_alloc(s, &b.mbbits, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.mbbitoff, 2 * G * nmb, err, errlen));
    TRY(dev_alloc(s, &b.skiprun, (size_t)1, err, errlen));
    TRY(dev_alloc(s, &b.qp, N, err, errlen));
    // raw slice payload: H.264 bounds a macroblock at 3200 bits; 512 B/MB leaves headroom
    const int rows_per_slice = (g.mbh + g.slices - 1) / g.slices + 1;
    b.rbsp_cap = ((size_t)rows_per_slice * g.mbw * 512 + 4096 + 15) & ~(size_t)15;
    TRY(dev_alloc(s, &b.rbsp, G * g.slices * b.rbsp_cap, err, errlen));
    TRY(dev_alloc(s, &b.slice_bits, G * g.slices, err, errlen));
    b.out_cap = N * frame_bytes_of(*pp) + (1 << 20);
    TRY(dev_alloc(s, &b.out, b.out_cap, err, errlen));
    TRY(dev_alloc(s, &b.out_cursor, (size_t)1, err, errlen));
    TRY(dev_alloc(s, &b.out_index, N * g.slices, err, errlen));
    TRY(dev_alloc(s, &b.out_index_hi, N * g.slices, err, errlen));
    TRY(dev_alloc(s, &b.frame_bits, N, err, errlen));
    TRY(dev_alloc(s, &b.error_flag, (size_t)1, err, errlen));
    TRY(dev_alloc(s, &b.rc_cum, G, err, errlen));
    TRY(dev_alloc(s, &b.rc_full, G, err, errlen));
    TRY(dev_alloc(s, &b.db_sync, G * (g.mbh + 1) + 1, err, errlen));
    TRY(dev_alloc(s, &b.icount, G * g.slices, err, errlen));
    CKS(cudaMemset(b.icount, 0, G * g.slices * sizeof(int)));
    if (g.cabac) {
        TRY(dev_alloc(s, &b.bins_cursor, (size_t)1, err, errlen));
        TRY(dev_alloc(s, &b.mbdesc, N * nmb, err, errlen));
        TRY(dev_alloc(s, &b.slice_bins, N * g.slices, err, errlen));
        TRY(dev_alloc(s, &b.crbsp_cursor, (size_t)1, err, errlen));
        TRY(dev_alloc(s, &b.sbins_cursor, (size_t)1, err, errlen));
        TRY(dev_alloc(s, &b.sslice_off, N * g.slices, err, errlen));
        TRY(dev_alloc(s, &b.cslice_off, N * g.slices, err, errlen));
        TRY(dev_alloc(s, &b.cslice_bytes, N * g.slices, err, errlen));
        TRY(alloc_cabac_arenas(s, err, errlen));
    }
    {
        uint32_t* ri = nullptr;
        TRY(dev_alloc(s, &ri, (size_t)g.mbh, err, errlen));
        std::vector<uint32_t> h((size_t)g.mbh);
        for (int r = 0; r < g.mbh; r++) {
            const int sl = vcp_slice_of_row(r, g.slices, g.mbh);
            h[r] = (uint32_t)vcp_slice_first_row(sl, g.slices, g.mbh) | ((uint32_t)sl << 16);
        }
        CKS(cudaMemcpy(ri, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        b.rowinfo = ri;
    }
    {
        // Tensor maps (vcp_tma.cuh): the half-res and full-res originals as (x, row, frame), the reconstruction
        // ring as (x, row, plane, slot).  Coordinates are plane coordinates (borders included), so a window
        // that leaves the picture simply reads the replicated border.
        const uint64_t hrows = (uint64_t)(g.ch / 2 + 2 * VCP_PAD1), yrows = (uint64_t)(g.ch + 2 * VCP_PAD);
        const uint64_t hd[3] = {(uint64_t)g.hs, hrows, N}, hst[2] = {(uint64_t)g.hs, g.hsize};
        const uint64_t yd[3] = {(uint64_t)g.ys, yrows, N}, yst[2] = {(uint64_t)g.ys, g.ysize};
        const uint64_t rd[4] = {(uint64_t)g.ys, yrows, VCP_REC_PLANES, G * s->ring}, rst[3] = {(uint64_t)g.ys, g.ysize, VCP_REC_PLANES * g.ysize};
        const uint32_t b_hwin[3] = {VCP_L1_WIN_W, VCP_L1_WIN_H, 1}, b_hcur[3] = {VCP_L1_CUR_W, 8, 1};
        const uint32_t b_yref[3] = {VCP_L0_REF_W, VCP_L0_REF_H, 1}, b_ycur[3] = {16, 16, 1};
        const uint32_t b_rec4[4] = {VCP_RF_WIN_W, VCP_RF_WIN_H, 4, 1}, b_rec1[4] = {VCP_RF_WIN_W, VCP_RF_WIN_H, 1, 1}, b_rec1q[4] = {VCP_RFQ_WIN_W, VCP_RFQ_WIN_H, 1, 1};
        int bad = vcp_make_tmap(&s->tm.h_win, b.src_h, 3, hd, hst, b_hwin) | vcp_make_tmap(&s->tm.h_cur, b.src_h, 3, hd, hst, b_hcur) |
                  vcp_make_tmap(&s->tm.y_ref, b.src_y, 3, yd, yst, b_yref) | vcp_make_tmap(&s->tm.y_cur, b.src_y, 3, yd, yst, b_ycur) |
                  vcp_make_tmap(&s->tm.rec4, b.rec_y, 4, rd, rst, b_rec4) | vcp_make_tmap(&s->tm.rec1, b.rec_y, 4, rd, rst, b_rec1) |
                  vcp_make_tmap(&s->tm.rec1q, b.rec_y, 4, rd, rst, b_rec1q);
        if (bad) { set_err(err, errlen, "cuTensorMapEncodeTiled failed (%d): the motion search needs TMA (sm_90+ driver)", bad); vcpenc_session_destroy(s); return VCPENC_E_CUDA; }
    }
    if (pp->debug) {
        TRY(dev_alloc(s, &s->dbg_mv, N * nmb, err, errlen));
        TRY(dev_alloc(s, &s->dbg_type, N * nmb, err, errlen));
        TRY(dev_alloc(s, &s->dbg_cbp, N * nmb, err, errlen));
    }
    s->staging_frames = std::max(1, std::min(16, max_frames));
    s->in_w = pp->in_width > 0 ? pp->in_width : pp->width;
    s->in_h = pp->in_height > 0 ? pp->in_height : pp->height;
    s->in_fb = (size_t)vcp_in_frame_bytes(pp->in_fmt, s->in_w, s->in_h);
    s->need_scale = s->in_w != pp->width || s->in_h != pp->height;
    s->need_conv = pp->in_fmt != VCPENC_FMT_YUV420P || s->need_scale;
    if (s->need_conv) TRY(dev_alloc(s, &s->norm_a, (size_t)s->staging_frames * vcp_in_frame_bytes(VCPENC_FMT_YUV420P, s->in_w, s->in_h), err, errlen));
    if (s->need_scale) TRY(dev_alloc(s, &s->norm_b, (size_t)s->staging_frames * frame_bytes_of(*pp), err, errlen));
    for (int i = 0; i < 2; i++) {
        if (i == 0) TRY(dev_alloc(s, &s->raw_dev, (size_t)max_frames * s->in_fb, err, errlen));
        CKS(cudaEventCreateWithFlags(&s->staging_ready[i], cudaEventDisableTiming));
    }
    // recon must never hold uninitialised borders when a vector reads them
    CKS(cudaMemsetAsync(b.rec_y, 128, G * s->ring * VCP_REC_PLANES * g.ysize, s->st));
    CKS(cudaMemsetAsync(b.rec_u, 128, G * s->ring * g.csize, s->st));
    CKS(cudaMemsetAsync(b.rec_v, 128, G * s->ring * g.csize, s->st));
    CKS(cudaMemsetAsync(b.mvfp, 0, N * nmb * sizeof(short2), s->st));
    CKS(cudaStreamSynchronize(s->st));
    for (int q = 0; q < 2; q++) {
        VcpBufs v = b;
        v.mv += q * G * nmb; v.mvd += q * G * nmb; v.mbtype += q * G * nmb; v.cbp += q * G * nmb;
        v.modes += q * G * nmb; v.nnz += q * G * nmb * 24; v.levels += q * G * nmb * VCP_LV_STRIDE;
        v.mbbits += q * G * nmb; v.mbbitoff += q * G * nmb;
        s->bpar[q] = v;
    }
    if (g.hevc) { s->vps = vcp::make_hevc_vps_nal(*pp); s->sps = vcp::make_hevc_sps_nal(*pp); s->pps = vcp::make_hevc_pps_nal(*pp); }
    else { s->sps = vcp::make_sps_nal(*pp); s->pps = vcp::make_pps_nal(*pp); }
    *out = s;
    return VCPENC_OK;
#undef TRY
#undef CKS
}

// GOP groups of an encode of N frames: group k covers GOPs [first, last)
static int group_count(const vcpenc_session* s, int N) {
    const int ngop_total = (N + s->p.gop - 1) / s->p.gop;
    return std::max(1, std::min(s->ngroups, ngop_total));
}
static void group_gops(const vcpenc_session* s, int N, int ng, int k, int* gA, int* gB) {
    const int ngop_total = (N + s->p.gop - 1) / s->p.gop;
    *gA = (int)((long long)ngop_total * k / ng); *gB = (int)((long long)ngop_total * (k + 1) / ng);
}

static int run_encode(vcpenc_session* s, char* err, size_t errlen, int konly = -1, int phase = 3);

// ready / finished (both or neither): the caller is still filling `frames`; piece k is queued once *ready pictures exist,
// and the call gives up with VCPENC_E_CANCELLED when *finished is set before that (the producer stopped short).
// fused: the encode of GOP group k is queued right behind its piece (vcpenc_session_encode_gated), so that the first
// groups are being encoded while the producer still delivers the last ones.
static int upload_host(vcpenc_session* s, const uint8_t* frames, int nframes, bool wait, char* err, size_t errlen,
                       const volatile long* ready = nullptr, const volatile int* finished = nullptr, bool fused = false) {
    if (!s || !frames || nframes < 1 || nframes > s->max_frames) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    CK(cudaSetDevice(s->device));
    const size_t fb = s->in_fb;
    s->nframes = nframes;
    s->encoded = false;
    s->h_qp.resize(nframes);
    for (int n = 0; n < nframes; n++) s->h_qp[n] = initial_qp(s, n);
    CK(cudaMemcpyAsync(s->b.qp, s->h_qp.data(), nframes, cudaMemcpyHostToDevice, s->st_up));
    // All H2D copies are queued up front into a device buffer that holds the whole batch, so the DMA
    // runs at PCIe rate whatever the GPU is busy with (another session's encode delays K1, never the
    // copies); K1 trails the copies piece by piece.  Pieces = the GOP groups of the encode, so that a
    // streamed upload (wait = false) lets group k start its chain as soon as ITS frames have landed.
    const int ng = s->profile ? 1 : group_count(s, nframes);
    if (fused) {
        if (s->profile || !ready) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
        s->streamed = true;
        const int rc = run_encode(s, err, errlen, -2, 1);
        if (rc) return rc;
    }
    for (int k = 0; k < ng; k++) {
        int gA, gB;
        group_gops(s, nframes, ng, k, &gA, &gB);
        const int n0 = gA * s->p.gop, n1 = std::min(nframes, gB * s->p.gop);
        if (n1 <= n0) { CK(cudaEventRecord(s->ev_piece[k], s->st_up)); continue; }
        if (ready) {
            // gated: GOP by GOP, so that the copy trails the producer by one GOP, not by one piece
            for (int a = n0; a < n1; a += s->p.gop) {
                const int e = std::min(n1, a + s->p.gop);
                while (*ready < e) {
                    if (finished && *finished && *ready < e) {
                        if (fused) { run_encode(s, err, errlen, -2, 2); cudaStreamSynchronize(s->st); }   // drain what was queued
                        s->nframes = 0;     // nothing usable was declared: the caller uploads again
                        set_err(err, errlen, "input ended after %ld of %d pictures", (long)*ready, nframes);
                        return VCPENC_E_CANCELLED;
                    }
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                }
                std::atomic_thread_fence(std::memory_order_acquire);
                CK(cudaMemcpyAsync(s->raw_dev + (size_t)a * fb, frames + (size_t)a * fb, (size_t)(e - a) * fb, cudaMemcpyHostToDevice, s->st_copy));
            }
        } else
        CK(cudaMemcpyAsync(s->raw_dev + (size_t)n0 * fb, frames + (size_t)n0 * fb, (size_t)(n1 - n0) * fb, cudaMemcpyHostToDevice, s->st_copy));
        CK(cudaEventRecord(s->staging_ready[k & 1], s->st_copy));
        CK(cudaStreamWaitEvent(s->st_up, s->staging_ready[k & 1], 0));
        k1_chain(s, s->raw_dev + (size_t)n0 * fb, n0, n1 - n0, s->st_up);
        CK(cudaEventRecord(s->ev_piece[k], s->st_up));
        if (fused) { const int rc = run_encode(s, err, errlen, k, 0); if (rc) return rc; }
    }
    if (fused) { const int rc = run_encode(s, err, errlen, -2, 2); if (rc) return rc; }
    CK(cudaGetLastError());
    s->streamed = !wait && !s->profile;
    if (!s->streamed) {
        CK(cudaStreamSynchronize(s->st_up));
        if (s->profile) collect_profile(s);
    }
    return VCPENC_OK;
}

int vcpenc_session_upload(vcpenc_session* s, const uint8_t* frames, int nframes, char* err, size_t errlen) {
    return upload_host(s, frames, nframes, true, err, errlen);
}

// Same, but returns as soon as the copies are queued: the following vcpenc_session_encode starts each GOP
// group when its frames have arrived, so the encode overlaps the rest of the transfer.  `frames` (pinned
// host memory for a truly asynchronous copy) must stay valid until that encode returns.
int vcpenc_session_upload_async(vcpenc_session* s, const uint8_t* frames, int nframes, char* err, size_t errlen) {
    return upload_host(s, frames, nframes, false, err, errlen);
}

// vcpenc_session_upload_async for a buffer that is still being filled: see include/vcpenc.h
int vcpenc_session_upload_gated(vcpenc_session* s, const uint8_t* frames, int nframes, const volatile long* ready,
                                const volatile int* finished, char* err, size_t errlen) {
    if (!ready || !finished) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    return upload_host(s, frames, nframes, false, err, errlen, ready, finished);
}

int vcpenc_session_upload_device(vcpenc_session* s, const uint8_t* dframes, int nframes, float* ms, char* err, size_t errlen) {
    if (!s || !dframes || nframes < 1 || nframes > s->max_frames) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    CK(cudaSetDevice(s->device));
    const size_t fb = s->in_fb;
    s->nframes = nframes;
    s->encoded = false;
    s->streamed = false;
    s->h_qp.resize(nframes);
    for (int n = 0; n < nframes; n++) s->h_qp[n] = initial_qp(s, n);
    CK(cudaMemcpyAsync(s->b.qp, s->h_qp.data(), nframes, cudaMemcpyHostToDevice, s->st));
    CK(cudaEventRecord(s->ev0, s->st));
    for (int n0 = 0; n0 < nframes; n0 += 4096)
        k1_chain(s, dframes + (size_t)n0 * fb, n0, std::min(4096, nframes - n0), s->st);
    CK(cudaEventRecord(s->ev1, s->st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->st));
    if (ms) CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    if (s->profile) collect_profile(s);
    return VCPENC_OK;
}

uint64_t vcpenc_session_launch_count(vcpenc_session* s) { return s ? s->launches : 0; }

// timing experiments only (the output is garbage): VCPENC_DEBUG_SKIP = bit mask of kernels NOT launched,
// 1 pre-pass, 2 refine, 4 p_recon, 8 deblock, 16 hpel, 32 CAVLC, 64 mbinfo/pad -- the marginal cost of a kernel inside the overlapped step
static const int g_dbg_skip = [] { const char* e = getenv("VCPENC_DEBUG_SKIP"); return e ? atoi(e) : 0; }();

// The whole encode of the resident pictures, queued on the session's streams (nothing here waits for the device).
// konly / phase: the gated encode (vcpenc_session_encode_gated) issues it in pieces -- phase bit 0 = the prologue
// (cursors reset, fork from the session stream), bit 1 = the epilogue (every stream joined back), konly = k: the
// pre-pass and the chain of GOP group k alone, -2: no group.  The default is everything at once.
static int run_encode(vcpenc_session* s, char* err, size_t errlen, int konly, int phase) {
    const VcpGeom& g = s->g;
    const VcpBufs& b = s->b;
    const int N = s->nframes, gop = s->p.gop;
    if (phase & 1) {
        CK(cudaMemsetAsync(b.out_cursor, 0, sizeof(unsigned long long), s->st));
        CK(cudaMemsetAsync(b.error_flag, 0, sizeof(int), s->st));
        CK(cudaMemsetAsync(b.frame_bits, 0, (size_t)N * sizeof(uint32_t), s->st));
        if (g.cabac) {
            CK(cudaMemsetAsync(b.bins_cursor, 0, sizeof(unsigned long long), s->st));
            CK(cudaMemsetAsync(b.crbsp_cursor, 0, sizeof(unsigned long long), s->st));
            CK(cudaMemsetAsync(b.sbins_cursor, 0, sizeof(unsigned long long), s->st));
            CK(cudaMemsetAsync(b.slice_bins, 0, (size_t)N * g.slices * sizeof(uint32_t), s->st));
            CK(cudaMemsetAsync(b.out_index, 0, (size_t)N * g.slices * sizeof(uint2), s->st));
        }
    }
    const int T = std::min(gop, N);
    // GOP groups advance on their own streams: the latency-bound wavefront kernels (intra
    // recon, deblocking) of one group overlap the throughput-bound kernels of the others.
    // Per-kernel profiling wants clean timings, so it runs everything on one stream.
    const int ng = s->profile ? 1 : group_count(s, N);
    const int kA = konly == -1 ? 0 : (konly < 0 ? 0 : konly), kB = konly == -1 ? ng : (konly < 0 ? 0 : konly + 1);   // groups issued by this call
    const bool streamed = s->streamed && !s->profile;
    if (s->streamed && s->profile) CK(cudaStreamSynchronize(s->st_up));   // profiling was switched on after an asynchronous upload
    if (s->profile) {
        Prof pr(s, VCPENC_K_ME_PRE, 2);
        vcp_launch_me_prepass(g, b, s->tm, N, gop, -1, s->st);
    } else {
        // picture by picture and group by group on its own stream: chain step t of a group only waits for
        // the vectors of picture t of its GOPs.  Resident input: pictures outermost (every group gets its
        // first vectors early).  Streamed input: groups outermost, each behind the arrival of its frames.
        if (phase & 1) {
            CK(cudaEventRecord(s->ev_pre, s->st));
            CK(cudaStreamWaitEvent(s->st_pre, s->ev_pre, 0));
        }
        auto pre = [&](int k, int t) -> int {
            int gA, gB;
            group_gops(s, N, ng, k, &gA, &gB);
            s->launches += 2;   // L1 + L0 kernels
            if (!(g_dbg_skip & 1)) vcp_launch_me_prepass(g, b, s->tm, N, gop, t, s->st_pre, gA, gB);
            CK(cudaEventRecord(s->ev_pre_t[(size_t)k * s->pre_T + t], s->st_pre));
            return VCPENC_OK;
        };
        if (streamed) {
            for (int k = kA; k < kB; k++) {
                CK(cudaStreamWaitEvent(s->st_pre, s->ev_piece[k], 0));
                for (int t = 1; t < T; t++) { const int rc = pre(k, t); if (rc) return rc; }
            }
        } else {
            for (int t = 1; t < T; t++)
                for (int k = kA; k < kB; k++) { const int rc = pre(k, t); if (rc) return rc; }
        }
    }
    if (phase & 1) CK(cudaEventRecord(s->ev_pre, s->st));   // (the session stream has not moved since the prologue)
    for (int k = kA; k < kB; k++) {
        CK(cudaStreamWaitEvent(s->gst[k], s->ev_pre, 0));
        if (streamed) CK(cudaStreamWaitEvent(s->gst[k], s->ev_piece[k], 0));
    }
    for (int t = 0; t < gop && t < N; t++) {
        const int par = t & 1;
        const VcpBufs& bt = s->bpar[par];     // macroblock records of this step
        for (int k = kA; k < kB; k++) {
            cudaStream_t st = s->profile ? s->st : s->gst[k];
            cudaStream_t se = s->profile ? s->st : s->est[k];   // entropy stream
            VcpStep sp;
            sp.t = t; sp.gop = gop; sp.ring = s->ring; sp.nframes = N; sp.gop0 = s->gop_base;
            // GOPs of this group that own a frame at position t
            int gA, gB;
            group_gops(s, N, ng, k, &gA, &gB);
            const int active = (N - t + gop - 1) / gop;      // GOPs (from 0) that have frame t
            sp.g0 = gA;
            sp.ngop = std::min(gB, active) - gA;
            if (sp.ngop <= 0) continue;
            // the records of this parity were last read by the entropy pass of step t-2
            if (!s->profile && t >= 2) CK(cudaStreamWaitEvent(st, s->ev_ent[k][par], 0));
            if (g.hevc) {
                // same chain, HEVC kernels (k6_hevc.cu): no half-sample planes
                if (t == 0) { Prof pr(s, VCPENC_K_I_RECON, 1, st); vcp_launch_hevc_i_recon(g, bt, sp, st); }
                else {
                    if (!s->profile) CK(cudaStreamWaitEvent(st, s->ev_pre_t[(size_t)k * s->pre_T + t], 0));
                    { Prof pr(s, VCPENC_K_ME_REFINE, 1, st); vcp_launch_me_refine(g, bt, s->tm, sp, st); }
                    { Prof pr(s, VCPENC_K_P_RECON, 2, st); vcp_launch_hevc_p_recon(g, bt, sp, st); vcp_launch_hevc_i_fix(g, bt, sp, st); }
                    { Prof pr(s, VCPENC_K_MBINFO, 1, st); vcp_launch_hevc_cuinfo(g, bt, sp, st); }
                }
                if (g.hevc_sao) {
                    // the SAO parameters are slice data: deblocking and the SAO decision run before the entropy fork
                    if (g.deblock_idc != 1) { Prof pr(s, VCPENC_K_DEBLOCK, 2, st); vcp_launch_hevc_deblock(g, bt, sp, st); }
                    { Prof pr(s, VCPENC_K_DEBLOCK, 1, st); vcp_launch_hevc_sao(g, bt, sp, st); }
                }
            } else if (t == 0) {
                Prof pr(s, VCPENC_K_I_RECON, 1, st);
                vcp_launch_i_recon(g, bt, sp, st);
            } else {
                if (!s->profile) CK(cudaStreamWaitEvent(st, s->ev_pre_t[(size_t)k * s->pre_T + t], 0));
                { Prof pr(s, VCPENC_K_ME_REFINE, 1, st); if (!(g_dbg_skip & 2)) vcp_launch_me_refine(g, bt, s->tm, sp, st); }
                { Prof pr(s, VCPENC_K_P_RECON, 2, st); if (!(g_dbg_skip & 4)) { vcp_launch_p_recon(g, bt, sp, st); vcp_launch_i_fix(g, bt, sp, st); } }
                { Prof pr(s, VCPENC_K_MBINFO, 1, st); if (!(g_dbg_skip & 64)) vcp_launch_mbinfo(g, bt, sp, st); }
            }
            if (s->p.debug) {
                for (int gi = gA; gi < gA + sp.ngop; gi++) {
                    const size_t n = (size_t)gi * gop + t;
                    CK(cudaMemcpyAsync(s->dbg_mv + n * g.nmb, bt.mv + (size_t)gi * g.nmb, g.nmb * sizeof(short2), cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpyAsync(s->dbg_type + n * g.nmb, bt.mbtype + (size_t)gi * g.nmb, g.nmb, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpyAsync(s->dbg_cbp + n * g.nmb, bt.cbp + (size_t)gi * g.nmb, g.nmb, cudaMemcpyDeviceToDevice, st));
                }
            }
            // entropy coding only reads the records: it leaves the recon chain here
            if (!s->profile) {
                CK(cudaEventRecord(s->ev_rec[k][par], st));
                CK(cudaStreamWaitEvent(se, s->ev_rec[k][par], 0));
            }
            if (g.cabac) {
                { Prof pr(s, VCPENC_K_CABAC_BINS, g.rc_fb ? 3 : 2, se); if (g.hevc) vcp_launch_hevc_bins(g, bt, sp, se); else vcp_launch_cabac_bins(g, bt, sp, se); }
                if (g.rc_fb) { Prof pr(s, VCPENC_K_RC, 1, se); vcp_launch_rc_update(g, bt, sp, se); }
                // The arithmetic coder takes the bins in batches on its own streams: one lane per slice, long-running but
                // only a few warps wide.  The IDR pictures are a batch of their own on a stream of their own: their
                // slices carry ten times the bins of a P slice, and a later batch queued behind them on the same stream
                // would not start before they end (measured: 57 of 212 ms per step).  P pictures go kCabacBatch at a
                // time, rotating over the other streams.
                const int last_t = std::min(gop, N) - 1;
                const int next_active = std::min(gB, (N - (t + 1) + gop - 1) / gop) - gA;   // GOPs of the group with a picture t+1
                if (t == 0 || t % kCabacBatch == 0 || t == last_t || next_active <= 0) {
                    const int t0 = t == 0 ? 0 : (t - 1) / kCabacBatch * kCabacBatch + 1;
                    const int bidx = t == 0 ? 0 : 1 + (t - 1) / kCabacBatch;
                    cudaStream_t sc = s->profile ? s->st : s->cst[k][bidx == 0 || s->ncst < 2 ? 0 : 1 + (bidx - 1) % (s->ncst - 1)];
                    VcpStep sb = sp;
                    sb.ngop = std::min(gB, (N - t0 + gop - 1) / gop) - gA;   // GOPs of the group that own picture t0
                    if (!s->profile) { CK(cudaEventRecord(s->ev_bins[k], se)); CK(cudaStreamWaitEvent(sc, s->ev_bins[k], 0)); }
                    Prof pr(s, VCPENC_K_CABAC_CODE, 2, sc);
                    static const int dbg_skip = [] { const char* e = getenv("VCPENC_DEBUG_SKIP_CODER"); return e ? atoi(e) : 0; }();   // timing experiments only: the output is empty
                    if (!dbg_skip) vcp_launch_cabac_encode(g, bt, sb, t0, t + 1, sc);
                }
            } else {
                if (!(g_dbg_skip & 32)) {
                { Prof pr(s, VCPENC_K_CAVLC_COUNT, 1, se); vcp_launch_cavlc_count(g, bt, sp, se); }
                { Prof pr(s, VCPENC_K_CAVLC_SCAN, 1, se); vcp_launch_cavlc_scan(g, bt, sp, se); }
                { Prof pr(s, VCPENC_K_CAVLC_WRITE, 2, se); vcp_launch_cavlc_write(g, bt, sp, se); vcp_launch_nal_pack(g, bt, sp, se); }
                }
                if (g.rc_fb) { Prof pr(s, VCPENC_K_RC, 1, se); vcp_launch_rc_update(g, bt, sp, se); }
            }
            if (!s->profile) CK(cudaEventRecord(s->ev_ent[k][par], se));
            if (g.hevc_sao) { Prof pr(s, VCPENC_K_DEBLOCK, 1, st); vcp_launch_hevc_sao_copy(g, bt, sp, st); }   // after the fork: only the picture changes
            else if (g.deblock_idc != 1) {
                Prof pr(s, VCPENC_K_DEBLOCK, g.hevc ? 2 : 1, st);
                if (g.hevc) vcp_launch_hevc_deblock(g, bt, sp, st); else if (!(g_dbg_skip & 8)) vcp_launch_deblock(g, bt, sp, st);
            }
            { Prof pr(s, VCPENC_K_PAD, 1, st); if (!(g_dbg_skip & 64)) vcp_launch_pad(g, bt, sp, st); }
            // half-sample planes of this reconstruction for the next picture's search and prediction
            if (t + 1 < gop && t + 1 < N && (!g.hevc || g.hevc_subpel)) { Prof pr(s, VCPENC_K_HPEL, 1, st); if (!(g_dbg_skip & 16)) vcp_launch_hpel(g, bt, sp, st); }
        }
    }
    if (!s->profile && (phase & 2))
        for (int k = 0; k < ng; k++) {
            CK(cudaEventRecord(s->gev[k], s->gst[k]));
            CK(cudaStreamWaitEvent(s->st, s->gev[k], 0));
            CK(cudaEventRecord(s->gev[k], s->est[k]));
            CK(cudaStreamWaitEvent(s->st, s->gev[k], 0));
            if (g.cabac) for (int q = 0; q < s->ncst; q++) {
                CK(cudaEventRecord(s->gev[k], s->cst[k][q]));
                CK(cudaStreamWaitEvent(s->st, s->gev[k], 0));
            }
        }
    CK(cudaGetLastError());
    return VCPENC_OK;
}

// issued: the first pass is already queued on the session's streams (vcpenc_session_encode_gated)
static int encode_and_wait(vcpenc_session* s, float* ms, bool issued, char* err, size_t errlen) {
    int flag = 0;
    for (int attempt = 0;; attempt++) {
        int rc = 0;
        const auto h0 = std::chrono::steady_clock::now();
        if (!(issued && attempt == 0)) {
            CK(cudaEventRecord(s->ev0, s->st));
            rc = run_encode(s, err, errlen);
            if (rc) return rc;
            CK(cudaEventRecord(s->ev1, s->st));
        } else ms = nullptr;
        const auto h1 = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(s->st));
        if (getenv("VCPENC_TRACE_ENCODE")) {
            const auto h2 = std::chrono::steady_clock::now();
            fprintf(stderr, "[vcpenc] encode: host issue %.1f ms, then waited %.1f ms\n",
                    std::chrono::duration<double, std::milli>(h1 - h0).count(), std::chrono::duration<double, std::milli>(h2 - h1).count());
        }
        if (ms) CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
        if (s->profile) collect_profile(s);
        CK(cudaMemcpy(&flag, s->b.error_flag, sizeof flag, cudaMemcpyDeviceToHost));
        // CABAC arenas too small for this content (very low QP): grow x4 and run the pass again
        if ((flag == 3 || flag == 4) && attempt < 6 && s->bins_per_mb < 65536) {
            s->bins_per_mb = std::min(65536, s->bins_per_mb * 4);
            rc = alloc_cabac_arenas(s, err, errlen);
            if (rc) return rc;
            continue;
        }
        break;
    }
    if (flag) { set_err(err, errlen, "bitstream buffer overflow on device (flag %d)", flag); return VCPENC_E_OVERFLOW; }
    s->encoded = true;
    return VCPENC_OK;
}

int vcpenc_session_encode(vcpenc_session* s, float* ms, char* err, size_t errlen) {
    if (!s || s->nframes < 1) { set_err(err, errlen, "no frames uploaded"); return VCPENC_E_ARGS; }
    CK(cudaSetDevice(s->device));
    return encode_and_wait(s, ms, false, err, errlen);
}

// upload_gated + encode in one call: see include/vcpenc.h
int vcpenc_session_encode_gated(vcpenc_session* s, const uint8_t* frames, int nframes, const volatile long* ready,
                                const volatile int* finished, char* err, size_t errlen) {
    if (!ready || !finished) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    const int rc = upload_host(s, frames, nframes, false, err, errlen, ready, finished, true);
    if (rc) return rc;
    return encode_and_wait(s, nullptr, true, err, errlen);
}

int vcpenc_session_download(vcpenc_session* s, uint8_t* out, size_t out_cap, size_t* out_len,
                            vcpenc_frame_info* info, uint8_t* recon, char* err, size_t errlen) {
    if (!s || !s->encoded || !out || !out_len) { set_err(err, errlen, "nothing encoded"); return VCPENC_E_ARGS; }
    CK(cudaSetDevice(s->device));
    const VcpGeom& g = s->g;
    const int N = s->nframes, S = g.slices;
    unsigned long long used = 0;
    CK(cudaMemcpy(&used, s->b.out_cursor, sizeof used, cudaMemcpyDeviceToHost));
    if (used > s->h_out_cap) {
        if (s->h_out) cudaFreeHost(s->h_out);
        s->h_out = nullptr; s->h_out_cap = 0;
        const size_t cap = (size_t)used + (used >> 2) + 4096;
        CK(cudaMallocHost(reinterpret_cast<void**>(&s->h_out), cap));
        s->h_out_cap = cap;
    }
    std::vector<uint2> idx((size_t)N * S);
    std::vector<uint32_t> idx_hi((size_t)N * S);
    CK(cudaMemcpyAsync(s->h_out, s->b.out, (size_t)used, cudaMemcpyDeviceToHost, s->st));
    CK(cudaMemcpyAsync(idx.data(), s->b.out_index, idx.size() * sizeof(uint2), cudaMemcpyDeviceToHost, s->st));
    CK(cudaMemcpyAsync(idx_hi.data(), s->b.out_index_hi, idx_hi.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->st));
    if (s->g.rc_fb) CK(cudaMemcpyAsync(s->h_qp.data(), s->b.qp, (size_t)N, cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    static const uint8_t sc[4] = {0, 0, 0, 1};
    size_t o = 0;
    for (int n = 0; n < N; n++) {
        const bool idr = (n % s->p.gop) == 0;
        const size_t au0 = o;
        size_t need = 0;
        if (idr) need += 12 + s->vps.size() + s->sps.size() + s->pps.size();
        for (int k = 0; k < S; k++) need += idx[(size_t)n * S + k].y;
        if (o + need > out_cap) { set_err(err, errlen, "output buffer too small (%zu needed)", o + need); return VCPENC_E_OVERFLOW; }
        if (idr) {
            if (!s->vps.empty()) { memcpy(out + o, sc, 4); o += 4; memcpy(out + o, s->vps.data(), s->vps.size()); o += s->vps.size(); }
            memcpy(out + o, sc, 4); o += 4; memcpy(out + o, s->sps.data(), s->sps.size()); o += s->sps.size();
            memcpy(out + o, sc, 4); o += 4; memcpy(out + o, s->pps.data(), s->pps.size()); o += s->pps.size();
        }
        for (int k = 0; k < S; k++) {
            const uint2 e = idx[(size_t)n * S + k];
            const size_t off = ((size_t)idx_hi[(size_t)n * S + k] << 32) | e.x;
            memcpy(out + o, s->h_out + off, e.y);
            o += e.y;
        }
        if (info) {
            info[n].offset = au0; info[n].size = (uint32_t)(o - au0);
            info[n].is_idr = idr; info[n].qp = s->h_qp[n]; info[n].pad[0] = info[n].pad[1] = 0;
        }
    }
    *out_len = o;
    if (recon) {
        if (!s->p.debug) { set_err(err, errlen, "recon download needs params.debug=1"); return VCPENC_E_ARGS; }
        const size_t fb = frame_bytes_of(s->p);
        const int w = g.w, h = g.h, cw = (w + 1) / 2, chh = (h + 1) / 2;
        for (int n = 0; n < N; n++) {
            const int gi = n / s->p.gop, t = n % s->p.gop;
            const size_t slot = (size_t)gi * s->ring + (t % s->ring);
            uint8_t* d = recon + (size_t)n * fb;
            CK(cudaMemcpy2DAsync(d, w, vcp_rec_luma(s->b, g, (int)slot) + g.yoff, g.ys, w, h, cudaMemcpyDeviceToHost, s->st));
            CK(cudaMemcpy2DAsync(d + (size_t)w * h, cw, s->b.rec_u + slot * g.csize + g.coff, g.cs, cw, chh, cudaMemcpyDeviceToHost, s->st));
            CK(cudaMemcpy2DAsync(d + (size_t)w * h + (size_t)cw * chh, cw, s->b.rec_v + slot * g.csize + g.coff, g.cs, cw, chh, cudaMemcpyDeviceToHost, s->st));
        }
        CK(cudaStreamSynchronize(s->st));
    }
    return VCPENC_OK;
}

int vcpenc_session_set_first_gop(vcpenc_session* s, int first_gop) {
    if (!s || first_gop < 0) return VCPENC_E_ARGS;
    s->gop_base = first_gop;
    return VCPENC_OK;
}

void* vcpenc_host_alloc(size_t bytes) {
    void* p = nullptr;
    // portable: a task sharded across several GPUs uploads from one staging buffer
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void vcpenc_host_free(void* p) { if (p) cudaFreeHost(p); }

int vcpenc_session_profile(vcpenc_session* s, int enable) {
    if (!s) return VCPENC_E_ARGS;
    s->profile = enable != 0;
    if (enable) memset(s->stats, 0, sizeof s->stats);
    return VCPENC_OK;
}

int vcpenc_session_kernel_stats(vcpenc_session* s, vcpenc_kernel_stat* stats) {
    if (!s || !stats) return VCPENC_E_ARGS;
    memcpy(stats, s->stats, sizeof s->stats);
    return VCPENC_OK;
}

int vcpenc_session_debug_mbs(vcpenc_session* s, int16_t* mv_prepass, int16_t* mv_final, uint8_t* mb_type, uint8_t* cbp) {
    if (!s || !s->p.debug || !s->encoded) return VCPENC_E_ARGS;
    cudaSetDevice(s->device);
    const size_t n = (size_t)s->nframes * s->g.nmb;
    if (mv_prepass && cudaMemcpy(mv_prepass, s->b.mvfp, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return VCPENC_E_CUDA;
    if (mv_final && cudaMemcpy(mv_final, s->dbg_mv, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return VCPENC_E_CUDA;
    if (mb_type && cudaMemcpy(mb_type, s->dbg_type, n, cudaMemcpyDeviceToHost) != cudaSuccess) return VCPENC_E_CUDA;
    if (cbp && cudaMemcpy(cbp, s->dbg_cbp, n, cudaMemcpyDeviceToHost) != cudaSuccess) return VCPENC_E_CUDA;
    return VCPENC_OK;
}

int vcpenc_encode_frames(const vcpenc_params* p, int device, const uint8_t* frames, int nframes, uint8_t* out,
                         size_t out_cap, size_t* out_len, vcpenc_frame_info* info, uint8_t* recon,
                         volatile int* cancel, char* err, size_t errlen) {
    if (!p || !frames || nframes < 1 || !out || !out_len) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    vcpenc_params q = *p;
    if (recon) q.debug = 1;
    // bound HBM use: process whole GOPs in chunks
    const size_t per_frame = 5 * frame_bytes_of(q);
    size_t budget_frames = (size_t)(48ull << 30) / std::max<size_t>(per_frame, 1);
    int chunk = (int)std::min<size_t>(budget_frames, (size_t)nframes);
    if (chunk < nframes) chunk = std::max(q.gop, chunk / q.gop * q.gop);
    vcpenc_session* s = nullptr;
    int rc = vcpenc_session_create(&q, device, std::min(chunk, nframes), &s, err, errlen);
    if (rc) return rc;
    const size_t fb = frame_bytes_of(q);
    const size_t in_fb = (size_t)vcp_in_frame_bytes(q.in_fmt, q.in_width > 0 ? q.in_width : q.width, q.in_height > 0 ? q.in_height : q.height);
    size_t o = 0;
    for (int n0 = 0; n0 < nframes && !rc; n0 += chunk) {
        if (cancel && *cancel) { set_err(err, errlen, "任务被取消"); rc = VCPENC_E_CANCELLED; break; }
        const int cnt = std::min(chunk, nframes - n0);
        s->gop_base = q.first_gop + n0 / q.gop;
        rc = vcpenc_session_upload(s, frames + (size_t)n0 * in_fb, cnt, err, errlen);
        if (!rc) rc = vcpenc_session_encode(s, nullptr, err, errlen);
        size_t len = 0;
        if (!rc) rc = vcpenc_session_download(s, out + o, out_cap - o, &len, info ? info + n0 : nullptr,
                                               recon ? recon + (size_t)n0 * fb : nullptr, err, errlen);
        if (!rc && info) for (int i = 0; i < cnt; i++) info[n0 + i].offset += o;
        o += len;
    }
    vcpenc_session_destroy(s);
    if (!rc) *out_len = o;
    return rc;
}

}  // extern "C"
