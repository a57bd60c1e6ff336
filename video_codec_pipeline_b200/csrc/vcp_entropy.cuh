// Pieces shared by the two entropy coders (k5_cavlc.cu, k5_cabac.cu): CTA-wide scan, the
// sequential bit writer used for slice headers, the slice header itself (7.3.3) and NAL
// encapsulation with emulation prevention (7.4.1, Annex B).
#ifndef VCP_ENTROPY_CUH
#define VCP_ENTROPY_CUH

#include "vcp_dev.cuh"

namespace {

// ---- block-wide helpers ---------------------------------------------------------------------------
// exclusive prefix sum over the CTA (blockDim multiple of 32, <= 1024); returns the CTA total via `total`
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* wsum /*[33]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();  // protect wsum from the previous use
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? wsum[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < nw) wsum[lane] = wi - w;
        if (lane == 31) wsum[32] = wi;
    }
    __syncthreads();
    total = wsum[32];
    return wsum[warp] + incl - v;
}

// tiny sequential writer used by one thread for the slice header / trailer
struct SeqBits {
    uint8_t* buf; uint32_t pos;  // bit position
    __device__ __forceinline__ void put(int n, uint32_t v) {
        for (int i = n - 1; i >= 0; i--) {
            if ((v >> i) & 1) buf[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
            pos++;
        }
    }
    __device__ __forceinline__ void ue(uint32_t k) { const uint32_t x = k + 1; const int n = 31 - __clz(x); put(n, 0); put(n + 1, x); }
    __device__ __forceinline__ void se(int v) { ue(v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1)); }
};

__device__ __forceinline__ int slice_header_bits(const VcpGeom& g, int first_mb, bool idr, int frame_num, int idr_id, int qp, SeqBits* w) {
    // returns the bit count; writes when w != nullptr
    int n = 0;
#define UE(k) do { n += vcp_ue_len((unsigned)(k)); if (w) w->ue((uint32_t)(k)); } while (0)
#define SE(v) do { n += vcp_se_len(v); if (w) w->se(v); } while (0)
#define PUT(c, v) do { n += (c); if (w) w->put((c), (v)); } while (0)
    UE(first_mb);
    UE(idr ? 7 : 5);
    UE(0);
    PUT(8, (uint32_t)(frame_num & 255));
    if (idr) UE(idr_id);
    if (!idr) { PUT(1, 0); PUT(1, 0); }
    if (idr) { PUT(1, 0); PUT(1, 0); } else PUT(1, 0);
    if (g.cabac && !idr) UE(0);   // cabac_init_idc
    SE(qp - 26);
    UE(g.deblock_idc);
    if (g.deblock_idc != 1) { SE(0); SE(0); }
#undef UE
#undef SE
#undef PUT
    return n;
}

constexpr int PACK_THREADS = 256;
constexpr int PACK_BYTES = 16;   // payload bytes per thread per pass

// escape decision for 16 consecutive payload bytes starting at i0 (i0 % 16 == 0):
// byte i gets a 0x03 in front iff src[i] <= 3 and the run of zero bytes right before it has
// an even length >= 2 (equivalent to the sequential "two zeros then <= 3" rule with its reset).
__device__ __forceinline__ uint32_t escape_mask16(const uint8_t* __restrict__ src, uint32_t i0, uint32_t bytes, uint4& v) {
    v = *reinterpret_cast<const uint4*>(src + i0);   // rbsp slots are 16 B aligned and zero padded
    // zero run before byte i0
    uint32_t z = 0;
    while (z < i0 && src[i0 - 1 - z] == 0) z++;
    uint32_t mask = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t byte = (w[k >> 2] >> (8 * (k & 3))) & 255;
        if (i0 + k < bytes && byte <= 3 && z >= 2 && !(z & 1)) mask |= 1u << k;
        z = byte == 0 ? z + 1 : 0;
    }
    return mask;
}

// One CTA of PACK_THREADS wraps one slice RBSP (`bytes` bytes at `src`, 16 B aligned, readable up to
// the next multiple of 16) into a NAL unit: 00 00 00 01 | header | escaped RBSP, appended to the
// output arena; out_index[n][sl] receives (offset, size).
__device__ __forceinline__ void nal_pack_body(const VcpGeom& g, const VcpBufs& b, const uint8_t* __restrict__ src,
                                              uint32_t bytes, int n, int sl, bool idr) {
    __shared__ uint32_t wsum[33];
    __shared__ unsigned long long out_base;
    __shared__ int err_seen;
    if (threadIdx.x == 0) err_seen = *b.error_flag;
    __syncthreads();
    if (err_seen) return;
    const uint32_t npass = (bytes + PACK_THREADS * PACK_BYTES - 1) / (PACK_THREADS * PACK_BYTES);
    // pass 1: number of emulation-prevention bytes
    uint32_t cnt = 0;
    for (uint32_t ps = 0; ps < npass; ps++) {
        const uint32_t i0 = (ps * PACK_THREADS + threadIdx.x) * PACK_BYTES;
        uint4 v;
        if (i0 < bytes) cnt += __popc(escape_mask16(src, i0, bytes, v));
    }
    uint32_t nesc;
    block_excl_scan(cnt, wsum, nesc);
    const uint32_t hdr = g.hevc ? 6 : 5;   // start code + nal_unit_header (two bytes in HEVC)
    const uint32_t size = hdr + bytes + nesc;
    if (threadIdx.x == 0) {
        const unsigned long long o = atomicAdd(b.out_cursor, (unsigned long long)size);
        out_base = o;
        if (o + size > b.out_cap) atomicExch(b.error_flag, 2);
        b.out_index[(size_t)n * g.slices + sl] = make_uint2((uint32_t)o, size);
        b.out_index_hi[(size_t)n * g.slices + sl] = (uint32_t)(o >> 32);
    }
    __syncthreads();
    if (out_base + size > b.out_cap) return;
    uint8_t* dst = b.out + out_base;
    if (threadIdx.x < hdr) {
        // H.264: nal_ref_idc 3 | type 5 / nal_ref_idc 2 | type 1.  HEVC: IDR_W_RADL (19) / TRAIL_R (1), layer 0, temporal id 0
        const uint8_t h0 = g.hevc ? (uint8_t)(idr ? 19 << 1 : 1 << 1) : (uint8_t)(idr ? 0x65 : 0x41);
        dst[threadIdx.x] = threadIdx.x < 3 ? 0 : threadIdx.x == 3 ? 1 : threadIdx.x == 4 ? h0 : 1;
    }
    dst += hdr;
    uint32_t carry = 0;
    for (uint32_t ps = 0; ps < npass; ps++) {
        const uint32_t i0 = (ps * PACK_THREADS + threadIdx.x) * PACK_BYTES;
        uint4 v = make_uint4(0, 0, 0, 0);
        const uint32_t mask = i0 < bytes ? escape_mask16(src, i0, bytes, v) : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan((uint32_t)__popc(mask), wsum, tot);
        if (i0 < bytes) {
            uint8_t* d = dst + i0 + carry + ex;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t lim = bytes - i0 < PACK_BYTES ? bytes - i0 : PACK_BYTES;
            if (mask == 0 && lim == PACK_BYTES) {
#pragma unroll
                for (int k = 0; k < 16; k++) d[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            } else {
                uint32_t o = 0;
                for (uint32_t k = 0; k < lim; k++) {
                    if ((mask >> k) & 1) d[o++] = 3;
                    d[o++] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
        carry += tot;
    }
}


}  // namespace

#endif  // VCP_ENTROPY_CUH
