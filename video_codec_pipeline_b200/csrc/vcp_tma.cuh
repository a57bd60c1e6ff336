// TMA plumbing for the motion-search kernels (k2_me.cu): tensor maps over the picture planes, and the
// few PTX wrappers the kernels need (mbarrier + cp.async.bulk.tensor).
//
// Why TMA here: a search window is a small 2D box of a padded plane.  With loads, every lane computes
// addresses, fetches words and stores them to shared memory, holding registers while the data is in flight;
// with a tensor map one elected lane names the box by its coordinates and the copy engine lands it in shared
// memory -- the issue slots go to VABSDIFF4 instead of staging.
// Measured constraint (tools/tma_probe.cu, B200): the box must start on a 16-BYTE boundary of the innermost
// dimension (a u8 coordinate that is not a multiple of 16 faults with "illegal instruction"); boxes may hang
// over the tensor's edges (zero fill, the full byte count is still signalled).  So a window is fetched from the
// 16-byte boundary below it and the kernels carry the misalignment (0..15) into their shared-memory reads.
#ifndef VCP_TMA_CUH
#define VCP_TMA_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// box sizes (bytes x rows) the kernels and the maps agree on
#define VCP_L1_WIN_W 176      // half-res window of a 16-macroblock CTA: 16*8 + 2*12 + 8 = 160, + the 3 byte shifts, in 16-byte units
#define VCP_L1_WIN_H 32       // 8 + 2*12 rows
#define VCP_L1_CUR_W 128      // 16 half-res 8x8 blocks side by side
#define VCP_L0_REF_W 48       // 16 + 4 candidate columns + up to 15 bytes of misalignment
#define VCP_L0_REF_H 20       // 16 + 4 candidate rows
#define VCP_RF_WIN_W 48       // refine: x = -2 .. 17 around the vector + up to 15 bytes of misalignment
#define VCP_RF_WIN_H 20       // y = -2 .. 17
#define VCP_RFQ_WIN_W 48      // HEVC quarter-sample step: x = -4 .. 19 around the best full-sample vector + misalignment
#define VCP_RFQ_WIN_H 24      // y = -4 .. 19

struct VcpTmaps {
    CUtensorMap h_win;    // src_h  (x, row, frame)          box {VCP_L1_WIN_W, VCP_L1_WIN_H, 1}
    CUtensorMap h_cur;    // src_h                            box {VCP_L1_CUR_W, 8, 1}
    CUtensorMap y_ref;    // src_y  (x, row, frame)          box {VCP_L0_REF_W, VCP_L0_REF_H, 1}
    CUtensorMap y_cur;    // src_y                            box {16, 16, 1}
    CUtensorMap rec4;     // rec_y  (x, row, plane, slot)    box {VCP_RF_WIN_W, VCP_RF_WIN_H, 4, 1}
    CUtensorMap rec1;     // rec_y                            box {VCP_RF_WIN_W, VCP_RF_WIN_H, 1, 1}
    CUtensorMap rec1q;    // rec_y                            box {VCP_RFQ_WIN_W, VCP_RFQ_WIN_H, 1, 1}
};

// host: encode a tiled u8 map (rank 3 or 4), no swizzle, no interleave; returns 0 on success
int vcp_make_tmap(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t vcp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(vcp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(vcp_smem_u32(bar)), "r"(bytes) : "memory");
}
// Wait for phase `parity` (a probe suspends the thread for a hardware-chosen time; an explicit suspend-time hint of 2 us
// doubled me_refine's duration, measured, so none is given).  A box that never lands would wait forever: after ~2^26
// probes the kernel traps, so a bad coordinate or byte count shows up as a launch failure, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = vcp_smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
// Every lane of the waiting warp probes: the hardware parks the whole warp on the barrier.  Letting one lane wait and the
// others follow through __syncwarp() turned the wait into a divergent polling loop: me_refine 23 -> 49 ms per step (measured).
// rounded-up average of four unsigned bytes, (a + b + 1) >> 1 per byte: five plain integer instructions
// (the __vavgu4 intrinsic is emulated with about twice as many on sm_100a)
__device__ __forceinline__ uint32_t vcp_avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu); }
// generic-proxy accesses to shared memory before this point are ordered before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(vcp_smem_u32(dst)), "l"(m), "r"(vcp_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(vcp_smem_u32(dst)), "l"(m), "r"(vcp_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
#endif

#endif  // VCP_TMA_CUH
