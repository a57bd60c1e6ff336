// tma_probe — checks on a real B200 the assumptions csrc/k2_me.cu makes about cp.async.bulk.tensor:
// boxes at arbitrary byte positions, boxes partly (or mostly) outside the tensor (zero fill, full
// transaction byte count), a box wider than the tensor's innermost dimension, rank 3 and rank 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_probe tools/tma_probe.cu && build/tma_probe <case>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../video_codec_pipeline_b200/csrc/vcp_tma.cuh"

int vcp_make_tmap(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    return (int)((EncodeFn)p)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, base, d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

__device__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t a = vcp_smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 22); spin++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

__global__ void probe_kernel(const __grid_constant__ CUtensorMap m, int rank, int c0, int c1, int c2, int c3, uint32_t bytes, uint8_t* out, int* status) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        if (rank == 3) tma_load_3d(sm, &m, &bar, c0, c1, c2); else tma_load_4d(sm, &m, &bar, c0, c1, c2, c3);
    }
    const bool ok = wait_bounded(&bar, 0);
    if (threadIdx.x == 0) *status = ok ? 1 : -1;
    if (ok) for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

static uint8_t pat(uint64_t off) { return (uint8_t)((off * 2654435761u) >> 13); }

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    struct Case { int rank; uint64_t d[4]; uint32_t box[4]; int c[4]; const char* what; };
    const Case cases[] = {
        {3, {1024, 600, 3, 1}, {160, 32, 1, 1}, {5, 7, 1, 0}, "3D box inside, odd byte position"},
        {3, {1024, 600, 3, 1}, {160, 32, 1, 1}, {1000, 7, 1, 0}, "3D box partly beyond dim0"},
        {3, {64, 56, 3, 1}, {160, 32, 1, 1}, {3, 4, 0, 0}, "3D box WIDER than dim0"},
        {3, {64, 56, 3, 1}, {128, 8, 1, 1}, {16, 16, 1, 0}, "3D box wider than dim0 (cur tile)"},
        {3, {128, 112, 3, 1}, {16, 24, 1, 1}, {37, 95, 2, 0}, "3D 16x24 box, rows beyond dim1"},
        {4, {2048, 1152, 4, 2}, {32, 20, 4, 1}, {37, 41, 0, 1}, "4D box over four planes"},
        {4, {128, 112, 4, 2}, {32, 20, 1, 1}, {120, 100, 0, 1}, "4D one plane, partly beyond dims 0 and 1"},
        {3, {1024, 600, 3, 1}, {160, 32, 1, 1}, {-7, -3, 0, 0}, "3D negative coordinates"},
        {3, {1024, 600, 3, 1}, {176, 32, 1, 1}, {16, 7, 1, 0}, "3D x multiple of 16, odd row"},
        {3, {1024, 600, 3, 1}, {176, 32, 1, 1}, {1008, 590, 2, 0}, "3D x multiple of 16, beyond dims 0 and 1"},
        {3, {1024, 600, 3, 1}, {48, 20, 1, 1}, {-16, -3, 0, 0}, "3D negative coordinates, x multiple of 16"},
        {4, {2048, 1152, 4, 2}, {48, 20, 4, 1}, {32, 41, 0, 1}, "4D four planes, x multiple of 16"},
        {4, {2048, 1152, 4, 2}, {48, 20, 1, 1}, {2032, 1140, 0, 1}, "4D one plane, x multiple of 16, beyond dims 0 and 1"},
        {3, {1024, 600, 3, 1}, {48, 20, 1, 1}, {8, 5, 0, 0}, "3D x multiple of 8 only"},
    };
    const int ncases = (int)(sizeof cases / sizeof cases[0]);
    if (which < 0 || which >= ncases) { printf("cases 0..%d\n", ncases - 1); return 2; }
    const Case& k = cases[which];
    uint64_t total = 1, strides[3];
    for (int i = 0; i < k.rank; i++) { if (i) strides[i - 1] = total; total *= k.d[i]; }
    std::vector<uint8_t> h(total);
    for (uint64_t i = 0; i < total; i++) h[i] = pat(i);
    uint8_t *dsrc, *dout; int* dstat;
    cudaMalloc(&dsrc, total); cudaMemcpy(dsrc, h.data(), total, cudaMemcpyHostToDevice);
    uint32_t bytes = 1;
    for (int i = 0; i < k.rank; i++) bytes *= k.box[i];
    cudaMalloc(&dout, bytes); cudaMemset(dout, 0xEE, bytes); cudaMalloc(&dstat, 4); cudaMemset(dstat, 0, 4);
    CUtensorMap m;
    const int er = vcp_make_tmap(&m, dsrc, k.rank, k.d, strides, k.box);
    printf("case %d (%s): encode rc=%d\n", which, k.what, er);
    if (er) return 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe_kernel<<<1, 128, bytes + 128, 0>>>(m, k.rank, k.c[0], k.c[1], k.c[2], k.c[3], bytes, dout, dstat);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("  launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    int stat = 0; cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
    std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    printf("  barrier: %s\n", stat == 1 ? "completed" : "TIMED OUT (byte count never reached)");
    if (stat != 1) return 1;
    // expected: element (x0..,x1..,..) or zero when outside
    uint64_t bad = 0, idx = 0, zeros = 0;
    const uint32_t* b = k.box;
    for (uint32_t i3 = 0; i3 < (k.rank > 3 ? b[3] : 1); i3++)
        for (uint32_t i2 = 0; i2 < b[2]; i2++)
            for (uint32_t i1 = 0; i1 < b[1]; i1++)
                for (uint32_t i0 = 0; i0 < b[0]; i0++, idx++) {
                    const long long x[4] = {k.c[0] + (long long)i0, k.c[1] + (long long)i1, k.c[2] + (long long)i2, k.c[3] + (long long)i3};
                    bool in = true; uint64_t off = 0, mul = 1;
                    for (int q = 0; q < k.rank; q++) { if (x[q] < 0 || x[q] >= (long long)k.d[q]) in = false; off += (uint64_t)(in ? x[q] : 0) * mul; mul *= k.d[q]; }
                    const uint8_t want = in ? h[off] : 0;
                    zeros += !in;
                    if (o[idx] != want) bad++;
                }
    printf("  content: %llu mismatches of %u bytes (%llu outside the tensor, expected zero)\n", (unsigned long long)bad, bytes, (unsigned long long)zeros);
    return bad ? 1 : 0;
}
