// espflix_b200/csrc/ef_idct_tc.cu — the north_star's tensor-core experiment: the
// 8x8 IDCT of MpegDecoder::idct() (player.cpp:922-996) as ONE dense contraction on tcgen05 tensor cores.
//
// NOT on the decode path: the reference transform rounds inside every butterfly ((x*c + 128) >> 8), so it is not a
// linear map and no GEMM reproduces it bit for bit; K1b keeps the integer butterfly. This file measures what the
// linearised transform costs and how far it drifts: out[n][o] = rint(sum_i in[n][i] * L[o][i]) with
//   L = (P (x) P) / 256,  P = the 8-point butterfly of the reference with its constants 362/256, 473/256, 196/256
// as a [blocks x 64] x [64 x 64] GEMM, 128 blocks per CTA tile:
//   * operands in shared memory in the canonical K-major no-swizzle layout (8 rows x 16 bytes core matrices), fed by
//     TMA bulk copies (cp.async.bulk + mbarrier complete_tx) from an image a pre-pass wrote in exactly that layout;
//   * TF32 operands, split hi + lo on both sides (18-bit inputs, 22 bits of L): D = Ah*Bh + Ah*Bl + Al*Bh,
//     3 x 8 tcgen05.mma.kind::tf32 (M 128, N 64, K 8) issued by one thread, FP32 accumulator in TMEM (64 columns);
//   * tcgen05.commit -> mbarrier, epilogue: tcgen05.ld 32x32b (thread = block row, 64 columns), round to nearest, store.
// ef_idct_experiment() returns the residuals and the device time of the pre-pass and of the MMA kernel.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kTileM = 128;                       // blocks per tile (UMMA M)
constexpr int kN = 64, kK = 64;
constexpr int kAPlane = kTileM * kK * 4;          // 32 KB: one TF32 plane of a tile, chunk-major [16 chunks][128 rows][16 B]
constexpr int kBPlane = kN * kK * 4;              // 16 KB: [16 chunks][64 rows][16 B]
constexpr int kSmem = 2 * kAPlane + 2 * kBPlane + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory descriptor (cute::UMMA::SmemDescriptor, sm_100): start >> 4 | LBO >> 4 << 16 |
// SBO >> 4 << 32 | version 1 << 46. LBO = bytes between the two 16-byte K chunks of one MMA, SBO = bytes between
// 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// tf32 x tf32 -> f32, K-major A and B, M 128, N 64 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace

// pre-pass: int32 coefficient blocks -> the shared-memory image of the A operand, TF32 hi and lo planes per tile
// (value = hi + lo exactly: hi keeps the top 11 significant bits, lo the rest; |coefficient| < 2^18)
__global__ void ef_idct_tc_prep_kernel(const int32_t* __restrict__ coefs, int n_blocks, float* __restrict__ a_img)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (block row, 4-coefficient chunk)
    const int row = t >> 4, chunk = t & 15;
    const int tile = row / kTileM, r = row % kTileM;
    if (tile * kTileM >= ((n_blocks + kTileM - 1) / kTileM) * kTileM) return;
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float x = row < n_blocks ? (float)coefs[(size_t)row * 64 + chunk * 4 + k] : 0.0f;
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);       // TF32: 10 explicit mantissa bits
        hi[k] = h; lo[k] = x - h;
    }
    float* base = a_img + (size_t)tile * (2 * kAPlane / 4);
    *(float4*)(base + (chunk * kTileM + r) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *(float4*)(base + kAPlane / 4 + (chunk * kTileM + r) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(128, 1)
ef_idct_tc_kernel(const float* __restrict__ a_img, const float* __restrict__ b_img /* Bhi plane, Blo plane */, int n_tiles, int n_blocks, int32_t* __restrict__ out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                  // hi plane, lo plane
    uint8_t* sB = smem + 2 * kAPlane;
    uint64_t* bars = (uint64_t*)(smem + 2 * kAPlane + 2 * kBPlane);      // [0] operands landed, [1] MMAs done
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                     // TMEM: 64 columns x 128 lanes of FP32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    uint32_t phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) {
            // TMA-fed operand tiles: A (64 KB) every tile, the constant B (32 KB) once
            const uint32_t bytes = 2 * kAPlane + (first ? 2 * kBPlane : 0);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(bytes) : "memory");
            tma_load(sA, a_img + (size_t)tile * (2 * kAPlane / 4), 2 * kAPlane, &bars[0]);
            if (first) tma_load(sB, b_img, 2 * kBPlane, &bars[0]);
            mbar_wait(&bars[0], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // D = Ah*Bh + Ah*Bl + Al*Bh: 3 products x 8 K-steps of 8
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int prod = 0; prod < 3; prod++) {
                const uint32_t ab = a0 + (prod == 2 ? kAPlane : 0), bb = b0 + (prod == 1 ? kBPlane : 0);
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint64_t ad = make_desc(ab + ks * 2 * (kTileM * 16), kTileM * 16, 128);
                    const uint64_t bd = make_desc(bb + ks * 2 * (kN * 16), kN * 16, 128);
                    const uint32_t acc = (prod | ks) ? 1u : 0u;
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(kInstrDesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
        }
        first = false;
        // epilogue: every thread owns one block row (TMEM lane), 64 columns = its 64 residuals
        mbar_wait(&bars[1], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = tile * kTileM + warp * 32 + lane;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                           "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                           "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < n_blocks) {
                int4* dst = (int4*)(out + (size_t)row * 64 + half * 32);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    dst[q] = make_int4(__float2int_rn(__uint_as_float(v[4 * q])), __float2int_rn(__uint_as_float(v[4 * q + 1])),
                                       __float2int_rn(__uint_as_float(v[4 * q + 2])), __float2int_rn(__uint_as_float(v[4 * q + 3])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                 // TMEM and the A planes are free again
        phase ^= 1;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// host side: L planes in the B-operand image (row = output o, K = input i), entry point
static void build_b_image(const double* L /* [64 out][64 in] */, float* img /* Bhi plane then Blo plane */)
{
    for (int o = 0; o < 64; o++)
        for (int i = 0; i < 64; i++) {
            const float x = (float)L[o * 64 + i];
            union { float f; uint32_t u; } h;
            h.f = x; h.u &= 0xFFFFE000u;
            const float lo = (float)(L[o * 64 + i] - (double)h.f);
            union { float f; uint32_t u; } l;
            l.f = lo; l.u &= 0xFFFFE000u;
            const int chunk = i >> 2, e = i & 3;
            img[(chunk * kN + o) * 4 + e] = h.f;
            img[kBPlane / 4 + (chunk * kN + o) * 4 + e] = l.f;
        }
}

extern "C" int ef_idct_tc_run(int device, const int32_t* coefs, int n_blocks, const double* L, int32_t* out, float* prep_ms, float* mma_ms, int repeats)
{
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    const int n_tiles = (n_blocks + kTileM - 1) / kTileM;
    int32_t *d_in = nullptr, *d_out = nullptr;
    float *d_a = nullptr, *d_b = nullptr;
    float* h_b = new float[2 * kBPlane / 4];
    build_b_image(L, h_b);
    cudaError_t e = cudaMalloc(&d_in, (size_t)n_blocks * 256);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n_tiles * kTileM * 256);
    if (e == cudaSuccess) e = cudaMalloc(&d_a, (size_t)n_tiles * 2 * kAPlane);
    if (e == cudaSuccess) e = cudaMalloc(&d_b, 2 * kBPlane);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, coefs, (size_t)n_blocks * 256, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_b, h_b, 2 * kBPlane, cudaMemcpyHostToDevice);
    delete[] h_b;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ef_idct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaEvent_t ev[3];
    for (auto& x : ev) cudaEventCreate(&x);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float t_prep = 0, t_mma = 0;
    for (int r = 0; r < repeats && e == cudaSuccess; r++) {
        cudaEventRecord(ev[0]);
        const int threads = n_tiles * kTileM * 16;
        ef_idct_tc_prep_kernel<<<(threads + 255) / 256, 256>>>(d_in, n_blocks, d_a);
        cudaEventRecord(ev[1]);
        ef_idct_tc_kernel<<<n_tiles < sms ? n_tiles : sms, 128, kSmem>>>(d_a, d_b, n_tiles, n_blocks, d_out);
        cudaEventRecord(ev[2]);
        e = cudaDeviceSynchronize();
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
        if (r == 0 || a < t_prep) t_prep = a;
        if (r == 0 || b < t_mma) t_mma = b;
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n_blocks * 256, cudaMemcpyDeviceToHost);
    for (auto& x : ev) cudaEventDestroy(x);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_a); cudaFree(d_b);
    if (prep_ms) *prep_ms = t_prep;
    if (mma_ms) *mma_ms = t_mma;
    return e == cudaSuccess ? 0 : -2;
}
