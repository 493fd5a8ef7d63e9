// conv_igemm.cu -- 3x3 'same' convolutions of the restoration CNN as IMPLICIT GEMMs on the tcgen05 tensor cores
// (reference UNet.in_conv / out_conv, src/models/convolutional.py:175-176: Conv2d(in, hidden, 3, padding="same") and
// Conv2d(hidden, in, 3, padding="same")).
//
//     out[b, y, x, n] = bias[n] + sum_{ky, kx, c} in[b, y + ky - 1, x + kx - 1, c] * w[n, c, ky, kx]
//
// is a GEMM D[pixels, N] = A[pixels, 9 C] B[N, 9 C]^T whose A operand (the 3x3 neighbourhoods) never exists in memory:
//   * an M tile is 16 x 8 pixels; for each of the nine taps ONE bulk tensor copy (channels, x, y, batch[, channel
//     block]) with the tap's coordinate offset lands the shifted 16 x 8 window in shared memory -- coordinates outside
//     the image are zero-filled by the TMA engine, which IS the 'same' padding (and clips ragged image edges);
//   * every copy is a dense [128 pixels][8 channels] block = 16 UMMA core matrices (8 rows x 16 bytes) stacked along M:
//     the canonical no-swizzle K-major layout with SBO = 128 B (next 8 pixels) and LBO = 2 KB (next 8 contraction
//     indices = the next block), so a tap (or a channel block of a tap) is one K-chunk of the GEMM and the MMAs read
//     the copies in place;
//   * the weights are rearranged once (host side, a few KB) into the same [chunk][N][8] block form and stay in shared
//     memory for the whole persistent CTA;
//   * accumulators are double-buffered in TMEM, four epilogue warps add the bias and store each pixel's N outputs as
//     contiguous bytes (a warp covers two tile rows: 2 x 1 KB runs for N = 32).
// Instantiations: <8, 32>  3 (padded to 8) -> 32 channels: in_conv forward, out_conv input gradient;
//                 <32, 16> 32 -> 3 (N padded to 16) channels: out_conv forward, in_conv input gradient.  Here a pixel is
//                 64 bytes: one copy per tap lands the window as [128 pixels][32 channels] in the SWIZZLE_64B K-major
//                 layout (8-row groups of 512 B), two K = 16 MMAs per tap.  The first version split a tap into four
//                 16-byte-wide copies (one per 8-channel block): 4608 sixteen-byte rows per tile kept the TMA engine
//                 busy for 339 us per call, slower than the direct kernel it replaced (271 us).
// The unfolded copy ([pixels, 27 -> 32], a library pad + 9-slice cat per call) and the N = 32 GEMM behind it are gone.
#include "umma.cuh"
#include <algorithm>
#include <mutex>

namespace sei {

constexpr int kIgTX = 16, kIgTY = 8;          // pixel tile (M = 128)
constexpr int kIgThreads = 192;
constexpr uint32_t kIgChunkBytes = 128 * 16;  // [128 pixels][8 channels] bf16

struct IgemmParams {
    const __nv_bfloat16* wg;      // [NCHP][NOUT][8] bf16: weights in chunk form
    const float* bias;            // [out_valid] or null
    __nv_bfloat16* out;           // [B, H, W, out_stride]
    int B, H, W, out_stride, out_valid;
    int tiles_x, tiles_y;
    long long tiles;
};

__device__ __forceinline__ void tma_load_tap_4d(void* dst, const CUtensorMap* map, int c, int x, int y, int b, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(b)
                 : "memory");
}
// K-major operand without swizzle: core matrices of 8 rows x 16 bytes; lbo = byte distance between core matrices that
// are neighbours along K, sbo = between neighbours along M / N
__device__ __forceinline__ uint64_t umma_smem_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;                                            // layout type 0: no swizzle
}

// K-major operand stored by TMA with SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row groups of 512 B (SBO), the four
// 16-byte chunks of a row XOR-ed with bits 1-2 of the row index
__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                              // leading byte offset: unused for swizzled K-major operands
    d |= (uint64_t)(512u >> 4) << 32;                    // stride byte offset: 8 rows * 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                              // layout type SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int CINP, int NOUT> struct IgemmShape {
    static constexpr int CCH = CINP / 8;                  // 8-channel blocks per tap
    static constexpr int NCH = 9 * CCH;                   // K-chunks with data
    static constexpr int NCHP = (NCH + 1) & ~1;           // padded to whole K = 16 MMAs (the extra chunk stays zero)
    static constexpr int NK16 = NCHP / 2;
    static constexpr uint32_t A_BYTES = NCHP * kIgChunkBytes;
    static constexpr uint32_t W_BYTES = NCHP * NOUT * 16;
    static constexpr int STAGES = CINP == 8 ? 4 : 2;
    static constexpr uint32_t SMEM = STAGES * A_BYTES + W_BYTES + 1024;
};

template <int CINP, int NOUT>
__global__ void __launch_bounds__(kIgThreads, CINP == 8 ? 2 : 1) conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                       const __grid_constant__ IgemmParams p)
{
    using S = IgemmShape<CINP, NOUT>;
    constexpr int STAGES = S::STAGES;
    constexpr uint32_t TMEM_COLS = 2 * NOUT < 32 ? 32 : 2 * NOUT;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* stages = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* wsm = stages + (size_t)STAGES * S::A_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // weights (already in chunk form) and the zero chunks that pad K to a multiple of 16
    for (uint32_t i = threadIdx.x; i < S::W_BYTES / 16; i += kIgThreads) {
        uint32_t dst = i;
        if (CINP == 32) {                         // [tap][n][four 16-byte chunks], chunks swizzled like the TMA's 64-byte mode
            const uint32_t tap = i / (4 * NOUT), blk = (i / NOUT) % 4, n = i % NOUT;      // global order: [tap * 4 + blk][n]
            dst = tap * (4 * NOUT) + n * 4 + (blk ^ ((n >> 1) & 3u));
        }
        reinterpret_cast<uint4*>(wsm)[dst] = __ldg(reinterpret_cast<const uint4*>(p.wg) + i);
    }
    if (S::NCHP != S::NCH) {
        for (uint32_t i = threadIdx.x; i < STAGES * (kIgChunkBytes / 16); i += kIgThreads) {
            const uint32_t s = i / (kIgChunkBytes / 16), o = i % (kIgChunkBytes / 16);
            reinterpret_cast<uint4*>(stages + (size_t)s * S::A_BYTES + (size_t)S::NCH * kIgChunkBytes)[o] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's shared-memory reads
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer: nine shifted windows per tile =====
        if (lane == 0) {
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
                const int tx = (int)(t % p.tiles_x);
                const long long r = t / p.tiles_x;
                const int ty = (int)(r % p.tiles_y), b = (int)(r / p.tiles_y);
                const uint32_t s = it % STAGES, use = it / STAGES;
                if (it >= (uint32_t)STAGES) mbar_wait(&empty_bar[s], (use - 1) & 1);
                mbar_arrive_expect_tx(&full_bar[s], S::NCH * kIgChunkBytes);
                unsigned char* dst = stages + (size_t)s * S::A_BYTES;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int x = tx * kIgTX + tap % 3 - 1, y = ty * kIgTY + tap / 3 - 1;
                    tma_load_tap_4d(dst + (size_t)tap * S::CCH * kIgChunkBytes, &map_x, 0, x, y, b, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, NOUT, false);
            const uint32_t w_addr = smem_u32(wsm);
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_use = it >> 1;
                if (it >= 2) mbar_wait(&tmem_empty_bar[acc], (acc_use - 1) & 1);
                const uint32_t s = it % STAGES;
                mbar_wait(&full_bar[s], (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(stages + (size_t)s * S::A_BYTES);
                const uint32_t tmem_d = tmem_base + acc * NOUT;
                if (CINP == 8) {
#pragma unroll
                    for (int j = 0; j < S::NK16; ++j) {
                        const uint64_t da = umma_smem_desc_kmajor_noswizzle(a_addr + (uint32_t)j * 2u * kIgChunkBytes, kIgChunkBytes, 128u);
                        const uint64_t db = umma_smem_desc_kmajor_noswizzle(w_addr + (uint32_t)j * 2u * NOUT * 16u, NOUT * 16u, 128u);
                        umma_bf16(tmem_d, da, db, idesc, j != 0);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < S::NK16; ++j) {      // tap j / 2, channels 16 (j % 2) ..: +32 B inside the 64-byte rows
                        const uint64_t da = umma_smem_desc_sw64(a_addr + (uint32_t)(j >> 1) * 4u * kIgChunkBytes + (uint32_t)(j & 1) * 32u);
                        const uint64_t db = umma_smem_desc_sw64(w_addr + (uint32_t)(j >> 1) * NOUT * 64u + (uint32_t)(j & 1) * 32u);
                        umma_bf16(tmem_d, da, db, idesc, j != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ===== epilogue: warp q owns pixels 32 q .. 32 q + 31 of the tile (tile rows 2 q, 2 q + 1) =====
        const int q = warp & 3;
        float bias[NOUT];
#pragma unroll
        for (int n = 0; n < NOUT; ++n) bias[n] = (p.bias && n < p.out_valid) ? __ldg(p.bias + n) : 0.f;
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
            const int tx = (int)(t % p.tiles_x);
            const long long r = t / p.tiles_x;
            const int ty = (int)(r % p.tiles_y), b = (int)(r / p.tiles_y);
            const uint32_t acc = it & 1, acc_use = it >> 1;
            mbar_wait(&tmem_full_bar[acc], acc_use & 1);
            tc_fence_after();
            const int m = q * 32 + lane;
            const int y = ty * kIgTY + m / kIgTX, x = tx * kIgTX + m % kIgTX;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NOUT;
            float v[NOUT];
            if (NOUT == 32) {
                uint32_t rr[32];
                tmem_ld_32x32(taddr, rr);
                tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < 32; ++n) v[n % NOUT] = __uint_as_float(rr[n]) + bias[n % NOUT];
            } else {
                uint32_t rr[16];
                tmem_ld_32x16(taddr, rr);
                tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < 16; ++n) v[n % NOUT] = __uint_as_float(rr[n]) + bias[n % NOUT];
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);           // the accumulator stage is in registers
            if (y < p.H && x < p.W) {
                __nv_bfloat16* o = p.out + (((size_t)b * p.H + y) * p.W + x) * p.out_stride;
                if (NOUT == 32 && p.out_stride == 32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), t1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]),
                                       t2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), t3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        reinterpret_cast<uint4*>(o)[j] = pk;
                    }
                } else if (p.out_stride == 4) {          // 3 (+1 zero) output channels: one 8-byte store per pixel
                    uint2 pk;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], p.out_valid > 3 ? v[3] : 0.f);
                    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                    *reinterpret_cast<uint2*>(o) = pk;
                } else {
#pragma unroll
                    for (int n = 0; n < NOUT; ++n)
                        if (n < p.out_stride) o[n] = __float2bfloat16_rn(n < p.out_valid ? v[n] : 0.f);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn ig_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// x: [B, H, W, C] bf16 as (c, x, y, b); box = C channels x 16 x 8 x 1, zero fill outside: a copy lands as
// [128 pixels][C channels] -- C = 8: 16-byte rows, no swizzle (one K-chunk); C = 32: 64-byte rows, SWIZZLE_64B
static int make_map_taps(CUtensorMap* map, const void* base, int B, int H, int W, int C)
{
    EncodeTiledFn fn = ig_encode_fn();
    SEI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t px = (cuuint64_t)C * 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {px, (cuuint64_t)W * px, (cuuint64_t)H * W * px};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)kIgTX, (cuuint32_t)kIgTY, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3x3 taps, C=%d) failed with CUresult %d", C, (int)r);
    return 0;
}

template <int CINP, int NOUT>
static int launch_igemm(const CUtensorMap& mx, const IgemmParams& p, int sm_count, cudaStream_t st)
{
    using S = IgemmShape<CINP, NOUT>;
    SEI_CUDA(allow_smem(conv3x3_igemm_kernel<CINP, NOUT>, S::SMEM));
    const unsigned grid = (unsigned)std::min<long long>(p.tiles, (long long)sm_count * (CINP == 8 ? 2 : 1));
    conv3x3_igemm_kernel<CINP, NOUT><<<grid, kIgThreads, S::SMEM, st>>>(mx, p);
    return finish_launch("conv3x3_igemm_kernel");
}

}  // namespace sei

using namespace sei;

// out[b, y, x, :] = bias + 3x3 'same' convolution of x [B, H, W, Cin] (bf16, Cin = 8 or 32) with the weights wg given in
// chunk form [NCHP][N][8] bf16 (N = 32 for Cin = 8, N = 16 for Cin = 32; chunk = tap * (Cin / 8) + channel block, tap =
// ky * 3 + kx; element [chunk][n][j] = w[n][8 * block + j][ky][kx], zero where padded).  out: bf16 [B, H, W, out_stride]
// (out_stride = 32 for N = 32; 4 for N = 16 with out_valid <= 4 channels written, the rest zero).
extern "C" int sei_conv3x3_igemm_bf16(const void* x, const void* wg, const float* bias, void* out, int B, int H, int W, int Cin,
                                      int out_stride, int out_valid, void* stream)
{
    SEI_REQUIRE(x && wg && out, "null pointer argument");
    SEI_REQUIRE(B >= 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
    SEI_REQUIRE((Cin == 8 && out_stride == 32 && out_valid <= 32) || (Cin == 32 && out_stride == 4 && out_valid <= 4),
                "implicit-GEMM 3x3 convolution: supported shapes are 8 -> 32 channels and 32 -> <= 4 channels (Cin=%d, out_stride=%d)",
                Cin, out_stride);
    SEI_REQUIRE(aligned16(x) && aligned16(wg) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, "operands must be 16-byte aligned");
    if (B == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    SEI_REQUIRE(dp.cc_major == 10, "tcgen05 convolution needs an sm_100 device");
    CUtensorMap mx;
    rc = make_map_taps(&mx, x, B, H, W, Cin);
    if (rc) return rc;
    IgemmParams p;
    p.wg = static_cast<const __nv_bfloat16*>(wg); p.bias = bias; p.out = static_cast<__nv_bfloat16*>(out);
    p.B = B; p.H = H; p.W = W; p.out_stride = out_stride; p.out_valid = out_valid;
    p.tiles_x = (W + kIgTX - 1) / kIgTX; p.tiles_y = (H + kIgTY - 1) / kIgTY;
    p.tiles = (long long)B * p.tiles_y * p.tiles_x;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (Cin == 8) return launch_igemm<8, 32>(mx, p, dp.sm_count, st);
    return launch_igemm<32, 16>(mx, p, dp.sm_count, st);
}
