// bgemm.cu -- batched "small dense operator x channels-last activation" product, the restoration CNN's ideal
// (Fourier-domain) resamplers as matrix products (reference: src/models/convolutional.py IdealDownsample
// :113-133, IdealUpsample :54-92).
//
// rfft2 -> fftshift -> mask / zero-pad -> irfft2 (-> stride) is a fixed linear map of the image.  Because the
// reference shifts the half-spectrum axis as well and discards its ifftshift, the map is not one separable
// product but the sum of two:  out = Gr X P^T + Gi X Q^T  (models/resample.py derives the four matrices).
// On channels-last activations both contractions are the same primitive,
//
//     D_b[M, N] = A[M, K] * X_b[K, N]        b = 0 .. batches-1,   N contiguous (channels, or width x channels)
//
// with A shared by every batch entry (A = [P; Q] over batch = image rows, then A = [Gr | Gi] over batch = images).
// The primitive is HBM-bound (K <= 512), so the design is: the A block stays resident in shared memory for the
// life of a persistent CTA, X streams through a cp.async ring of KC x BN chunks that runs ahead across
// work items, the products run on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate; ldmatrix fragments
// from padded, conflict-free rows), and every warp stages its output tile in shared memory to store full
// 16-byte vectors.  Row indices of X and D may be split (r = ro * inner + ri, two strides) so that the two-term
// intermediate [B, 2, H, W', C] is written and read in place.
#include "sei_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace sei {

// Tile shapes per resident-row count MT: threads, output columns BN per work item, X rows KC per pipeline chunk and
// ring depth.  Small operators get wide work items so that every warp still owns a 16 x 64 tile per chunk (the first
// version used 64 columns for all of them and spent its time in the per-chunk barrier); the two large ones run 16
// warps so that one CTA per SM (the A block fills most of shared memory) still hides the ldmatrix -> mma latency.
template <int MT, bool WIDE> struct BgCfg;
template <> struct BgCfg<256, false> { static constexpr int NW = 8, WM = 8, BN = 64, KC = 64, ST = 4; };   // 32 x 64 warp tiles
template <> struct BgCfg<128, false> { static constexpr int NW = 8, WM = 8, BN = 64, KC = 64, ST = 4; };   // 16 x 64
template <> struct BgCfg<64, false> { static constexpr int NW = 8, WM = 4, BN = 128, KC = 64, ST = 4; };
template <> struct BgCfg<32, false> { static constexpr int NW = 8, WM = 2, BN = 256, KC = 32, ST = 4; };
template <> struct BgCfg<16, false> { static constexpr int NW = 8, WM = 1, BN = 512, KC = 32, ST = 3; };
// WIDE (N >= 2 x the narrow BN): 64 x 64 warp tiles, 4 + 4 ldmatrix per 32 mma instead of 2 + 4 per 16 -- the A and X
// fragments are re-read from shared memory by every warp that shares them, and that traffic, not the tensor pipe,
// bounded the 32 x 64 version (ncu: profiles/)
template <> struct BgCfg<256, true> { static constexpr int NW = 8, WM = 4, BN = 128, KC = 32, ST = 4; };
template <> struct BgCfg<128, true> { static constexpr int NW = 8, WM = 2, BN = 256, KC = 32, ST = 3; };
template <> struct BgCfg<64, true> : BgCfg<64, false> {};
template <> struct BgCfg<32, true> : BgCfg<32, false> {};
template <> struct BgCfg<16, true> : BgCfg<16, false> {};

struct BgemmParams {
    const __nv_bfloat16* A;      // [m_blocks * MT][Kpad], zero padded
    const __nv_bfloat16* X;
    __nv_bfloat16* D;
    int M, K, N, Kpad;
    int b_inner, k_inner, m_inner, n_tiles;
    long long items;             // batches * n_tiles
    long long x_bo, x_bi, x_ko, x_ki;
    long long d_bo, d_bi, d_mo, d_mi;
};

// cp_async16 / cp_async_commit / cp_async_wait: sei_common.cuh

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT: rows of A resident per CTA (blockIdx.y selects the block).  Warp grid WM x WN over the MT x BN tile.
template <int MT, bool WIDE>
__global__ void __launch_bounds__(BgCfg<MT, WIDE>::NW * 32, 1) bgemm_kernel(const __grid_constant__ BgemmParams p)
{
    using Cfg = BgCfg<MT, WIDE>;
    constexpr int NT = Cfg::NW * 32, BN = Cfg::BN, KC = Cfg::KC, ST = Cfg::ST;
    constexpr int WM = Cfg::WM, WN = Cfg::NW / WM;
    constexpr int TM = MT / WM;                        // warp tile rows (32 for MT = 256, else 16)
    constexpr int TN = BN / WN;                        // warp tile columns (32 or 64)
    constexpr int MI = TM / 16, NI = TN / 8;
    constexpr int XP = BN + 8;                         // X stage pitch (bf16): rows 16 B apart mod 128 B, ldmatrix conflict-free
    constexpr int DP = TN + 8;                         // staging pitch (bf16)
    constexpr int VEC = KC * BN / 8;                   // 16-byte vectors per chunk
    static_assert(VEC % NT == 0 && NI % 2 == 0, "tile shape");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int apitch = p.Kpad + 8;
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);                           // [MT][Kpad + 8]
    __nv_bfloat16* sX = sA + (size_t)MT * apitch;                                              // [ST][KC][XP]
    __nv_bfloat16* sD = sX + (size_t)ST * KC * XP;                                             // [warps][16][DP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int m_block = blockIdx.y;
    const int kch = (p.K + KC - 1) / KC;

    // ---- resident A block (zero padded on the host: no bounds checks)
    {
        const __nv_bfloat16* gA = p.A + (size_t)m_block * MT * p.Kpad;
        const int vec_per_row = p.Kpad / 8;
        for (int i = tid; i < MT * vec_per_row; i += NT) {
            const int r = i / vec_per_row, v = i - r * vec_per_row;
            cp_async16(sA + (size_t)r * apitch + v * 8, gA + (size_t)r * p.Kpad + v * 8, true);
        }
        cp_async_commit();
    }

    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_items = first < p.items ? (p.items - first + stride - 1) / stride : 0;
    const long long total_chunks = my_items * kch;

    // one pipeline chunk: KC rows x BN columns of X (rows beyond K / columns beyond N are zero-filled)
    auto issue = [&](long long g) {
        const long long it = g / kch;
        const int kc = (int)(g - it * kch);
        const long long item = first + it * stride;
        const long long bt = item / p.n_tiles;
        const int nt = (int)(item - bt * p.n_tiles);
        const long long bo = bt / p.b_inner, bi = bt - bo * p.b_inner;
        const __nv_bfloat16* xb = p.X + bo * p.x_bo + bi * p.x_bi;
        __nv_bfloat16* dst = sX + (size_t)(g % ST) * KC * XP;
#pragma unroll
        for (int j = 0; j < VEC / NT; ++j) {
            const int idx = tid + j * NT;
            const int row = idx / (BN / 8), cv = idx % (BN / 8);
            const int k = kc * KC + row, n = nt * BN + cv * 8;
            const bool valid = k < p.K && n < p.N;
            const int ko = valid ? k / p.k_inner : 0, ki = valid ? k - ko * p.k_inner : 0;
            cp_async16(dst + row * XP + cv * 8, xb + ko * p.x_ko + ki * p.x_ki + (valid ? n : 0), valid);
        }
    };

#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < total_chunks) issue(s);
        cp_async_commit();
    }

    float acc[MI][NI][4];
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
        for (int b = 0; b < NI; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    // lane-dependent fragment addresses
    const int a_row = wm * TM + (lane & 15), a_kofs = (lane >> 4) * 8;
    const int b_krow = (lane & 15), b_nofs = wn * TN + (lane >> 4) * 8;
    __nv_bfloat16* myD = sD + (size_t)warp * 16 * DP;

    for (long long g = 0; g < total_chunks; ++g) {
        cp_async_wait<ST - 2>();
        __syncthreads();
        if (g + ST - 1 < total_chunks) issue(g + ST - 1);
        cp_async_commit();

        const long long it = g / kch;
        const int kc = (int)(g - it * kch);
        const __nv_bfloat16* xs = sX + (size_t)(g % ST) * KC * XP;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
            uint32_t af[MI][4];
#pragma unroll
            for (int a = 0; a < MI; ++a)
                ldsm_x4(af[a], sA + (size_t)(a_row + a * 16) * apitch + kc * KC + ks * 16 + a_kofs);
#pragma unroll
            for (int b = 0; b < NI / 2; ++b) {
                uint32_t bf[4];
                ldsm_x4_trans(bf, xs + (size_t)(ks * 16 + b_krow) * XP + b_nofs + b * 16);
#pragma unroll
                for (int a = 0; a < MI; ++a) {
                    mma_bf16(acc[a][2 * b], af[a], bf[0], bf[1]);
                    mma_bf16(acc[a][2 * b + 1], af[a], bf[2], bf[3]);
                }
            }
        }

        if (kc == kch - 1) {
            // ---- epilogue of this work item: fragments -> per-warp staging tile -> 16-byte global stores
            const long long item = first + it * stride;
            const long long bt = item / p.n_tiles;
            const int nt = (int)(item - bt * p.n_tiles);
            const long long bo = bt / p.b_inner, bi = bt - bo * p.b_inner;
            __nv_bfloat16* db = p.D + bo * p.d_bo + bi * p.d_bi;
            constexpr int VPR = TN / 8;                    // 16-byte vectors per tile row
#pragma unroll
            for (int a = 0; a < MI; ++a) {                 // 16 rows at a time through the per-warp staging tile
                __syncwarp();
#pragma unroll
                for (int b = 0; b < NI; ++b) {
                    const int r = lane >> 2, c = b * 8 + (lane & 3) * 2;
                    *reinterpret_cast<__nv_bfloat162*>(myD + r * DP + c) = __floats2bfloat162_rn(acc[a][b][0], acc[a][b][1]);
                    *reinterpret_cast<__nv_bfloat162*>(myD + (r + 8) * DP + c) = __floats2bfloat162_rn(acc[a][b][2], acc[a][b][3]);
                    acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;
                }
                __syncwarp();
                for (int i = lane; i < 16 * VPR; i += 32) {
                    const int r = i / VPR, v = i - r * VPR;
                    const int m = m_block * MT + wm * TM + a * 16 + r, n = nt * BN + wn * TN + v * 8;
                    if (m < p.M && n < p.N) {
                        const int mo = m / p.m_inner, mi = m - mo * p.m_inner;
                        *reinterpret_cast<uint4*>(db + mo * p.d_mo + mi * p.d_mi + n) =
                            *reinterpret_cast<const uint4*>(myD + r * DP + v * 8);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <int MT, bool WIDE> static size_t bgemm_smem_t(int Kpad)
{
    using Cfg = BgCfg<MT, WIDE>;
    constexpr int TN = Cfg::BN / (Cfg::NW / Cfg::WM);
    return ((size_t)MT * (Kpad + 8) + (size_t)Cfg::ST * Cfg::KC * (Cfg::BN + 8) + (size_t)Cfg::NW * 16 * (TN + 8)) * 2;
}

// the resident-row count is chosen with the larger (WIDE) footprint so that both variants of a tile size fit
static size_t bgemm_smem(int MT, int Kpad)
{
    switch (MT) {
    case 256: return std::max(bgemm_smem_t<256, false>(Kpad), bgemm_smem_t<256, true>(Kpad));
    case 128: return std::max(bgemm_smem_t<128, false>(Kpad), bgemm_smem_t<128, true>(Kpad));
    case 64: return bgemm_smem_t<64, false>(Kpad);
    case 32: return bgemm_smem_t<32, false>(Kpad);
    default: return bgemm_smem_t<16, false>(Kpad);
    }
}

template <int MT, bool WIDE>
static int launch_bgemm_cfg(BgemmParams p, int m_blocks, int sm_count, cudaStream_t st)
{
    using Cfg = BgCfg<MT, WIDE>;
    const size_t smem = bgemm_smem_t<MT, WIDE>(p.Kpad);
    SEI_CUDA(allow_smem(bgemm_kernel<MT, WIDE>, smem));
    p.n_tiles = (p.N + Cfg::BN - 1) / Cfg::BN;
    p.items *= p.n_tiles;                                   // caller passes items = batches
    const int by_threads = 2048 / (Cfg::NW * 32);
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(std::min(4, by_threads), (size_t)227 * 1024 / (smem + 1024)));
    const long long per_block = std::max(1, sm_count * ctas_per_sm / m_blocks);
    dim3 grid((unsigned)std::min<long long>(p.items, per_block), (unsigned)m_blocks);
    bgemm_kernel<MT, WIDE><<<grid, Cfg::NW * 32, smem, st>>>(p);
    return finish_launch("bgemm_kernel");
}

template <int MT>
static int launch_bgemm(const BgemmParams& p, int m_blocks, int sm_count, cudaStream_t st)
{
    if ((MT == 256 && p.N >= 128) || (MT == 128 && p.N >= 128)) return launch_bgemm_cfg<MT, true>(p, m_blocks, sm_count, st);
    return launch_bgemm_cfg<MT, false>(p, m_blocks, sm_count, st);
}

int bgemm_tc_try_launch(const void* A, const void* X, void* D, int M, int K, int N, int Kpad, int rows_a,
                        long long batches, int b_inner, long long x_bo, long long x_bi,
                        int k_inner, long long x_ko, long long x_ki,
                        long long d_bo, long long d_bi, int m_inner, long long d_mo, long long d_mi, cudaStream_t st);

}  // namespace sei

using namespace sei;

// rows of A resident per CTA for an (M, Kpad) operator: the smallest tile that covers M, shrunk until it fits
extern "C" int sei_bgemm_tile_rows(int M, int Kpad)
{
    DeviceProps dp;
    if (get_device_props(&dp)) return 0;
    int mt = 16;
    while (mt < M && mt < 256) mt *= 2;
    while (mt > 16 && bgemm_smem(mt, Kpad) + 1024 > (size_t)dp.smem_optin) mt /= 2;
    return bgemm_smem(mt, Kpad) + 1024 <= (size_t)dp.smem_optin ? mt : 0;
}

extern "C" int sei_bgemm_bf16(const void* A, const void* X, void* D, int M, int K, int N, int Kpad, int tile_rows,
                              long long batches, int b_inner, long long x_bo, long long x_bi,
                              int k_inner, long long x_ko, long long x_ki,
                              long long d_bo, long long d_bi, int m_inner, long long d_mo, long long d_mi, void* stream)
{
    SEI_REQUIRE(A && X && D, "null pointer argument");
    SEI_REQUIRE(M > 0 && K > 0 && N > 0 && batches >= 0, "bad shape M=%d K=%d N=%d", M, K, N);
    SEI_REQUIRE(N % 8 == 0, "N=%d must be a multiple of 8 (16-byte rows)", N);
    SEI_REQUIRE(Kpad % 64 == 0 && Kpad >= K, "Kpad=%d must be a multiple of 64 and >= K=%d", Kpad, K);
    SEI_REQUIRE(b_inner > 0 && k_inner > 0 && m_inner > 0, "inner sizes must be positive");
    SEI_REQUIRE(aligned16(A) && aligned16(X) && aligned16(D), "operands must be 16-byte aligned");
    SEI_REQUIRE(((x_bo | x_bi | x_ko | x_ki | d_bo | d_bi | d_mo | d_mi) & 7) == 0, "strides must be multiples of 8 elements");
    SEI_REQUIRE(tile_rows == sei_bgemm_tile_rows(M, Kpad) && tile_rows > 0, "tile_rows %d does not match this operator", tile_rows);
    if (batches == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    // tcgen05 / TMEM / TMA kernel (gemm.cu) for the shapes its tensor maps can express; mma.sync kernel otherwise
    {
        const int rows_a = ((M + tile_rows - 1) / tile_rows) * tile_rows;
        rc = bgemm_tc_try_launch(A, X, D, M, K, N, Kpad, rows_a, batches, b_inner, x_bo, x_bi, k_inner, x_ko, x_ki,
                                 d_bo, d_bi, m_inner, d_mo, d_mi, st);
        if (rc != 1) return rc;
    }
    BgemmParams p;
    p.A = static_cast<const __nv_bfloat16*>(A);
    p.X = static_cast<const __nv_bfloat16*>(X);
    p.D = static_cast<__nv_bfloat16*>(D);
    p.M = M; p.K = K; p.N = N; p.Kpad = Kpad;
    p.b_inner = b_inner; p.k_inner = k_inner; p.m_inner = m_inner;
    p.n_tiles = 0;
    p.items = batches;           // multiplied by the column tiles of the chosen shape in launch_bgemm
    p.x_bo = x_bo; p.x_bi = x_bi; p.x_ko = x_ko; p.x_ki = x_ki;
    p.d_bo = d_bo; p.d_bi = d_bi; p.d_mo = d_mo; p.d_mi = d_mi;
    const int m_blocks = (M + tile_rows - 1) / tile_rows;
    switch (tile_rows) {
    case 256: return launch_bgemm<256>(p, m_blocks, dp.sm_count, st);
    case 128: return launch_bgemm<128>(p, m_blocks, dp.sm_count, st);
    case 64: return launch_bgemm<64>(p, m_blocks, dp.sm_count, st);
    case 32: return launch_bgemm<32>(p, m_blocks, dp.sm_count, st);
    default: return launch_bgemm<16>(p, m_blocks, dp.sm_count, st);
    }
}
