// gemm.cu -- bf16 x bf16 -> fp32-accumulate GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// the dense contraction of the restoration CNN (reference: src/models/convolutional.py, the 1x1
// convolutions :40-42,106,143 and, through an unfold, the 3x3 in/out convolutions :175-176).
//
//     D[M, N] = A[M, K] * B[N, K]^T (+ bias[N])          A, B: bf16 row-major (K contiguous, "TN")
//
// A pointwise convolution on channels-last activations is exactly this GEMM with M = B*H*W pixels,
// K = C_in, N = C_out, B = the (C_out, C_in) weight -- no layout change, and dgrad / wgrad are the
// same kernel with the operand roles permuted.
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0   TMA producer: one elected lane streams 128x64 (A) and BNx64 (B) bf16 boxes into a
//            ring of shared-memory stages (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier tx completion)
//   warp 1   allocates BN TMEM columns, then one elected lane issues tcgen05.mma (UMMA 128 x BN x 16,
//            both operands from shared-memory descriptors, fp32 accumulator in TMEM) and releases each
//            stage with tcgen05.commit
//   warps 2-5  epilogue: tcgen05.ld the accumulator (each warp its 32-lane quarter), add the bias,
//            convert, store 16-byte vectors
// Out-of-range rows/columns/k are zero-filled by TMA on load and masked on store.
#include "sei_common.cuh"
#include "umma.cuh"
#include "gelu.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace sei {

constexpr int kGemmThreads = 192;

struct GemmParams {
    void* D;
    const float* bias;
    int M, N, K, ldd;
    int kb_per_split;    // k-blocks per grid.z slice (split-K: fp32 output only, slices accumulate with atomics)
    int splits;
    int tma_store;       // bf16 output staged through shared memory and written with bulk tensor stores
    int accumulate;      // fp32 output: add to D instead of overwriting it (weight gradients accumulated in place)
    const __nv_bfloat16* gelu_h;   // optional [M, N] (row pitch ld_h): D = (A B^T) * gelu'(gelu_h)  (TMA-store epilogue only)
    int ld_h;
    const __nv_bfloat16* res;      // optional [M, N] (row pitch ld_r): D = A B^T + bias + res_scale * res  (bf16 output),
    int ld_r;                      //   or, with res_mul, D = (A B^T + bias) * res
    float res_scale;
    int res_mul;
    int gelu_dual;                 // D = gelu(A B^T + bias) and D2 (second tensor map) = gelu'(A B^T + bias)
    int raster_gn;                 // CTA-pair kernels: n-tiles per group of the tile order (raster_tile)
    const float* row_scale;        // optional [row_period]: the bias of row r is bias[n] * row_scale[r % row_period]
    int row_period;
};



__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8])
{
    uint4 pk;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]),
                   t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
    return pk;
}

// The GELU backward fused into the input-gradient GEMM of the layer that follows it (reference ConvBlock: conv2 -> GELU
// -> conv3): the epilogue thread that owns a row chunk of dL/da = gy W3 multiplies it by gelu'(h) read from the saved
// pre-activation, so dL/dh is written directly (no separate pass over the 4C-wide gradient).
__device__ __forceinline__ void epilogue_gelu_bwd(const GemmParams& p, int row, int col0, uint32_t (&r0)[32], uint32_t (&r1)[32])
{
    if (row >= p.M) return;
    const uint4* hp = reinterpret_cast<const uint4*>(p.gelu_h + (size_t)row * p.ld_h + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (col0 + 8 * j + 8 > p.N) break;
        const uint4 raw = __ldg(hp + j);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 hv = __bfloat1622float2(h2[e]);
            float Phi, phi;
            const int cc = 8 * j + 2 * e;
            gelu_parts(hv.x, Phi, phi);
            const float d0 = fmaf(hv.x, phi, Phi);
            gelu_parts(hv.y, Phi, phi);
            const float d1 = fmaf(hv.y, phi, Phi);
            if (cc < 32) {
                r0[cc] = __float_as_uint(__uint_as_float(r0[cc]) * d0);
                r0[cc + 1] = __float_as_uint(__uint_as_float(r0[cc + 1]) * d1);
            } else {
                r1[cc - 32] = __float_as_uint(__uint_as_float(r1[cc - 32]) * d0);
                r1[cc - 31] = __float_as_uint(__uint_as_float(r1[cc - 31]) * d1);
            }
        }
    }
}

// epilogue of one 32-column chunk held in registers: bias, convert, store (masked at the matrix edge)
template <bool OUT_F32>
__device__ __forceinline__ void gemm_store_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, int col0, bool add_bias)
{
    if (row >= p.M || col0 >= p.N) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias && add_bias) {
        const float rsc = p.row_scale ? __ldg(p.row_scale + (row % p.row_period)) : 1.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] = fmaf(rsc, __ldg(p.bias + col0 + j), v[j]);
    }
    if (!OUT_F32 && p.res) {
        const __nv_bfloat16* rp = p.res + (size_t)row * p.ld_r + col0;
        if (col0 + 32 <= p.N && (p.ld_r & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(rp + j));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[j + 2 * e] = fmaf(p.res_scale, f.x, v[j + 2 * e]);
                    v[j + 2 * e + 1] = fmaf(p.res_scale, f.y, v[j + 2 * e + 1]);
                }
            }
        } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) v[j] = fmaf(p.res_scale, __bfloat162float(rp[j]), v[j]);
        }
    }
    if (OUT_F32 && (p.splits > 1 || p.accumulate)) {
        float* dst = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0;
        if (col0 + 32 <= p.N && (p.ldd & 3) == 0) {
            // 16-byte vector reductions (red.global.add.v4.f32): a lane owns one row, so scalar atomics would touch
            // 32 sectors with 4 bytes each per instruction
#pragma unroll
            for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
        }
    } else if (OUT_F32) {
        float* dst = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0;
        if (col0 + 32 <= p.N && (p.ldd & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = v[j];
        }
    } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col0;
        if (col0 + 32 <= p.N && (p.ldd & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]),
                               t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + j) = pk;
            }
        } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = __float2bfloat16_rn(v[j]);
        }
    }
}


// Epilogue of one 32-row x 64-column chunk of the accumulator on the bulk-tensor-store path (one warp): tcgen05.ld the
// chunk, add bias / residual, convert to bf16, stage it in the warp's SWIZZLE_128B slab (16-byte chunk j of row r at
// position j ^ (r & 7): conflict-free) and hand the slab to the TMA engine.  The bias row of the chunk is staged once in
// shared memory by the warp (two coalesced loads, then 16 broadcast LDS.128 per lane) -- the first version issued 64
// predicated scalar loads + compares + adds per lane and chunk, three times the instructions of the conversion itself
// -- and the residual row chunk (128 contiguous bytes per lane) is requested BEFORE the accumulator wait so that the two
// latencies overlap.
// The bias values of a warp's column chunks, staged in shared memory when the tile's column offset changes (with m-fastest
// / grouped tile orders: rarely).  The first round-2 version reloaded the chunk's 64 values from global memory for every
// chunk, which put one L2 round trip into the latency chain of every 4 KB store.
template <int NCH>
__device__ __forceinline__ void gemm_stage_bias(const GemmParams& p, float* sbw, int n0, int sub, int nsub, int lane)
{
    __syncwarp();
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int col0 = n0 + 64 * (sub + ci * nsub);
        sbw[ci * 64 + lane] = col0 + lane < p.N ? __ldg(p.bias + col0 + lane) : 0.f;
        sbw[ci * 64 + 32 + lane] = col0 + 32 + lane < p.N ? __ldg(p.bias + col0 + 32 + lane) : 0.f;
    }
    __syncwarp();
}

template <int NSLAB>
__device__ __forceinline__ void gemm_epilogue_chunk64(const GemmParams& p, const CUtensorMap* map_d, const CUtensorMap* map_d2,
                                                      uint32_t taddr, int row0, int col0, int lane, bool add_bias,
                                                      unsigned char* slab, float* sb)
{
    const int row = row0 + lane;
    uint4 rr[8];
    const bool has_res = p.res != nullptr;
    if (has_res) {
        // The residual / multiplier chunk (32 rows x 128 B) is copied into the warp's own slab with coalesced 16-byte
        // asynchronous copies (a lane moves chunk lane % 8 of rows lane / 8 + 4 i: four whole rows per instruction, the
        // slab's swizzled layout) while the accumulator is fetched; each lane then reads ITS row back.  Per-lane row
        // loads straight from global memory touch 32 sectors per instruction and re-fetch every sector (the kernel
        // leaves no room for L1): the input-gradient GEMM with the multiplier took 598 us against 162 us without.
        if (lane == 0) bulk_wait_read<0>();      // the slab is free (the previous store has read it)
        __syncwarp();
        const int ch = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = (lane >> 3) + 4 * i;
            const bool ok = row0 + rl < p.M && col0 + 8 * ch + 8 <= p.N;
            cp_async16(slab + rl * 128 + ((ch ^ (rl & 7)) << 4),
                       p.res + (ok ? (size_t)(row0 + rl) * p.ld_r + col0 + 8 * ch : 0), ok);
        }
        cp_async_commit();
    }
    add_bias = add_bias && sb != nullptr;          // sb: the chunk's 64 bias values, staged by the caller (see gemm_stage_bias)
    // bias behind a resampler (Downsample with the resampler applied first): the constant image bias[n] became
    // bias[n] * R(1)[pixel], a per-row factor
    const float rsc = (add_bias && p.row_scale) ? __ldg(p.row_scale + (min(row, p.M - 1) % p.row_period)) : 1.0f;
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(taddr, r0);
    tmem_ld_32x32(taddr + 32u, r1);
    tmem_ld_wait();
    if (p.gelu_h) epilogue_gelu_bwd(p, row, col0, r0, r1);
    if (has_res) {
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) rr[j] = *reinterpret_cast<const uint4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4));
    }
    if (lane == 0) bulk_wait_read<NSLAB - 1>();   // the store that last read this slab has finished reading
    __syncwarp();                            // (also: the bias row is visible to every lane)
    const bool dual = p.gelu_dual != 0;
    uint4 dpk[8];                            // gelu'(h) of the chunk (dual mode), stored after the activations
#pragma unroll
    for (int j = 0; j < 8; ++j) {            // 8 chunks of 8 bf16
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int cc = 8 * j + e;
            v[e] = __uint_as_float(cc < 32 ? r0[cc] : r1[cc - 32]);
        }
        if (add_bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(sb + 8 * j), b1 = *reinterpret_cast<const float4*>(sb + 8 * j + 4);
            v[0] = fmaf(rsc, b0.x, v[0]); v[1] = fmaf(rsc, b0.y, v[1]); v[2] = fmaf(rsc, b0.z, v[2]); v[3] = fmaf(rsc, b0.w, v[3]);
            v[4] = fmaf(rsc, b1.x, v[4]); v[5] = fmaf(rsc, b1.y, v[5]); v[6] = fmaf(rsc, b1.z, v[6]); v[7] = fmaf(rsc, b1.w, v[7]);
        }
        if (has_res) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
            if (p.res_mul) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[2 * e] *= f.x;
                    v[2 * e + 1] *= f.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[2 * e] = fmaf(p.res_scale, f.x, v[2 * e]);
                    v[2 * e + 1] = fmaf(p.res_scale, f.y, v[2 * e + 1]);
                }
            }
        }
        if (dual) {
            float dv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                // the pre-activation is rounded to bf16 first: that is the value the unfused path stores and both of its
                // passes read
                const float hq = __bfloat162float(__float2bfloat16_rn(v[e]));
                float Phi, phi;
                gelu_parts(hq, Phi, phi);
                dv[e] = fmaf(hq, phi, Phi);
                v[e] = hq * Phi;
            }
            dpk[j] = pack8_bf16(dv);
        }
        *reinterpret_cast<uint4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v);
    }
    fence_proxy_async();                     // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    const bool in_range = col0 < p.N && row0 < p.M;
    if (lane == 0 && in_range) {
        tma_store_2d(map_d, slab, col0, row0);
        bulk_commit();
    }
    if (dual) {
        if (lane == 0) bulk_wait_read<0>();  // the activations have left the slab
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = dpk[j];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && in_range) {
            tma_store_2d(map_d2, slab, col0, row0);
            bulk_commit();
        }
    }
}

// Persistent kernel: gridDim.x CTAs (one per SM) walk the work items (output tile x K-split), m fastest so that the
// CTAs running together share the B (weight) tile through L2.  The shared-memory ring and its mbarrier phases
// run continuously across tiles, and the accumulator is double-buffered in TMEM (2 x BN columns): the epilogue
// warps drain tile i (tcgen05.ld -> bias -> store) while the MMA thread is already accumulating tile i+1.
// MN = false: A[M,K], B[N,K] row-major (K contiguous).  MN = true: A[K,M], B[K,N] row-major (the contraction index is
// the row: D = A^T B), used by the weight gradient so that gy and x are read in place (no transposed copies).
// EW epilogue warps (4 or 8): warp 2 + i owns TMEM lane quarter (2 + i) % 4 and every (EW / 4)-th column chunk.  A
// short-K tile finishes its MMAs in a few hundred clocks and then waits for 32 - 64 KB of output to leave: with four
// warps that drain (tcgen05.ld -> convert -> slab -> bulk store, each step waiting on the one before) bounded the
// expanding pointwise convolutions of the shallow levels at half of the HBM rate.
template <int BN, int STAGES, bool OUT_F32, bool MN = false, int EW = 4, int MINB = 1>
__global__ void __launch_bounds__(64 + 32 * EW, MINB)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_d2,
                    const __grid_constant__ GemmParams p)
{
    constexpr int BM = kGemmBM, BK = kGemmBK;
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;               // two accumulator stages; 64..512, a power of two
    constexpr uint32_t SLAB_BYTES = 32 * 128;            // epilogue staging: 32 rows x 64 bf16 per warp and buffer
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    constexpr int BCH = BN >= 64 ? (BN / 64 + EW / 4 - 1) / (EW / 4) : 1;     // 64-column chunks per epilogue warp
    __shared__ __align__(16) float bias_s[EW * BCH * 64];

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);      // SWIZZLE_128B tiles: 1024-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int nk_total = (p.K + BK - 1) / BK;
    const long long tiles_mn = (long long)tiles_m * tiles_n;
    const long long total = tiles_mn * p.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 32 * EW);       // every epilogue thread arrives
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (long long w = blockIdx.x; w < total; w += gridDim.x) {
                const int split = (int)(w / tiles_mn);
                const long long rem = w - split * tiles_mn;
                const int m0 = (int)(rem % tiles_m) * BM, n0 = (int)(rem / tiles_m) * BN;
                const int kb0 = split * p.kb_per_split;
                const int nk = min(nk_total, kb0 + p.kb_per_split) - kb0;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES, use = it / STAGES;
                    if (it >= STAGES) mbar_wait(&empty_bar[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    unsigned char* a_dst = tiles + (size_t)s * STAGE_BYTES;
                    if (MN) {
                        // boxes of 64 MN-elements x BK k-rows (8 KB each), one per 64-element MN block
#pragma unroll
                        for (int mb = 0; mb < BM / 64; ++mb)
                            tma_load_2d(a_dst + mb * (BK * 128), &map_a, m0 + 64 * mb, (kb0 + kb) * BK, &full_bar[s]);
#pragma unroll
                        for (int nb = 0; nb < BN / 64; ++nb)
                            tma_load_2d(a_dst + A_BYTES + nb * (BK * 128), &map_b, n0 + 64 * nb, (kb0 + kb) * BK, &full_bar[s]);
                    } else {
                        tma_load_2d(a_dst, &map_a, (kb0 + kb) * BK, m0, &full_bar[s]);
                        tma_load_2d(a_dst + A_BYTES, &map_b, (kb0 + kb) * BK, n0, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, MN);
            uint32_t it = 0, tcount = 0;
            for (long long w = blockIdx.x; w < total; w += gridDim.x, ++tcount) {
                const int split = (int)(w / tiles_mn);
                const int kb0 = split * p.kb_per_split;
                const int nk = min(nk_total, kb0 + p.kb_per_split) - kb0;
                const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
                if (tcount >= 2) mbar_wait(&tmem_empty_bar[acc], (acc_use - 1) & 1);     // epilogue drained this stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    if (MN) {
                        const uint64_t da = umma_smem_desc_mn_sw128(a_addr, BK * 128),
                                       db = umma_smem_desc_mn_sw128(a_addr + A_BYTES, BK * 128);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)   // 16 k-rows = 2048 B further = +128 in the (addr >> 4) field
                            umma_bf16(tmem_d, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                    } else {
                        const uint64_t da = umma_smem_desc_sw128(a_addr), db = umma_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)   // +32 B along K inside the swizzle row = +2 in the (addr >> 4) field
                            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);              // frees the stage when these MMAs have read it
                }
                umma_commit(&tmem_full_bar[acc]);            // accumulator stage complete
            }
        }
    } else {
        // ===== epilogue: warps 2.. own TMEM lane quarters (warp % 4) and interleaved column chunks =====
        const int q = warp & 3, sub = (warp - 2) >> 2;
        constexpr int NSUB = EW / 4;
        uint32_t tcount = 0, chunk_count = 0;
        constexpr int NSLAB = EW == 4 ? 2 : 1;     // staging slabs per warp (with eight warps the other warps provide the overlap)
        unsigned char* slabs = tiles + (size_t)STAGES * STAGE_BYTES + (size_t)(warp - 2) * NSLAB * SLAB_BYTES;
        float* sbw = bias_s + (warp - 2) * BCH * 64;
        int bias_n0 = -1;
        for (long long w = blockIdx.x; w < total; w += gridDim.x, ++tcount) {
            const int split = (int)(w / tiles_mn);
            const long long rem = w - split * tiles_mn;
            const int m0 = (int)(rem % tiles_m) * BM, n0 = (int)(rem / tiles_m) * BN;
            const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
            const bool tma_path = !OUT_F32 && p.tma_store && BN >= 64;
            if (tma_path && p.bias && n0 != bias_n0) {       // (before the accumulator wait: the loads overlap the MMAs)
                gemm_stage_bias<BCH>(p, sbw, n0, sub, NSUB, lane);
                bias_n0 = n0;
            }
            mbar_wait(&tmem_full_bar[acc], acc_use & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            if (tma_path) {
                int ci = 0;
#pragma unroll 1
                for (int c = 64 * sub; c < BN; c += 64 * NSUB, ++ci)
                    gemm_epilogue_chunk64<NSLAB>(p, &map_d, &map_d2, taddr + (uint32_t)c, m0 + q * 32, n0 + c, lane, split == 0,
                                                 slabs + (size_t)(chunk_count++ % NSLAB) * SLAB_BYTES,
                                                 p.bias ? sbw + ci * 64 : nullptr);
            } else {
#pragma unroll 1
                for (int c = 32 * sub; c < BN; c += 32 * NSUB) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + (uint32_t)c, r);
                    tmem_ld_wait();
                    gemm_store_chunk<OUT_F32>(p, r, row, n0 + c, split == 0);
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);               // this thread is done reading the accumulator stage
        }
    }
    if (!OUT_F32 && p.tma_store && warp >= 2 && lane == 0) bulk_wait_all();   // outstanding stores still read shared memory
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- CTA-pair kernel (cta_group::2)
// A 128 x 256 tile per SM reads (128 + 256) x 16 bf16 = 12 KB of shared memory per UMMA, 192 B per tensor-core clock
// against the 128 B/clk an SM's shared memory delivers: the single-CTA kernel above tops out near 70 % of the tensor
// peak (measured: 1.20 of 1.69 PFLOP/s).  Here two CTAs of a cluster (one TPC) share a 256 x 256 tile: each stages
// its own 128 rows of A and HALF of B (128 of the 256 columns), the leader CTA issues one tcgen05.mma.cta_group::2
// (M = 256) that reads both halves in place, and each CTA's TMEM receives its 128 rows of the accumulator.  Shared
// memory traffic per CTA drops to 8 KB per UMMA = 128 B/clk.
//   * both CTAs run a TMA producer; every load signals the LEADER's full barrier (peer bit cleared), which expects
//     the bytes of both CTAs;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arrives on the empty / accumulator-full barriers of BOTH CTAs;
//   * the epilogue threads of both CTAs arrive on the leader's accumulator-empty barrier (256 arrivals).
// TN operands (K contiguous), bf16 output through the TMA-store epilogue, no split-K: the forward and input-gradient
// GEMMs of the deep layers.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask),
                   "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all earlier MMAs of this thread completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}


// Tile order of the CTA-pair kernels: n-tiles in groups of kRasterGN; inside a group n runs fastest, then m; then the
// next group.  The 74 clusters that run together then cover ~9 m-tiles x 8 n-tiles (68 MB of
// operand strips at K = 8192: they fit the 126 MB L2), the group's B strips stay resident while A streams through once
// per group.  With plain m-fastest order a wave touched EVERY A strip (134 MB > L2) and 2-3 B strips: ncu measured
// 9.2 GB of DRAM reads per launch of the deepest layer against 0.67 GB of operands (profiles/r02_ncu_gemm_2cta_s4*).
constexpr int kRasterGN = 8;
__device__ __forceinline__ void raster_tile(long long w, int tiles_m, int tiles_n, int gn, int& mt, int& nt)
{
    const long long per_group = (long long)gn * tiles_m;
    const int group = (int)(w / per_group);
    const int in_group = (int)(w - (long long)group * per_group);
    const int gn0 = group * gn;
    const int gsize = min(gn, tiles_n - gn0);
    mt = in_group / gsize;
    nt = gn0 + in_group - mt * gsize;
}

// B_MN: the B operand is given as [K, N] row-major (N contiguous: the UMMA MN-major layout, TMA boxes of 64 n x BK k-rows,
// two per CTA) instead of [N, K] -- D = A B.  The input-gradient GEMMs read the weight matrix in place that way; round 1
// kept a transposed bf16 copy of every weight (1.3 GB, rewritten by the optimizer every step: 1.8 ms of transposes).
template <int STAGES, int EW = 4, bool B_MN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_d2,
                         const __grid_constant__ GemmParams p)
{
    constexpr int BM = kGemmBM, BK = kGemmBK, BN = 256, BNH = 128;      // per CTA: 128 rows of A, 128 columns of B
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;               // two accumulator stages of 256 fp32 columns
    constexpr uint32_t SLAB_BYTES = 32 * 128;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    constexpr int BCH = (256 / 64 + EW / 4 - 1) / (EW / 4);      // 64-column chunks per epilogue warp
    __shared__ __align__(16) float bias_s[EW * BCH * 64];

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = (p.N + BN - 1) / BN;
    const int nk = (p.K + BK - 1) / BK;
    const long long total = (long long)tiles_m * tiles_n;
    const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 64 * EW);       // the epilogue threads of both CTAs
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own rows of A, own half of B; completion on the leader's barrier =====
        if (lane == 0) {
            uint32_t it = 0;
            for (long long w = cluster_id; w < total; w += n_clusters) {
                int mt, nt;
                raster_tile(w, tiles_m, tiles_n, p.raster_gn, mt, nt);
                const int m0 = mt * 2 * BM + (int)rank * BM, n0 = nt * BN + (int)rank * BNH;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES, use = it / STAGES;
                    if (it >= STAGES) mbar_wait(&empty_bar[s], (use - 1) & 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    unsigned char* a_dst = tiles + (size_t)s * STAGE_BYTES;
                    tma_load_2d_2sm(a_dst, &map_a, kb * BK, m0, &full_bar[s]);
                    if (B_MN) {
#pragma unroll
                        for (int nb = 0; nb < BNH / 64; ++nb)
                            tma_load_2d_2sm(a_dst + A_BYTES + nb * (BK * 128), &map_b, n0 + 64 * nb, kb * BK, &full_bar[s]);
                    } else {
                        tma_load_2d_2sm(a_dst + A_BYTES, &map_b, kb * BK, n0, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = B_MN ? umma_idesc_bf16_kmn(2 * BM, BN) : umma_idesc_bf16(2 * BM, BN, false);
            uint32_t it = 0, tcount = 0;
            for (long long w = cluster_id; w < total; w += n_clusters, ++tcount) {
                const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
                if (tcount >= 2) mbar_wait(&tmem_empty_bar[acc], (acc_use - 1) & 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t da = umma_smem_desc_sw128(a_addr);
                    const uint64_t db = B_MN ? umma_smem_desc_mn_sw128(a_addr + A_BYTES, BK * 128) : umma_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)     // B_MN: 16 k-rows = 2048 B further = +128 in the (addr >> 4) field
                        umma_bf16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(B_MN ? 128 * k : 2 * k), idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty_bar[s]);              // frees this stage in both CTAs
                }
                umma_commit_2sm(&tmem_full_bar[acc]);            // accumulator complete: both epilogues
            }
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows of the accumulator =====
        const int q = warp & 3, sub = (warp - 2) >> 2;
        constexpr int NSUB = EW / 4;
        uint32_t tcount = 0, chunk_count = 0;
        constexpr int NSLAB = EW == 4 ? 2 : 1;
        unsigned char* slabs = tiles + (size_t)STAGES * STAGE_BYTES + (size_t)(warp - 2) * NSLAB * SLAB_BYTES;
        float* sbw = bias_s + (warp - 2) * BCH * 64;
        int bias_n0 = -1;
        for (long long w = cluster_id; w < total; w += n_clusters, ++tcount) {
            int mt, nt;
            raster_tile(w, tiles_m, tiles_n, p.raster_gn, mt, nt);
            const int m0 = mt * 2 * BM + (int)rank * BM, n0 = nt * BN;
            const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
            if (p.bias && n0 != bias_n0) {
                gemm_stage_bias<BCH>(p, sbw, n0, sub, NSUB, lane);
                bias_n0 = n0;
            }
            mbar_wait(&tmem_full_bar[acc], acc_use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            int ci = 0;
#pragma unroll 1
            for (int c = 64 * sub; c < BN; c += 64 * NSUB, ++ci)
                gemm_epilogue_chunk64<NSLAB>(p, &map_d, &map_d2, taddr + (uint32_t)c, m0 + q * 32, n0 + c, lane, true,
                                             slabs + (size_t)(chunk_count++ % NSLAB) * SLAB_BYTES,
                                             p.bias ? sbw + ci * 64 : nullptr);
            tc_fence_before();
            mbar_arrive_leader(&tmem_empty_bar[acc]);       // this thread is done reading the accumulator stage
        }
        if (lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();                                    // the peer's shared memory / barriers stay valid until both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

// CTA-pair kernel for the weight gradient: D[M, N] (fp32) (+)= A[K, M]^T B[K, N], both operands MN-major (read in place),
// M = 256 per cluster (128 rows of D per CTA), N = 256 (each CTA stages 128 of the 256 columns of B), split-K across
// clusters with vector-atomic accumulation.  Same barrier protocol as gemm_bf16_tn_2cta_kernel; the operand tiles are
// the MN-major boxes of the single-CTA kernel (64 MN-elements x BK k-rows, 8 KB each): two per operand and CTA.
template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_atb_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const __grid_constant__ GemmParams p)
{
    constexpr int BM = kGemmBM, BK = kGemmBK, BN = 256, BNH = 128;
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t BOX_BYTES = BK * 128;             // one 64-element MN block
    constexpr uint32_t TMEM_COLS = 2 * BN;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = (p.N + BN - 1) / BN;
    const int nk_total = (p.K + BK - 1) / BK;
    const long long tiles_mn = (long long)tiles_m * tiles_n;
    const long long total = tiles_mn * p.splits;
    const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 256);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (long long w = cluster_id; w < total; w += n_clusters) {
                const int split = (int)(w / tiles_mn);
                const long long rem = w - split * tiles_mn;
                int mt, nt;
                raster_tile(rem, tiles_m, tiles_n, p.raster_gn, mt, nt);
                const int m0 = mt * 2 * BM + (int)rank * BM, n0 = nt * BN + (int)rank * BNH;
                const int kb0 = split * p.kb_per_split;
                const int nk = min(nk_total, kb0 + p.kb_per_split) - kb0;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES, use = it / STAGES;
                    if (it >= STAGES) mbar_wait(&empty_bar[s], (use - 1) & 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    unsigned char* a_dst = tiles + (size_t)s * STAGE_BYTES;
                    const int k0 = (kb0 + kb) * BK;
#pragma unroll
                    for (int mb = 0; mb < BM / 64; ++mb) tma_load_2d_2sm(a_dst + mb * BOX_BYTES, &map_a, m0 + 64 * mb, k0, &full_bar[s]);
#pragma unroll
                    for (int nb = 0; nb < BNH / 64; ++nb)
                        tma_load_2d_2sm(a_dst + A_BYTES + nb * BOX_BYTES, &map_b, n0 + 64 * nb, k0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, true);
            uint32_t it = 0, tcount = 0;
            for (long long w = cluster_id; w < total; w += n_clusters, ++tcount) {
                const int split = (int)(w / tiles_mn);
                const int kb0 = split * p.kb_per_split;
                const int nk = min(nk_total, kb0 + p.kb_per_split) - kb0;
                const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
                if (tcount >= 2) mbar_wait(&tmem_empty_bar[acc], (acc_use - 1) & 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t da = umma_smem_desc_mn_sw128(a_addr, BOX_BYTES), db = umma_smem_desc_mn_sw128(a_addr + A_BYTES, BOX_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_2sm(tmem_d, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty_bar[s]);
                }
                umma_commit_2sm(&tmem_full_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t tcount = 0;
        for (long long w = cluster_id; w < total; w += n_clusters, ++tcount) {
            const int split = (int)(w / tiles_mn);
            const long long rem = w - split * tiles_mn;
            int mt, nt;
            raster_tile(rem, tiles_m, tiles_n, p.raster_gn, mt, nt);
            const int m0 = mt * 2 * BM + (int)rank * BM, n0 = nt * BN;
            const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
            mbar_wait(&tmem_full_bar[acc], acc_use & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c, r);
                tmem_ld_wait();
                gemm_store_chunk<true>(p, r, row, n0 + c, false);
            }
            tc_fence_before();
            mbar_arrive_leader(&tmem_empty_bar[acc]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- batched operator product on tcgen05
// D_b[M, N] = A[M, K] * X_b[K, N]  (csrc/bgemm.cu; the CNN's ideal resamplers).  The operator A (M <= 256 rows per
// CTA, M * K <= 64 K elements) is loaded ONCE per persistent CTA into shared memory as K-major SWIZZLE_128B tiles and
// is the UMMA "A" operand; X_b is the "B" operand read in place as an MN-major tile (its rows are the contraction
// index, its columns contiguous), streamed through a TMA ring; the accumulators (one 128-row block per 128 rows of A,
// double-buffered) live in TMEM and leave through the swizzled-slab / TMA-store epilogue.  Batch entries and the
// split row indices of X and D (r = ro * inner + ri) are dimensions of 5-D tensor maps, so the two-term resampler
// intermediate [B, 2, H, W', C] is read and written in place.  HBM-bound by construction: a 256 x 128 x 256 item is
// 1 k tensor-core clocks for 128 KB of traffic.
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
                   "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
struct BgemmTcParams {
    int M, K, N;                 // logical sizes
    int mb, kb;                  // 128-row blocks of A per CTA (1 or 2), 64-column k-blocks
    int n_tiles, b_inner;
    int k_inner, m_inner;        // split sizes of the X rows / D rows
    long long items;             // batches * n_tiles
};

// RES: the operator block is resident in shared memory (loaded once); otherwise its 128 x 64 tiles travel through the
// ring next to the X tiles (operators too large to keep: K > 512 at two row blocks), re-read from L2 per work item.
template <int BN, int STAGES, bool RES>
__global__ void __launch_bounds__(kGemmThreads, 1)
bgemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_d, const __grid_constant__ BgemmTcParams p)
{
    constexpr int BK = kGemmBK;
    constexpr uint32_t A_TILE = 128 * BK * 2, X_STAGE = BN * BK * 2, SLAB_BYTES = 32 * 128;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2], a_bar;
    __shared__ uint32_t tmem_base_slot;

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* sA = smem_raw + (((raw + 1023u) & ~1023u) - raw);           // [mb][kb] tiles of 128 x 64, K-major SW128
    const uint32_t a_stage = RES ? 0u : (uint32_t)p.mb * A_TILE;                // operator tiles inside a ring stage
    const uint32_t stage_bytes = a_stage + X_STAGE;
    unsigned char* sX = sA + (RES ? (size_t)p.mb * p.kb * A_TILE : 0);          // [STAGES]([mb] A tiles,) [BN/64] boxes of 64 k x 64 n
    unsigned char* slabs0 = sX + (size_t)STAGES * stage_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_row0 = blockIdx.y * p.mb * 128;
    const uint32_t tmem_cols = 2u * (uint32_t)p.mb * BN;                        // 128 .. 512: a power of two

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_d);
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
        mbar_init(&a_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    // decode of the k-th 64-row block of X / a 32-row group of D into split coordinates
    const int kbox_i = min(64, p.k_inner);

    if (warp == 0) {
        if (lane == 0) {
            if (RES) {          // resident operator
                mbar_arrive_expect_tx(&a_bar, (uint32_t)(p.mb * p.kb) * A_TILE);
                for (int mb = 0; mb < p.mb; ++mb)
                    for (int kb = 0; kb < p.kb; ++kb)
                        tma_load_2d(sA + (size_t)(mb * p.kb + kb) * A_TILE, &map_a, kb * BK, m_row0 + mb * 128, &a_bar);
            }
            uint32_t it = 0;
            for (long long w = blockIdx.x; w < p.items; w += gridDim.x) {
                const long long bt = w / p.n_tiles;
                const int n0 = (int)(w - bt * p.n_tiles) * BN;
                const int bo = (int)(bt / p.b_inner), bi = (int)(bt - (long long)bo * p.b_inner);
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, use = it / STAGES;
                    if (it >= STAGES) mbar_wait(&empty_bar[s], (use - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                    unsigned char* st = sX + (size_t)s * stage_bytes;
                    if (!RES)
                        for (int mb = 0; mb < p.mb; ++mb)
                            tma_load_2d(st + (size_t)mb * A_TILE, &map_a, kb * BK, m_row0 + mb * 128, &full_bar[s]);
                    const int k0 = kb * BK;
                    const int ko = k0 / p.k_inner, ki = k0 - ko * p.k_inner;
#pragma unroll
                    for (int nb = 0; nb < BN / 64; ++nb)
                        tma_load_5d(st + a_stage + nb * (BK * 128), &map_x, n0 + 64 * nb, ki, ko, bi, bo, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_kmn(128, BN);
            if (RES) mbar_wait(&a_bar, 0);
            tc_fence_after();
            uint32_t it = 0, tcount = 0;
            for (long long w = blockIdx.x; w < p.items; w += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
                if (tcount >= 2) mbar_wait(&tmem_empty_bar[acc], (acc_use - 1) & 1);
                tc_fence_after();
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    unsigned char* st = sX + (size_t)s * stage_bytes;
                    const uint64_t dx = umma_smem_desc_mn_sw128(smem_u32(st + a_stage), BK * 128);
                    for (int mb = 0; mb < p.mb; ++mb) {
                        const uint64_t da = umma_smem_desc_sw128(smem_u32(RES ? sA + (size_t)(mb * p.kb + kb) * A_TILE
                                                                              : st + (size_t)mb * A_TILE));
                        const uint32_t tmem_d = tmem_base + (acc * (uint32_t)p.mb + (uint32_t)mb) * BN;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16(tmem_d, da + (uint64_t)(2 * k), dx + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        uint32_t tcount = 0, chunk_count = 0;
        unsigned char* slabs = slabs0 + (size_t)(warp - 2) * 2 * SLAB_BYTES;
        const int mbox_i = min(32, p.m_inner);
        for (long long w = blockIdx.x; w < p.items; w += gridDim.x, ++tcount) {
            const long long bt = w / p.n_tiles;
            const int n0 = (int)(w - bt * p.n_tiles) * BN;
            const int bo = (int)(bt / p.b_inner), bi = (int)(bt - (long long)bo * p.b_inner);
            const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
            mbar_wait(&tmem_full_bar[acc], acc_use & 1);
            tc_fence_after();
            for (int mb = 0; mb < p.mb; ++mb) {
                const int m0 = m_row0 + mb * 128 + q * 32;                      // first of this warp's 32 rows
                const int mo = m0 / p.m_inner, mi = m0 - mo * p.m_inner;
#pragma unroll 1
                for (int c = 0; c < BN; c += 64) {
                    uint32_t r0[32], r1[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * (uint32_t)p.mb + (uint32_t)mb) * BN + (uint32_t)c;
                    tmem_ld_32x32(taddr, r0);
                    tmem_ld_32x32(taddr + 32, r1);
                    tmem_ld_wait();
                    unsigned char* slab = slabs + (size_t)(chunk_count++ & 1) * SLAB_BYTES;
                    if (lane == 0) bulk_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int cc = 8 * j + e;
                            v[e] = __uint_as_float(cc < 32 ? r0[cc] : r1[cc - 32]);
                        }
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]),
                                       t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        *reinterpret_cast<uint4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = pk;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && n0 + c < p.N && m0 < p.M) {
                        tma_store_5d(&map_d, slab, n0 + c, mi, mo, bi, bo);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (lane == 0) bulk_wait_all();
        (void)mbox_i;
    }
    (void)kbox_i;
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
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

// 2-D bf16 row-major matrix [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows rows, SWIZZLE_128B
static int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows)
{
    // box = 64 columns (128 B, the swizzle span) x box_rows rows
    EncodeTiledFn fn = get_encode_fn();
    SEI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
    return 0;
}

// MN-major operand [K rows, MN cols] row-major with leading dimension ld: box = 64 MN-elements x BK k-rows
static int make_map_bf16_mn(CUtensorMap* map, const void* base, long long k_rows, long long mn_cols, long long ld)
{
    EncodeTiledFn fn = get_encode_fn();
    SEI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)mn_cols, (cuuint64_t)k_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)kGemmBK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (MN-major) failed with CUresult %d", (int)r);
    return 0;
}

template <int BN, int STAGES>
static int launch_gemm_mn(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int sm_count, cudaStream_t st)
{
    constexpr size_t smem = (size_t)STAGES * (kGemmBM + BN) * kGemmBK * 2 + 1024 + 4 * 2 * 32 * 128;
    const long long tiles = (long long)((p.M + kGemmBM - 1) / kGemmBM) * ((p.N + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(tiles * p.splits, sm_count);
    SEI_CUDA(allow_smem(gemm_bf16_tn_kernel<BN, STAGES, true, true>, smem));
    gemm_bf16_tn_kernel<BN, STAGES, true, true><<<grid, kGemmThreads, smem, st>>>(ma, mb, ma, ma, p);
    return finish_launch("gemm_bf16_mn_kernel");
}

// shared memory of gemm_bf16_tn_kernel: operand ring + alignment slack + (bf16 TMA-store path only) two 4 KB slabs per
// epilogue warp
template <int BN, int STAGES, int EW>
constexpr size_t gemm_smem_bytes(bool out_f32)
{
    return (size_t)STAGES * (kGemmBM + BN) * kGemmBK * 2 + 1024 + ((out_f32 || BN < 64) ? 0 : (size_t)8 * 32 * 128);
}

template <int BN, int STAGES, int EW>
static int launch_gemm_ew(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const CUtensorMap& md2,
                          const GemmParams& p, bool out_f32, int sm_count, cudaStream_t st)
{
    const long long tiles = (long long)((p.M + kGemmBM - 1) / kGemmBM) * ((p.N + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(tiles * p.splits, sm_count);
    constexpr int threads = 64 + 32 * EW;
    if (out_f32) {
        constexpr size_t smem = gemm_smem_bytes<BN, STAGES, EW>(true);
        SEI_CUDA(allow_smem(gemm_bf16_tn_kernel<BN, STAGES, true, false, EW>, smem));
        gemm_bf16_tn_kernel<BN, STAGES, true, false, EW><<<grid, threads, smem, st>>>(ma, mb, md, md2, p);
    } else {
        constexpr size_t smem = gemm_smem_bytes<BN, STAGES, EW>(false);
        static_assert(smem + 4096 <= 227 * 1024, "operand ring + epilogue slabs + static barriers / bias strips exceed the shared memory of an SM");
        SEI_CUDA(allow_smem(gemm_bf16_tn_kernel<BN, STAGES, false, false, EW>, smem));
        gemm_bf16_tn_kernel<BN, STAGES, false, false, EW><<<grid, threads, smem, st>>>(ma, mb, md, md2, p);
    }
    return finish_launch("gemm_bf16_tn_kernel");
}

// n-tiles per group of the CTA-pair kernels' tile order (SEI_GEMM_RASTER_GN: tuning override; a huge value = m-fastest)
static int gemm_raster_gn()
{
    const char* e = getenv("SEI_GEMM_RASTER_GN");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : kRasterGN;
}

// Short-K, wide-output products (the expanding pointwise convolutions of the two shallowest levels: K <= 128): two
// CTAs per SM with a two-stage ring and four epilogue warps each -- two producers, two MMA issuers and two independent
// accumulator pipelines per SM instead of one (SEI_GEMM_SHORTK=0: the one-CTA shape)
static int launch_gemm_short_k(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const CUtensorMap& md2,
                               const GemmParams& p, int sm_count, cudaStream_t st)
{
    constexpr int BN = 128, STAGES = 2, EW = 4;
    constexpr size_t smem = (size_t)STAGES * (kGemmBM + BN) * kGemmBK * 2 + 1024 + (size_t)EW * 2 * 32 * 128;
    const long long tiles = (long long)((p.M + kGemmBM - 1) / kGemmBM) * ((p.N + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(tiles, 2ll * sm_count);
    SEI_CUDA(allow_smem(gemm_bf16_tn_kernel<BN, STAGES, false, false, EW, 2>, smem));
    gemm_bf16_tn_kernel<BN, STAGES, false, false, EW, 2><<<grid, 64 + 32 * EW, smem, st>>>(ma, mb, md, md2, p);
    return finish_launch("gemm_bf16_tn_kernel");
}

// epilogue warps per CTA (SEI_GEMM_EW=4 restores the round-1 shape for A/B measurements)
static int gemm_epilogue_warps()
{
    const char* e = getenv("SEI_GEMM_EW");
    return (e && *e == '4') ? 4 : 8;
}

// 5-D bf16 tensor map {n, ri, ro, bi, bo}: element (n, r = ro * r_inner + ri, batch = bo * b_inner + bi) of a tensor whose
// rows / batch entries are split with two strides each; box = 64 n x box_ri x box_ro x 1 x 1, SWIZZLE_128B.
static int make_map_bf16_5d(CUtensorMap* map, const void* base, long long n, int r_inner, long long r_outer, int b_inner,
                            long long b_outer, long long s_ri, long long s_ro, long long s_bi, long long s_bo,
                            int box_ri, int box_ro)
{
    EncodeTiledFn fn = get_encode_fn();
    SEI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[5] = {(cuuint64_t)n, (cuuint64_t)r_inner, (cuuint64_t)r_outer, (cuuint64_t)b_inner, (cuuint64_t)b_outer};
    // unused dimensions (size 1) still need a valid (non-zero, 16-byte multiple) stride
    auto st = [](long long v) { return (cuuint64_t)(v > 0 ? v : 8) * 2; };
    cuuint64_t strides[4] = {st(s_ri), st(s_ro), st(s_bi), st(s_bo)};
    cuuint32_t box[5] = {64u, (cuuint32_t)box_ri, (cuuint32_t)box_ro, 1u, 1u};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (5-D) failed with CUresult %d", (int)r);
    return 0;
}

static bool pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

// Returns 1 if the tcgen05 kernel does not take this shape (the caller then uses the mma.sync kernel), 0 on a
// successful launch, or an error code.
int bgemm_tc_try_launch(const void* A, const void* X, void* D, int M, int K, int N, int Kpad, int rows_a,
                        long long batches, int b_inner, long long x_bo, long long x_bi,
                        int k_inner, long long x_ko, long long x_ki,
                        long long d_bo, long long d_bi, int m_inner, long long d_mo, long long d_mi, cudaStream_t st)
{
    const char* off = getenv("SEI_BGEMM_NO_TC");
    if (off && *off == '1') return 1;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    if (dp.cc_major != 10) return 1;
    // shape rules: see the kernel comment
    const int mb_total = (M + 127) / 128;
    int mb = std::min(2, mb_total);                                         // 128-row blocks per CTA
    const int kb = Kpad / 64;
    // (rows of A beyond rows_a are out of the tensor map's bounds: TMA zero-fills them)
    bool resident = true;
    if ((long long)mb * kb * 16384 > 128 * 1024) mb = 1;                    // one row block resident (K <= 512)
    if ((long long)mb * kb * 16384 > 128 * 1024) {                          // stream the operator through the ring
        resident = false;
        mb = std::min(2, mb_total);
    }
    if (N % 8 != 0 || K % 16 != 0 || batches % b_inner != 0) return 1;
    if (!(k_inner % 64 == 0 || (pow2(k_inner) && k_inner < 64)) || K % k_inner != 0) return 1;
    if (!(m_inner % 32 == 0 || (pow2(m_inner) && m_inner < 32)) || M % m_inner != 0) return 1;
    if (M % 32 != 0 && M > 32) return 1;
    if (x_ki <= 0 || d_mi <= 0) return 1;
    const int bn = N <= 64 ? 64 : 128;
    CUtensorMap ma, mx, md;
    rc = make_map_bf16(&ma, A, rows_a, Kpad, Kpad, 128);
    if (rc) return rc;
    const int kbox_i = std::min(64, k_inner), mbox_i = std::min(32, m_inner);
    rc = make_map_bf16_5d(&mx, X, N, k_inner, K / k_inner, b_inner, batches / b_inner, x_ki, x_ko, x_bi, x_bo, kbox_i, 64 / kbox_i);
    if (rc) return rc;
    rc = make_map_bf16_5d(&md, D, N, m_inner, M / m_inner, b_inner, batches / b_inner, d_mi, d_mo, d_bi, d_bo, mbox_i, 32 / mbox_i);
    if (rc) return rc;
    BgemmTcParams p;
    p.M = M; p.K = K; p.N = N; p.mb = mb; p.kb = kb;
    p.n_tiles = (N + bn - 1) / bn; p.b_inner = b_inner; p.k_inner = k_inner; p.m_inner = m_inner;
    p.items = batches * p.n_tiles;
    const int m_blocks = (mb_total + mb - 1) / mb;
    const unsigned gx = (unsigned)std::min<long long>(p.items, std::max(1, dp.sm_count / m_blocks));
    constexpr int ST = 3;
    const size_t slabs = 4 * 2 * 32 * 128 + 1024;
    const dim3 grid(gx, m_blocks);
    if (resident) {
        const size_t smem = (size_t)mb * kb * 16384 + (size_t)ST * bn * 64 * 2 + slabs;
        // narrow tiles (N <= 64: the 32-channel level) move 8 KB per stage: a deeper ring keeps more bytes in flight
        constexpr int ST_DEEP = 6;
        const size_t smem_deep = (size_t)mb * kb * 16384 + (size_t)ST_DEEP * 64 * 64 * 2 + slabs;
        const char* nodeep = getenv("SEI_BGEMM_NO_DEEP_RING");
        if (bn == 64 && smem_deep + 2048 <= (size_t)dp.smem_optin && !(nodeep && *nodeep == '1')) {
            SEI_CUDA(allow_smem(bgemm_tc_kernel<64, ST_DEEP, true>, smem_deep));
            bgemm_tc_kernel<64, ST_DEEP, true><<<grid, kGemmThreads, smem_deep, st>>>(ma, mx, md, p);
        } else if (bn == 64) {
            SEI_CUDA(allow_smem(bgemm_tc_kernel<64, ST, true>, smem));
            bgemm_tc_kernel<64, ST, true><<<grid, kGemmThreads, smem, st>>>(ma, mx, md, p);
        } else {
            SEI_CUDA(allow_smem(bgemm_tc_kernel<128, ST, true>, smem));
            bgemm_tc_kernel<128, ST, true><<<grid, kGemmThreads, smem, st>>>(ma, mx, md, p);
        }
    } else {
        const size_t smem = (size_t)ST * ((size_t)mb * 16384 + (size_t)bn * 64 * 2) + slabs;
        if (bn == 64) {
            SEI_CUDA(allow_smem(bgemm_tc_kernel<64, ST, false>, smem));
            bgemm_tc_kernel<64, ST, false><<<grid, kGemmThreads, smem, st>>>(ma, mx, md, p);
        } else {
            SEI_CUDA(allow_smem(bgemm_tc_kernel<128, ST, false>, smem));
            bgemm_tc_kernel<128, ST, false><<<grid, kGemmThreads, smem, st>>>(ma, mx, md, p);
        }
    }
    return finish_launch("bgemm_tc_kernel");
}

}  // namespace sei

using namespace sei;

// Split-K factor for `tiles` output tiles walked by `units` persistent CTAs (or CTA pairs) with a static stride: the
// launch takes waves x (k-blocks per item + pipeline fill/drain), waves = ceil(items / units).  Picks the factor that
// minimises it (smallest on ties: fewer atomic passes over D), e.g. 16 tiles on 74 pairs -> 9 slices (144 items, two
// full waves) rather than 10 (160 items, a third wave at 16 % occupancy).
static void choose_splits(long long tiles, int units, int nk, int* splits, int* kb_per_split)
{
    long long best_cost = -1;
    const int smax = std::max(1, std::min(512, nk / 4));
    for (int sp = 1; sp <= smax; ++sp) {
        const int kbps = (nk + sp - 1) / sp;
        const int real = (nk + kbps - 1) / kbps;
        const long long waves = (tiles * real + units - 1) / units;
        const long long cost = waves * (kbps + 4);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            *splits = real;
            *kb_per_split = kbps;
        }
    }
}

struct GemmEpilogueExtra {
    const void* gelu_h = nullptr;      // multiply by gelu'(gelu_h)
    long long ld_h = 0;
    const void* res = nullptr;         // add res_scale * res (or multiply by res: res_mul)
    long long ld_r = 0;
    float res_scale = 1.0f;
    int res_mul = 0;
    void* d2 = nullptr;                // gelu_dual: second output gelu'(.) (same shape and pitch as D)
    const float* row_scale = nullptr;  // per-row factor of the bias (period row_period)
    int row_period = 1;
    bool b_kn = false;                 // B is [K, N] row-major (read in place as an MN-major operand): D = A B
};

static int gemm_bf16_tn_impl(const void* A, const void* B, void* D, const float* bias, long long M, int N, int K,
                            long long lda, long long ldb, long long ldd, int out_f32, int tile_n,
                            const GemmEpilogueExtra& ex, void* stream);

extern "C" int sei_gemm_bf16_tn(const void* A, const void* B, void* D, const float* bias, long long M, int N, int K,
                                long long lda, long long ldb, long long ldd, int out_f32, int tile_n, void* stream)
{
    return gemm_bf16_tn_impl(A, B, D, bias, M, N, K, lda, ldb, ldd, out_f32, tile_n, GemmEpilogueExtra(), stream);
}

// D (bf16) = A B^T + bias + res_scale * R: a pointwise convolution whose output is added to a tensor of the same shape in
// the GEMM epilogue (ConvBlock's `x + x1`, reference src/models/convolutional.py:51; UNet's skip / inner-residual
// additions :206-215) instead of a separate pass over both tensors.  R: bf16 [M, N] with row pitch ld_r (multiple of 8).
extern "C" int sei_gemm_bf16_tn_residual(const void* A, const void* B, void* D, const float* bias, const void* R,
                                         float res_scale, long long M, int N, int K, long long lda, long long ldb,
                                         long long ldd, long long ld_r, void* stream)
{
    SEI_REQUIRE(R != nullptr && aligned16(R) && ld_r % 8 == 0 && ld_r >= N, "R must be 16-byte aligned with a row pitch >= N that is a multiple of 8");
    SEI_REQUIRE(N % 8 == 0, "the fused residual needs N %% 8 == 0 (N=%d)", N);
    GemmEpilogueExtra ex;
    ex.res = R; ex.ld_r = ld_r; ex.res_scale = res_scale;
    return gemm_bf16_tn_impl(A, B, D, bias, M, N, K, lda, ldb, ldd, 0, 0, ex, stream);
}

// D (bf16) = A B^T + bias[n] * row_scale[m % period]: the pointwise convolution of Downsample (reference
// src/models/convolutional.py:136-150) when the ideal resampler is applied BEFORE it -- the constant image bias[n]
// becomes bias[n] * R(1)[pixel]; added in the GEMM epilogue instead of an in-place pass over the output.
extern "C" int sei_gemm_bf16_tn_rowscaled_bias(const void* A, const void* B, void* D, const float* bias, const float* row_scale,
                                               int period, long long M, int N, int K, long long lda, long long ldb,
                                               long long ldd, void* stream)
{
    SEI_REQUIRE(bias && row_scale && period > 0, "bias, row_scale and a positive period are required");
    GemmEpilogueExtra ex;
    ex.row_scale = row_scale; ex.row_period = period;
    return gemm_bf16_tn_impl(A, B, D, bias, M, N, K, lda, ldb, ldd, 0, 0, ex, stream);
}

// D (bf16) = A Bkn (* Mult): B given as [K, N] row-major and read in place (MN-major UMMA operand) by the CTA-pair kernel:
// the input gradient of a pointwise convolution, gx = gy W with W = the (C_out x C_in) weight itself -- no transposed
// copy of the weights.  Mult (optional, bf16 [M, N], row pitch ld_m): element-wise multiplier in the epilogue (the stored
// gelu').  Needs M >= 256, N >= 256.
extern "C" int sei_gemm_bf16_nn(const void* A, const void* Bkn, const void* Mult, void* D, long long M, int N, int K,
                                long long lda, long long ldb, long long ldd, long long ld_m, void* stream)
{
    GemmEpilogueExtra ex;
    ex.b_kn = true;
    if (Mult) {
        SEI_REQUIRE(aligned16(Mult) && ld_m % 8 == 0 && ld_m >= N, "Mult must be 16-byte aligned with a row pitch >= N that is a multiple of 8");
        ex.res = Mult; ex.ld_r = ld_m; ex.res_mul = 1;
    }
    return gemm_bf16_tn_impl(A, Bkn, D, nullptr, M, N, K, lda, ldb, ldd, 0, 256, ex, stream);
}

// Aout (bf16) = gelu(A B^T + bias), Dout (bf16) = gelu'(A B^T + bias): ConvBlock.conv2 and ConvBlock.gelu (reference
// src/models/convolutional.py:40-41) as one kernel -- the pre-activation never reaches memory; the stored derivative
// turns the GELU backward into the multiplier epilogue below (one erf evaluation per element and step instead of two).
extern "C" int sei_gemm_bf16_tn_gelu_dual(const void* A, const void* B, const float* bias, void* Aout, void* Dout,
                                          long long M, int N, int K, long long lda, long long ldb, long long ldo, void* stream)
{
    SEI_REQUIRE(Aout && Dout && aligned16(Dout), "null / misaligned output");
    SEI_REQUIRE(N > 32 && ldo % 8 == 0, "the fused GELU needs N > 32 and an output pitch that is a multiple of 8 (N=%d)", N);
    GemmEpilogueExtra ex;
    ex.d2 = Dout;
    return gemm_bf16_tn_impl(A, B, Aout, bias, M, N, K, lda, ldb, ldo, 0, 0, ex, stream);
}

// D (bf16) = (A B^T) * Mult, element-wise in the epilogue: the input gradient of ConvBlock.conv3 multiplied by the stored
// gelu' (the GELU backward without a pass of its own).  Mult: bf16 [M, N] with row pitch ld_m (multiple of 8).
extern "C" int sei_gemm_bf16_tn_mul(const void* A, const void* B, const void* Mult, void* D, long long M, int N, int K,
                                    long long lda, long long ldb, long long ldd, long long ld_m, void* stream)
{
    SEI_REQUIRE(Mult != nullptr && aligned16(Mult) && ld_m % 8 == 0 && ld_m >= N, "Mult must be 16-byte aligned with a row pitch >= N that is a multiple of 8");
    SEI_REQUIRE(N > 32 && N % 8 == 0 && ldd % 8 == 0, "the multiplier epilogue needs N > 32, N %% 8 == 0 and ldd %% 8 == 0 (N=%d)", N);
    GemmEpilogueExtra ex;
    ex.res = Mult; ex.ld_r = ld_m; ex.res_mul = 1;
    return gemm_bf16_tn_impl(A, B, D, nullptr, M, N, K, lda, ldb, ldd, 0, 0, ex, stream);
}

// D (bf16) = (A B^T) * gelu'(H): the input gradient of `conv3` with the GELU backward applied in the epilogue.
// H: bf16 [M, N] with row pitch ld_h (multiple of 8); needs N % 64 == 0 and ldd % 8 == 0 (TMA-store epilogue).
extern "C" int sei_gemm_bf16_tn_gelu_bwd(const void* A, const void* B, void* D, const void* H, long long M, int N, int K,
                                         long long lda, long long ldb, long long ldd, long long ld_h, void* stream)
{
    SEI_REQUIRE(H != nullptr && aligned16(H) && ld_h % 8 == 0 && ld_h >= N, "H must be 16-byte aligned with a row pitch >= N that is a multiple of 8");
    SEI_REQUIRE(N % 64 == 0 && ldd % 8 == 0, "the fused GELU backward needs N %% 64 == 0 and ldd %% 8 == 0 (N=%d)", N);
    GemmEpilogueExtra ex;
    ex.gelu_h = H; ex.ld_h = ld_h;
    return gemm_bf16_tn_impl(A, B, D, nullptr, M, N, K, lda, ldb, ldd, 0, 0, ex, stream);
}

static int gemm_bf16_tn_impl(const void* A, const void* B, void* D, const float* bias, long long M, int N, int K,
                            long long lda, long long ldb, long long ldd, int out_f32, int tile_n,
                            const GemmEpilogueExtra& ex, void* stream)
{
    const void* gelu_h = ex.gelu_h;
    const long long ld_h = ex.ld_h;
    SEI_REQUIRE(A && B && D, "null pointer argument");
    SEI_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "bad shape M=%lld N=%d K=%d", M, N, K);
    SEI_REQUIRE(lda >= K && ldb >= (ex.b_kn ? N : K) && ldd >= N, "leading dimensions smaller than the rows");
    SEI_REQUIRE(!ex.b_kn || (!out_f32 && M >= 256 && N >= 256 && N % 8 == 0 && ldd % 8 == 0),
                "the [K, N] form of B is taken by the CTA-pair kernel only (bf16 output, M >= 256, N >= 256, N %% 8 == 0)");
    SEI_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "lda/ldb must be multiples of 8 bf16 (16-byte TMA row pitch); pad K");
    SEI_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D), "A, B, D must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    SEI_REQUIRE(dp.cc_major == 10, "tcgen05 GEMM needs an sm_100 device");
    int bn = tile_n;
    if (bn == 0) bn = N >= 256 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
    SEI_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "tile_n must be 0, 32, 64, 128 or 256");
    SEI_REQUIRE((M + kGemmBM - 1) / kGemmBM <= 2147483647ll && (N + bn - 1) / bn <= 65535, "grid too large");
    CUtensorMap ma, mb;
    rc = make_map_bf16(&ma, A, M, K, lda, kGemmBM);
    if (rc) return rc;
    if (!ex.b_kn) {
        rc = make_map_bf16(&mb, B, N, K, ldb, bn);
        if (rc) return rc;
    }
    GemmParams p;
    p.D = D; p.bias = bias; p.M = (int)M; p.N = N; p.K = K; p.ldd = (int)ldd; p.accumulate = 0;
    p.gelu_h = static_cast<const __nv_bfloat16*>(gelu_h); p.ld_h = (int)ld_h;
    p.res = static_cast<const __nv_bfloat16*>(ex.res); p.ld_r = (int)ex.ld_r; p.res_scale = ex.res_scale;
    p.res_mul = ex.res_mul; p.gelu_dual = ex.d2 ? 1 : 0; p.raster_gn = gemm_raster_gn();
    p.row_scale = ex.row_scale; p.row_period = ex.row_period;
    SEI_REQUIRE(!ex.res || !out_f32, "the fused residual is a bf16-output epilogue");
    // bf16 output through shared memory + bulk tensor stores when the row pitch allows a tensor map
    const char* nts = getenv("SEI_GEMM_NO_TMA_STORE");
    p.tma_store = (!out_f32 && bn >= 64 && ldd % 8 == 0 && !(nts && *nts == '1')) ? 1 : 0;
    SEI_REQUIRE(!gelu_h || p.tma_store, "the fused GELU backward needs the TMA-store epilogue (N >= 64, ldd %% 8 == 0)");
    SEI_REQUIRE(!(ex.d2 || ex.res_mul) || p.tma_store, "the fused GELU / multiplier epilogues need the TMA-store path (N > 32, ldd %% 8 == 0)");
    CUtensorMap md = ma, md2 = ma;
    if (p.tma_store) {
        rc = make_map_bf16(&md, D, M, N, ldd, 32);
        if (rc) return rc;
        if (ex.d2) {
            rc = make_map_bf16(&md2, ex.d2, M, N, ldd, 32);
            if (rc) return rc;
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // split-K for the weight-gradient shapes (few output tiles, very long K = pixels): fp32 output only
    const int nk = (K + kGemmBK - 1) / kGemmBK;
    const long long tiles = ((M + kGemmBM - 1) / kGemmBM) * (long long)((N + bn - 1) / bn);
    p.splits = 1;
    p.kb_per_split = nk;
    if (out_f32 && tiles < dp.sm_count && nk >= 8 && ldd == N) {
        choose_splits(tiles, dp.sm_count, nk, &p.splits, &p.kb_per_split);
        if (p.splits > 1) SEI_CUDA(cudaMemsetAsync(D, 0, (size_t)M * N * sizeof(float), st));
    }
    // CTA-pair kernel for the large bf16-output TN products (forward / input gradient of the deep layers)
    const char* no2 = getenv("SEI_GEMM_NO_2CTA");
    if (bn == 256 && p.tma_store && p.splits == 1 && M >= 256 && N >= 256 && dp.sm_count >= 2 && !(no2 && *no2 == '1')) {
        constexpr int ST2 = 5;
        CUtensorMap mb2;
        rc = ex.b_kn ? make_map_bf16_mn(&mb2, B, K, N, ldb) : make_map_bf16(&mb2, B, N, K, ldb, 128);
        if (rc) return rc;
        const long long ctiles = ((M + 255) / 256) * (long long)((N + 255) / 256);
        const unsigned grid = 2u * (unsigned)std::min<long long>(ctiles, dp.sm_count / 2);
        if (ex.b_kn) {
            constexpr size_t smem2 = (size_t)ST2 * (kGemmBM + 128) * kGemmBK * 2 + 1024 + 8 * 32 * 128;
            SEI_CUDA(allow_smem(gemm_bf16_tn_2cta_kernel<ST2, 8, true>, smem2));
            gemm_bf16_tn_2cta_kernel<ST2, 8, true><<<grid, 64 + 32 * 8, smem2, st>>>(ma, mb2, md, md2, p);
        } else if (gemm_epilogue_warps() == 8) {
            constexpr size_t smem2 = (size_t)ST2 * (kGemmBM + 128) * kGemmBK * 2 + 1024 + 8 * 32 * 128;
            static_assert(smem2 + 4096 <= 227 * 1024, "CTA-pair kernel: ring + slabs + barriers exceed the shared memory of an SM");
            SEI_CUDA(allow_smem(gemm_bf16_tn_2cta_kernel<ST2, 8>, smem2));
            gemm_bf16_tn_2cta_kernel<ST2, 8><<<grid, 64 + 32 * 8, smem2, st>>>(ma, mb2, md, md2, p);
        } else {
            constexpr size_t smem2 = (size_t)ST2 * (kGemmBM + 128) * kGemmBK * 2 + 1024 + 4 * 2 * 32 * 128;
            SEI_CUDA(allow_smem(gemm_bf16_tn_2cta_kernel<ST2, 4>, smem2));
            gemm_bf16_tn_2cta_kernel<ST2, 4><<<grid, 64 + 32 * 4, smem2, st>>>(ma, mb2, md, md2, p);
        }
        return finish_launch("gemm_bf16_tn_2cta_kernel");
    }
    SEI_REQUIRE(!ex.b_kn, "the [K, N] form of B needs the CTA-pair kernel (disabled by SEI_GEMM_NO_2CTA or a split)");
    const char* sk = getenv("SEI_GEMM_SHORTK");
    if (bn == 128 && nk <= 2 && !out_f32 && p.tma_store && !(sk && *sk == '0'))
        return launch_gemm_short_k(ma, mb, md, md2, p, dp.sm_count, st);
    if (gemm_epilogue_warps() == 8) {        // eight epilogue warps; one operand stage fewer where the slabs need the room
        switch (bn) {
        case 32: return launch_gemm_ew<32, 8, 8>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
        case 64: return launch_gemm_ew<64, 7, 8>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
        case 128: return launch_gemm_ew<128, 5, 8>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
        default: return launch_gemm_ew<256, 3, 8>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
        }
    }
    switch (bn) {
    case 32: return launch_gemm_ew<32, 8, 4>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
    case 64: return launch_gemm_ew<64, 7, 4>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
    case 128: return launch_gemm_ew<128, 5, 4>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
    default: return launch_gemm_ew<256, 3, 4>(ma, mb, md, md2, p, out_f32 != 0, dp.sm_count, st);
    }
}


// D[M, N] (fp32) = A[K, M]^T * B[K, N]: both operands MN-major (the contraction index is the row).  The weight
// gradient of a pointwise convolution: A = dL/dy (pixels x C_out), B = x (pixels x C_in), D = dL/dW (C_out x C_in).
static int gemm_bf16_atb_impl(const void* A, const void* B, float* D, long long K, int M, int N,
                             long long lda, long long ldb, int accumulate, void* stream);

extern "C" int sei_gemm_bf16_atb(const void* A, const void* B, float* D, long long K, int M, int N,
                                 long long lda, long long ldb, void* stream)
{
    return gemm_bf16_atb_impl(A, B, D, K, M, N, lda, ldb, 0, stream);
}

// D += A^T B: the weight gradient added straight into the parameter's gradient buffer (fp32 atomic adds from the
// epilogue), instead of a temporary plus a separate accumulation pass over all parameters
extern "C" int sei_gemm_bf16_atb_accumulate(const void* A, const void* B, float* D, long long K, int M, int N,
                                            long long lda, long long ldb, void* stream)
{
    return gemm_bf16_atb_impl(A, B, D, K, M, N, lda, ldb, 1, stream);
}

static int gemm_bf16_atb_impl(const void* A, const void* B, float* D, long long K, int M, int N,
                             long long lda, long long ldb, int accumulate, void* stream)
{
    SEI_REQUIRE(A && B && D, "null pointer argument");
    SEI_REQUIRE(M > 0 && N > 0 && K > 0 && K < (1ll << 31), "bad shape M=%d N=%d K=%lld", M, N, K);
    SEI_REQUIRE(lda >= M && ldb >= N, "leading dimensions smaller than the rows");
    SEI_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "lda/ldb must be multiples of 8 bf16 (16-byte TMA row pitch)");
    SEI_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D), "A, B, D must be 16-byte aligned");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    SEI_REQUIRE(dp.cc_major == 10, "tcgen05 GEMM needs an sm_100 device");
    const int bn = N >= 256 ? 256 : (N > 64 ? 128 : 64);
    CUtensorMap ma, mb;
    rc = make_map_bf16_mn(&ma, A, K, M, lda);
    if (rc) return rc;
    rc = make_map_bf16_mn(&mb, B, K, N, ldb);
    if (rc) return rc;
    GemmParams p;
    p.D = D; p.bias = nullptr; p.M = M; p.N = N; p.K = (int)K; p.ldd = N; p.tma_store = 0; p.accumulate = accumulate;
    p.gelu_h = nullptr; p.ld_h = 0; p.res = nullptr; p.ld_r = 0; p.res_scale = 0.f; p.res_mul = 0; p.gelu_dual = 0; p.raster_gn = gemm_raster_gn(); p.row_scale = nullptr; p.row_period = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nk = (int)((K + kGemmBK - 1) / kGemmBK);
    const long long tiles = (long long)((M + kGemmBM - 1) / kGemmBM) * ((N + bn - 1) / bn);
    // CTA-pair kernel for the wide weight matrices (256 x 256 tiles per cluster)
    const char* no2 = getenv("SEI_GEMM_NO_2CTA_MN");
    if (bn == 256 && M >= 256 && dp.sm_count >= 2 && !(no2 && *no2 == '1')) {
        constexpr int ST2 = 6;
        const long long ctiles = (long long)((M + 255) / 256) * ((N + 255) / 256);
        const int clusters = dp.sm_count / 2;
        p.splits = 1;
        p.kb_per_split = nk;
        if (nk >= 8) choose_splits(ctiles, clusters, nk, &p.splits, &p.kb_per_split);
        if (p.splits > 1 && !accumulate) SEI_CUDA(cudaMemsetAsync(D, 0, (size_t)M * N * sizeof(float), st));
        constexpr size_t smem2 = (size_t)ST2 * (kGemmBM + 128) * kGemmBK * 2 + 1024;
        const unsigned grid = 2u * (unsigned)std::min<long long>(ctiles * p.splits, clusters);
        SEI_CUDA(allow_smem(gemm_bf16_atb_2cta_kernel<ST2>, smem2));
        gemm_bf16_atb_2cta_kernel<ST2><<<grid, kGemmThreads, smem2, st>>>(ma, mb, p);
        return finish_launch("gemm_bf16_mn_2cta_kernel");
    }
    p.splits = 1;
    p.kb_per_split = nk;
    if (nk >= 8) {
        choose_splits(tiles, dp.sm_count, nk, &p.splits, &p.kb_per_split);
        if (p.splits > 1 && !accumulate) SEI_CUDA(cudaMemsetAsync(D, 0, (size_t)M * N * sizeof(float), st));
    }
    switch (bn) {
    case 64: return launch_gemm_mn<64, 8>(ma, mb, p, dp.sm_count, st);
    case 128: return launch_gemm_mn<128, 6>(ma, mb, p, dp.sm_count, st);
    default: return launch_gemm_mn<256, 4>(ma, mb, p, dp.sm_count, st);
    }
}
