// dwconv_tile.cu -- depthwise 7x7 convolution (reference ConvBlock.conv1, src/models/convolutional.py:36-38:
// Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)) and its weight gradient on channels-last bf16 tensors whose
// channel count is a multiple of 64: shared-memory halo tiles staged by ONE tensor-map TMA copy per tile.
//
// 49 fp32 FMAs per element and nothing to contract over: CUDA-core work bounded by the FMA rate (3.3 GFMA per call at
// batch 32 = 92 us), not by HBM.  The register-window kernel of round 1 (cnn_elem.cu: every output row re-reads its
// seven 14-column input windows and 49 tap vectors through L1) ran at a third of that rate; ncu showed the L1 data path
// as the limiter (49 B of L1 traffic per output element).  Here:
//   * a CTA owns a 16 x 16 pixel tile of 64 channels.  Its (16 + 6)^2 x 64 input tile is ONE 4-D bulk tensor copy
//     (c, x, y, b) whose out-of-range coordinates (the 3-pixel zero padding of the convolution, ragged image edges) are
//     zero-filled by the TMA engine: no bounds logic anywhere in the compute loop;
//   * a lane owns a channel PAIR (one 32-bit shared word: a warp reads the 128 contiguous bytes of a pixel, conflict
//     free); a thread owns 2 output rows x 16 output columns of it and keeps the 64 accumulators in registers;
//   * per input row it loads the 22-pixel window once (22 LDS.32) and uses it for both output rows: 448 FMAs per 22 + 14
//     shared loads (taps: 7 LDS.64 per row pair from a [49][64] fp32 table) -- the shared pipe runs at about half of
//     its rate while the FMA pipe is the limiter.
// The input gradient is the same kernel with flipped taps; the block's residual gradient is added in the store.
// The weight gradient uses the same tiles (x with halo, dL/dy without): warp ky of a CTA slides a 7-wide register
// window along each tile row and keeps dW[c][ky][0..6] of its channel pair in registers across all tiles of a
// persistent CTA; the eighth warp sums the bias gradient.
#include "sei_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <mutex>

namespace sei {

constexpr int kDtTW = 16;                 // output tile width
constexpr int kDtThreads = 256;

// G = 1: 64 channels per CTA, a warp instruction reads one pixel (32 lanes x one channel pair).
// G = 2: 32 channels per CTA (the network's first level); lanes 0-15 and 16-31 work on NEIGHBOURING ROWS, and the tile
//        pitches (23 and 17 pixels of 16 words) are odd, so the two half-warps read complementary halves of the banks.
template <int G> struct DwTile {
    static constexpr int CB = 64 / G, CBW = CB / 2;          // channels / 32-bit words per pixel
    static constexpr int TH = 16 * G;                        // output rows per CTA (8 warps x 2 rows x G)
    static constexpr int IW = kDtTW + 6 + (G - 1), IH = TH + 6;
    static constexpr int RW = kDtTW + (G - 1);               // pitch of the halo-free tiles (dL/dy, residual)
    static constexpr uint32_t XBYTES = IH * IW * CB * 2, WBYTES = 49 * CB * 4, RBYTES = TH * RW * CB * 2;
    static_assert(XBYTES % 128 == 0 && WBYTES % 128 == 0 && RBYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

struct DwTileParams {
    const float* wt;              // [49][C] fp32 taps
    const float* bias;            // [C] or null
    const __nv_bfloat16* res;     // [B, H, W, C] or null
    float res_scale;
    __nv_bfloat16* y;
    float* partial;               // wgrad: [slots][C * 50]
    int B, H, W, C, tiles_x, tiles_y, cblocks;
    long long tiles;              // B * tiles_y * tiles_x (per channel block)
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
                   "r"(c3)
                 : "memory");
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t w)
{
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

template <int G>
__global__ void __launch_bounds__(kDtThreads, 2) dwconv7_tile_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                      const __grid_constant__ CUtensorMap map_r,
                                                                      const __grid_constant__ DwTileParams p)
{
    using T = DwTile<G>;
    constexpr int CB = T::CB, CBW = T::CBW, IW = T::IW, Q = kDtTW, R = 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t* sx = reinterpret_cast<const uint32_t*>(base);                                  // [IH][IW][CBW]
    float* sw = reinterpret_cast<float*>(base + T::XBYTES);                                         // [49][CB]
    const uint32_t* sr = reinterpret_cast<const uint32_t*>(base + T::XBYTES + T::WBYTES);           // [TH][RW][CBW]

    long long t = blockIdx.x;
    const int cb = (int)(t % p.cblocks); t /= p.cblocks;
    const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
    const int ty = (int)(t % p.tiles_y);
    const int b = (int)(t / p.tiles_y);
    const int c0 = cb * CB, x0 = tx * kDtTW, y0 = ty * T::TH;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&bar, T::XBYTES + (p.res ? T::RBYTES : 0u));
        tma_load_4d(base, &map_x, c0, x0 - 3, y0 - 3, b, &bar);
        // the residual tile travels with the input tile: reading it from global memory in the store loop exposed one
        // memory latency per output pixel of a thread (362 us against 186 us without the residual)
        if (p.res) tma_load_4d(base + T::XBYTES + T::WBYTES, &map_r, c0, x0, y0, b, &bar);
    }
    for (int i = threadIdx.x; i < 49 * CB / 4; i += kDtThreads) {
        const int tap = i / (CB / 4), v = i - tap * (CB / 4);
        reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(p.wt + (size_t)tap * p.C + c0) + v);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / CBW, cp = lane - g * CBW;
    const int r0 = R * G * warp + g;                // tile-local output rows r0 and r0 + G
    float acc[R][Q][2];
    {
        float2 bv = make_float2(0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float2*>(p.bias + c0) + cp);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                acc[r][q][0] = bv.x;
                acc[r][q][1] = bv.y;
            }
    }
    const uint32_t* sxl = sx + cp;                                             // this lane's channel pair
    const float2* swl = reinterpret_cast<const float2*>(sw) + cp;
#pragma unroll 1
    for (int iy = 0; iy < G + 7; ++iy) {
        const uint32_t* row = sxl + (size_t)((r0 + iy) * IW) * CBW;
        float2 win[Q + 6];
#pragma unroll
        for (int j = 0; j < Q + 6; ++j) win[j] = bf2_to_f2(row[j * CBW]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ky = iy - r * G;
            if (ky >= 0 && ky < 7) {                // warp-uniform
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float2 w = swl[(ky * 7 + kx) * CBW];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        acc[r][q][0] = fmaf(w.x, win[q + kx].x, acc[r][q][0]);
                        acc[r][q][1] = fmaf(w.y, win[q + kx].y, acc[r][q][1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int yl = r0 + r * G, y = y0 + yl;
        if (y >= p.H) continue;
        const size_t rowoff = (((size_t)b * p.H + y) * p.W + x0) * p.C + c0;
        uint32_t* orow = reinterpret_cast<uint32_t*>(p.y + rowoff) + cp;
        const uint32_t* rrow = p.res ? sr + (size_t)(yl * T::RW) * CBW + cp : nullptr;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (x0 + q < p.W) {
                float a0 = acc[r][q][0], a1 = acc[r][q][1];
                if (rrow) {
                    const float2 rv = bf2_to_f2(rrow[q * CBW]);
                    a0 = fmaf(p.res_scale, rv.x, a0);
                    a1 = fmaf(p.res_scale, rv.y, a1);
                }
                const __nv_bfloat162 o = __floats2bfloat162_rn(a0, a1);
                orow[(size_t)q * (p.C / 2)] = *reinterpret_cast<const uint32_t*>(&o);
            }
        }
    }
}

// Weight / bias gradient partials.  grid = (slots, C / CB); CTA (slot, cb) walks tiles slot, slot + slots, ... of its
// channel block.  Warp ky < 7: dW[c][ky][0..6] of the lane's channel pair (14 accumulators, kept across tiles); for every
// tile row it holds the 22-pixel window of x[y + ky - 3] and the 16 values of dL/dy[y] in registers: 224 FMAs per 38
// shared loads.  Warp 7: bias gradient.  partial[slot][c * 49 + ky * 7 + kx], then [C * 49 + c].
template <int G>
__global__ void __launch_bounds__(kDtThreads, 2) dwconv7_wgrad_tile_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                            const __grid_constant__ CUtensorMap map_g,
                                                                            const __grid_constant__ DwTileParams p)
{
    using T = DwTile<G>;
    constexpr int CB = T::CB, CBW = T::CBW, IW = T::IW, RW = T::RW, Q = kDtTW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t* sx = reinterpret_cast<const uint32_t*>(base);                     // [IH][IW][CBW] channel pairs
    const uint32_t* sg = reinterpret_cast<const uint32_t*>(base + T::XBYTES);         // [TH][RW][CBW]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / CBW, cp = lane - g * CBW;
    const int cb = blockIdx.y, c0 = cb * CB;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    float acc[7][2];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k][0] = acc[k][1] = 0.f;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int tx = (int)(t % p.tiles_x);
        const long long r = t / p.tiles_x;
        const int ty = (int)(r % p.tiles_y);
        const int b = (int)(r / p.tiles_y);
        if (threadIdx.x == 0) {
            fence_proxy_async();                  // the tiles were read by generic loads in the previous iteration
            mbar_arrive_expect_tx(&bar, T::XBYTES + T::RBYTES);
            tma_load_4d(base, &map_x, c0, tx * kDtTW - 3, ty * T::TH - 3, b, &bar);
            tma_load_4d(base + T::XBYTES, &map_g, c0, tx * kDtTW, ty * T::TH, b, &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        if (warp < 7) {
            const int ky = warp;
#pragma unroll 1
            for (int yy = 0; yy < T::TH / G; ++yy) {
                const int y = G * yy + g;
                const uint32_t* xr = sx + (size_t)((y + ky) * IW) * CBW + cp;
                const uint32_t* gr = sg + (size_t)(y * RW) * CBW + cp;
                float2 win[Q + 6], gv[Q];
#pragma unroll
                for (int j = 0; j < Q + 6; ++j) win[j] = bf2_to_f2(xr[j * CBW]);
#pragma unroll
                for (int q = 0; q < Q; ++q) gv[q] = bf2_to_f2(gr[q * CBW]);
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        acc[kx][0] = fmaf(gv[q].x, win[q + kx].x, acc[kx][0]);
                        acc[kx][1] = fmaf(gv[q].y, win[q + kx].y, acc[kx][1]);
                    }
            }
        } else {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int i = 0; i < T::TH / G * Q; ++i) {
                const int y = G * (i / Q) + g, xq = i % Q;
                const float2 v = bf2_to_f2(sg[(size_t)(y * RW + xq) * CBW + cp]);
                s0 += v.x;
                s1 += v.y;
            }
            acc[0][0] += s0;
            acc[0][1] += s1;
        }
        __syncthreads();                          // every warp is done with the tiles before the next copy lands
    }
    if (G == 2) {                                 // the two half-warps hold the same channel pairs (different rows)
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            acc[k][0] += __shfl_down_sync(0xffffffffu, acc[k][0], 16);
            acc[k][1] += __shfl_down_sync(0xffffffffu, acc[k][1], 16);
        }
        if (g != 0) return;
    }
    float* out = p.partial + (size_t)blockIdx.x * ((size_t)p.C * 50);
    const int c = c0 + 2 * cp;
    if (warp < 7) {
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            out[(size_t)c * 49 + warp * 7 + kx] = acc[kx][0];
            out[(size_t)(c + 1) * 49 + warp * 7 + kx] = acc[kx][1];
        }
    } else {
        out[(size_t)p.C * 49 + c] = acc[0][0];
        out[(size_t)p.C * 49 + c + 1] = acc[0][1];
    }
}

// ---------------------------------------------------------------- weight gradient of the 3x3 edge convolutions
// dW[co][ci][ky][kx] = sum_p g[p][co] * x[p + (ky - 1, kx - 1)][ci],  db[co] = sum_p g[p][co]   (co < 3, ci < 32)
// for UNet.out_conv (reference src/models/convolutional.py:176; g = dL/dy, 4 bf16 per pixel) and, with the roles of the
// image and the gradient exchanged, UNet.in_conv (:175).  864 numbers, each a sum over every pixel: 1.8 GFMA per call.
// Shared-memory tiles like the depthwise kernels above: x tile (32 rows + halo) x (16 + halo, pitch 19) x 32 channels by
// one TMA copy with zero fill, the 8-byte gradient pixels by ordinary loads.  Warp (slice, ky), half-warp g, lane = channel
// pair: it slides a 3-pixel register window along the rows 2 i + g of its slice: per pixel one new x word and one
// broadcast gradient word feed 18 FMAs.  The round-1 kernel (a thread per (ci / 8, tap), every operand through L1) took
// 680 us per call.
constexpr int kCwTW = 16, kCwTH = 32, kCwIW = 19, kCwIH = kCwTH + 2, kCwWarps = 9, kCwThreads = 32 * kCwWarps;
constexpr uint32_t kCwXBytes = kCwIH * kCwIW * 64;
static_assert(kCwXBytes % 128 == 0, "gradient tile must stay aligned");

struct CwParams {
    const uint2* g4;              // [B, H, W] pixels of 4 bf16
    float* partial;               // [3 * gridDim.x][nW + 3]
    int B, H, W, tiles_x, tiles_y;
    long long tiles;
};

__global__ void __launch_bounds__(kCwThreads, 2) conv3_wgrad_tile_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                          const __grid_constant__ CwParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t* sx = reinterpret_cast<const uint32_t*>(base);                 // [IH][IW][16 channel pairs]
    uint2* sg = reinterpret_cast<uint2*>(base + kCwXBytes);                        // [TH][TW]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = warp % 3, slice = warp / 3, g = lane >> 4, cp = lane & 15;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    float acc[3][3][2], gb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int co = 0; co < 3; ++co) acc[kx][co][0] = acc[kx][co][1] = 0.f;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int tx = (int)(t % p.tiles_x);
        const long long r = t / p.tiles_x;
        const int ty = (int)(r % p.tiles_y), b = (int)(r / p.tiles_y);
        const int x0 = tx * kCwTW, y0 = ty * kCwTH;
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar, kCwXBytes);
            tma_load_4d(base, &map_x, 0, x0 - 1, y0 - 1, b, &bar);
        }
        for (int i = threadIdx.x; i < kCwTH * kCwTW; i += kCwThreads) {
            const int yy = y0 + i / kCwTW, xx = x0 + i % kCwTW;
            sg[i] = (yy < p.H && xx < p.W) ? __ldg(p.g4 + ((size_t)b * p.H + yy) * p.W + xx) : make_uint2(0u, 0u);
        }
        __syncthreads();
        mbar_wait(&bar, phase);
        phase ^= 1u;
#pragma unroll 1
        for (int pi = slice; pi < kCwTH / 2; pi += 3) {
            const int row = 2 * pi + g;
            const uint32_t* xr = sx + (size_t)((row + ky) * kCwIW) * 16 + cp;
            const uint2* gr = sg + row * kCwTW;
            float2 w0 = bf2_to_f2(xr[0]), w1 = bf2_to_f2(xr[16]);
#pragma unroll
            for (int c = 0; c < kCwTW; ++c) {
                const float2 w2 = bf2_to_f2(xr[(c + 2) * 16]);
                const uint2 gq = gr[c];
                const float2 g01 = bf2_to_f2(gq.x), g2_ = bf2_to_f2(gq.y);
                const float gv[3] = {g01.x, g01.y, g2_.x};
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    acc[0][co][0] = fmaf(gv[co], w0.x, acc[0][co][0]); acc[0][co][1] = fmaf(gv[co], w0.y, acc[0][co][1]);
                    acc[1][co][0] = fmaf(gv[co], w1.x, acc[1][co][0]); acc[1][co][1] = fmaf(gv[co], w1.y, acc[1][co][1]);
                    acc[2][co][0] = fmaf(gv[co], w2.x, acc[2][co][0]); acc[2][co][1] = fmaf(gv[co], w2.y, acc[2][co][1]);
                }
                if (ky == 0) {
                    gb[0] += gv[0]; gb[1] += gv[1]; gb[2] += gv[2];
                }
                w0 = w1;
                w1 = w2;
            }
        }
        __syncthreads();
    }
    // the two half-warps hold the same channel pairs (rows of the other parity)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            acc[kx][co][0] += __shfl_down_sync(0xffffffffu, acc[kx][co][0], 16);
            acc[kx][co][1] += __shfl_down_sync(0xffffffffu, acc[kx][co][1], 16);
        }
#pragma unroll
    for (int co = 0; co < 3; ++co) gb[co] += __shfl_down_sync(0xffffffffu, gb[co], 16);
    if (g != 0) return;
    constexpr int nW = 3 * 32 * 9;
    float* out = p.partial + ((size_t)blockIdx.x * 3 + slice) * (nW + 3);
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            out[(co * 32 + 2 * cp) * 9 + ky * 3 + kx] = acc[kx][co][0];
            out[(co * 32 + 2 * cp + 1) * 9 + ky * 3 + kx] = acc[kx][co][1];
        }
    if (ky == 0 && cp == 0) {
#pragma unroll
        for (int co = 0; co < 3; ++co) out[nW + co] = gb[co];
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn dw_encode_fn()
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

// [B, H, W, C] bf16 as a 4-D tensor (c, x, y, b); box = cb channels x bw x bh x 1, no swizzle, zero fill outside
static int make_map_bhwc(CUtensorMap* map, const void* base, int B, int H, int W, int C, int cb, int bw, int bh)
{
    EncodeTiledFn fn = dw_encode_fn();
    SEI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (B, H, W, C) failed with CUresult %d", (int)r);
    return 0;
}

bool dwconv7_tile_supported(int C) { return C >= 32 && C % 32 == 0; }

static int dw_tile_g(int C) { return C % 64 == 0 ? 1 : 2; }

int dwconv7_tile_wgrad_slots(int C, int sm_count) { return std::max(1, 2 * sm_count / (C / (64 / dw_tile_g(C)))); }

template <int G>
static int dw_tile_forward(const void* x, const float* wt, const float* bias, const void* res, float res_scale, void* y, int B,
                           int H, int W, int C, cudaStream_t st)
{
    using T = DwTile<G>;
    CUtensorMap mx, mr;
    int rc = make_map_bhwc(&mx, x, B, H, W, C, T::CB, T::IW, T::IH);
    if (rc) return rc;
    mr = mx;
    if (res) {
        rc = make_map_bhwc(&mr, res, B, H, W, C, T::CB, T::RW, T::TH);
        if (rc) return rc;
    }
    DwTileParams p = {};
    p.wt = wt; p.bias = bias; p.res = static_cast<const __nv_bfloat16*>(res); p.res_scale = res_scale;
    p.y = static_cast<__nv_bfloat16*>(y);
    p.B = B; p.H = H; p.W = W; p.C = C;
    p.tiles_x = (W + kDtTW - 1) / kDtTW; p.tiles_y = (H + T::TH - 1) / T::TH; p.cblocks = C / T::CB;
    p.tiles = (long long)B * p.tiles_y * p.tiles_x;
    const long long ctas = p.tiles * p.cblocks;
    SEI_REQUIRE(ctas < (1ll << 31), "depthwise convolution: grid too large");
    const size_t smem = T::XBYTES + T::WBYTES + 128 + (res ? T::RBYTES : 0u);
    SEI_CUDA(allow_smem(dwconv7_tile_kernel<G>, T::XBYTES + T::WBYTES + 128 + T::RBYTES));
    dwconv7_tile_kernel<G><<<(unsigned)ctas, kDtThreads, smem, st>>>(mx, mr, p);
    return finish_launch("dwconv7_tile_kernel");
}

int dwconv7_tile_forward(const void* x, const float* wt, const float* bias, const void* res, float res_scale, void* y, int B,
                         int H, int W, int C, cudaStream_t st)
{
    return dw_tile_g(C) == 1 ? dw_tile_forward<1>(x, wt, bias, res, res_scale, y, B, H, W, C, st)
                             : dw_tile_forward<2>(x, wt, bias, res, res_scale, y, B, H, W, C, st);
}

template <int G>
static int dw_tile_wgrad(const void* gy, const void* x, float* partial, int slots, int B, int H, int W, int C, cudaStream_t st)
{
    using T = DwTile<G>;
    CUtensorMap mx, mg;
    int rc = make_map_bhwc(&mx, x, B, H, W, C, T::CB, T::IW, T::IH);
    if (rc) return rc;
    rc = make_map_bhwc(&mg, gy, B, H, W, C, T::CB, T::RW, T::TH);
    if (rc) return rc;
    DwTileParams p = {};
    p.partial = partial;
    p.B = B; p.H = H; p.W = W; p.C = C;
    p.tiles_x = (W + kDtTW - 1) / kDtTW; p.tiles_y = (H + T::TH - 1) / T::TH; p.cblocks = C / T::CB;
    p.tiles = (long long)B * p.tiles_y * p.tiles_x;
    constexpr size_t smem = T::XBYTES + T::RBYTES + 128;
    SEI_CUDA(allow_smem(dwconv7_wgrad_tile_kernel<G>, smem));
    dwconv7_wgrad_tile_kernel<G><<<dim3((unsigned)slots, (unsigned)p.cblocks), kDtThreads, smem, st>>>(mx, mg, p);
    return finish_launch("dwconv7_wgrad_tile_kernel");
}

// partial: [slots][C * 50] floats with slots = dwconv7_tile_wgrad_slots(C, sm_count)
int dwconv7_tile_wgrad(const void* gy, const void* x, float* partial, int slots, int B, int H, int W, int C, cudaStream_t st)
{
    return dw_tile_g(C) == 1 ? dw_tile_wgrad<1>(gy, x, partial, slots, B, H, W, C, st)
                             : dw_tile_wgrad<2>(gy, x, partial, slots, B, H, W, C, st);
}

int conv3_wgrad_tile_slots(int sm_count) { return 3 * 2 * sm_count; }

// partial: [conv3_wgrad_tile_slots()][867] floats; x: bf16 [B, H, W, 32], g4: [B, H, W, 4] bf16 (3 channels used)
int conv3_wgrad_tile(const void* g4, const void* x, float* partial, int B, int H, int W, int sm_count, cudaStream_t st)
{
    CUtensorMap mx;
    int rc = make_map_bhwc(&mx, x, B, H, W, 32, 32, kCwIW, kCwIH);
    if (rc) return rc;
    CwParams p;
    p.g4 = static_cast<const uint2*>(g4); p.partial = partial; p.B = B; p.H = H; p.W = W;
    p.tiles_x = (W + kCwTW - 1) / kCwTW; p.tiles_y = (H + kCwTH - 1) / kCwTH;
    p.tiles = (long long)B * p.tiles_y * p.tiles_x;
    constexpr size_t smem = kCwXBytes + kCwTH * kCwTW * 8 + 128;
    SEI_CUDA(allow_smem(conv3_wgrad_tile_kernel, smem));
    conv3_wgrad_tile_kernel<<<2 * sm_count, kCwThreads, smem, st>>>(mx, p);
    return finish_launch("conv3_wgrad_tile_kernel");
}

}  // namespace sei
