// scale.cu -- the random scale transform (reference: src/transforms.py:5-109).
//
// padded_downsampling_transform = get_downsampling_grid (:27-43) + F.grid_sample(bicubic,
// reflection, align_corners=True) (:77-83).  The grid is a per-image axis-aligned affine map, so
// the 4x4 bicubic gather is SEPARABLE: row taps depend only on the output row, column taps only
// on the output column.  The grid is never materialised; its coordinates are recomputed with the
// reference's exact fp32 rounding sequence (scale_src_coord).
//
// Tiled kernel (scale_band_kernel): one CTA = TH output rows x full width of one plane.  The
// source rows the band touches (a contiguous range after reflection, <= TH/rate + 4 rows) are
// staged by bulk async copies (TMA engine); a vertical 4-tap pass (128-bit shared loads) is
// followed by a horizontal 4-tap gather within shared memory.  Bands whose source range does not
// fit the staging buffer (rates < 0.5) read their taps from global memory instead.
// Direct kernel: one thread per output element, for shapes the tiled path does not take.
#include "tile_ops.cuh"
#include <algorithm>
#include <stdlib.h>

namespace sei {

constexpr int kScaleThreads = 256;

struct ScaleParams {
    const float* x;
    float* out;
    const float* rate;
    const float* center;
    int C, S, TH, nbands, SRC_MAX;
    float two_over_S;
    long long total_bands;
};

// de-interleaved column layout of the vertical-pass result: even columns first, odd columns HALF floats later
// (HALF = S/2 + 16: the two halves start 16 banks apart).  At rate 0.5 the horizontal gather of 32 neighbouring
// outputs reads every second column -- a 2-way bank conflict in a plain row (ncu: 3.6 M conflicts on 4.4 M shared
// loads, the shared-memory pipe at 74 %) and consecutive words here.
__host__ __device__ constexpr int scale_tmp_half(int S) { return S / 2 + 16; }
__host__ __device__ constexpr int scale_tmp_pitch(int S) { return 2 * scale_tmp_half(S); }
__device__ __forceinline__ int scale_tmp_pos(int c, int S) { return (c >> 1) + (c & 1) * scale_tmp_half(S); }

// Persistent, double-buffered: a CTA walks bands blockIdx.x, blockIdx.x + gridDim.x, ...  While band i is resampled,
// warp 0 has already computed the row taps of band i+1, derived its source-row range with warp shuffles and issued
// the bulk copy into the other staging buffer, and the remaining warps have written its column taps; the first
// version loaded, waited and computed strictly in sequence (17 % of its stall samples sat in the mbarrier wait).
template <int ST, int THT>
__global__ void __launch_bounds__(kScaleThreads, 2) scale_band_kernel(const __grid_constant__ ScaleParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    __shared__ int s_lo[2], s_n[2], s_img[2];

    const int S = ST ? ST : p.S;
    const int TH = THT ? THT : p.TH;
    const int TP = scale_tmp_pitch(S);
    float* sSrc = reinterpret_cast<float*>(smem_raw);                         // [2][SRC_MAX][S]
    float* sTmp = sSrc + (size_t)2 * p.SRC_MAX * S;                            // [TH][TP]
    AxisTap* colT = reinterpret_cast<AxisTap*>(sTmp + (size_t)TH * TP);      // [2][S]   (idx = de-interleaved position)
    AxisTap* rowT = colT + 2 * S;                                              // [2][TH]  (idx relative to the staged rows)
    const long long total = p.total_bands;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // every CTA owns a contiguous range of bands, so consecutive bands mostly belong to the same image and share
    // their column taps (a tap costs ~120 instructions -- reflections need integer divisions -- against ~10 per
    // output of the passes themselves): colT[slot] is recomputed only when the image changes
    const long long w_begin = total * blockIdx.x / gridDim.x, w_end = total * (blockIdx.x + 1) / gridDim.x;
    int cur_img = -1, cur_slot = 1;

    // taps of band w into buffer `buf`; warp 0: row taps, source range, bulk copy.  Requires sSrc[buf] to be free.
    auto prepare = [&](long long w, int buf) {
        const int band = (int)(w % p.nbands);
        const long long plane = w / p.nbands;
        const int b = (int)(plane / p.C);
        const int r0 = band * TH, th = min(TH, S - r0);
        const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        if (warp == 0) {
            const float cy = __ldg(p.center + 2 * b + 1);
            AxisTap t;
            int lo = S, hi = -1;
            if (lane < th) {
                scale_axis_tap(r0 + lane, S, p.two_over_S, inv_rate, cy, t);
                lo = min(min(t.idx[0], t.idx[1]), min(t.idx[2], t.idx[3]));
                hi = max(max(t.idx[0], t.idx[1]), max(t.idx[2], t.idx[3]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            const int nsrc = hi - lo + 1;
            const bool staged = nsrc <= p.SRC_MAX;
            if (lane < th) {
                if (staged) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) t.idx[a] -= lo;
                }
                rowT[buf * TH + lane] = t;
            }
            if (lane == 0) {
                s_lo[buf] = lo;
                s_n[buf] = staged ? nsrc : 0;
                if (staged) {
                    const uint32_t row_bytes = (uint32_t)S * 4u;
                    fence_proxy_async();          // the buffer was read by generic loads two bands ago
                    mbar_arrive_expect_tx(&bar[buf], (uint32_t)nsrc * row_bytes);
                    bulk_load_rows_circular(reinterpret_cast<unsigned char*>(sSrc + (size_t)buf * p.SRC_MAX * S),
                                            reinterpret_cast<const unsigned char*>(p.x + (size_t)plane * S * S), S, row_bytes,
                                            lo, nsrc, &bar[buf]);
                }
            }
        }
        if (b != cur_img) {                // uniform across the CTA
            cur_img = b;
            cur_slot ^= 1;
            if (warp != 0) {
                const float cx = __ldg(p.center + 2 * b);
                for (int j = threadIdx.x - 32; j < S; j += kScaleThreads - 32) {
                    AxisTap t;
                    scale_axis_tap(j, S, p.two_over_S, inv_rate, cx, t);
#pragma unroll
                    for (int a = 0; a < 4; ++a) t.idx[a] = scale_tmp_pos(t.idx[a], S);
                    colT[cur_slot * S + j] = t;
                }
            }
        }
        if (threadIdx.x == 0) s_img[buf] = cur_slot;
    };

    if (w_begin < w_end) prepare(w_begin, 0);
    uint32_t phases = 0u;                  // bit b: parity of buffer b's barrier (a phase completes only for staged bands)
    int it = 0;
    for (long long w = w_begin; w < w_end; ++w, ++it) {
        const int buf = it & 1;
        if (w + 1 < w_end) prepare(w + 1, buf ^ 1);
        __syncthreads();                                        // taps of this band (written one iteration ago) visible
        const int band = (int)(w % p.nbands);
        const long long plane = w / p.nbands;
        const int r0 = band * TH, th = min(TH, S - r0);
        const bool staged = s_n[buf] > 0;
        const AxisTap* rT = rowT + buf * TH;
        if (staged) {
            mbar_wait(&bar[buf], (phases >> buf) & 1u);
            phases ^= 1u << buf;
        }

        // ---- vertical pass: sTmp[r][pos(c)] = sum_a wy[r][a] * src[iy[r][a]][c]
        // (two instantiations so that the staged path compiles to shared loads and the other to read-only global loads)
        const int CW = S >> 2;
        auto vpass = [&](const float* __restrict__ src, auto ld) {
            for (int item = threadIdx.x; item < th * CW; item += kScaleThreads) {
                const int r = item / CW, c4 = item - r * CW;
                const AxisTap t = rT[r];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float4 v = ld(reinterpret_cast<const float4*>(src + (size_t)t.idx[a] * S + 4 * c4));
                    acc.x = fmaf(t.w[a], v.x, acc.x); acc.y = fmaf(t.w[a], v.y, acc.y);
                    acc.z = fmaf(t.w[a], v.z, acc.z); acc.w = fmaf(t.w[a], v.w, acc.w);
                }
                float* d = sTmp + (size_t)r * TP + 2 * c4;
                *reinterpret_cast<float2*>(d) = make_float2(acc.x, acc.z);                               // columns 4c4, 4c4+2
                *reinterpret_cast<float2*>(d + scale_tmp_half(S)) = make_float2(acc.y, acc.w);           // columns 4c4+1, 4c4+3
            }
        };
        if (staged) vpass(sSrc + (size_t)buf * p.SRC_MAX * S, [](const float4* q) { return *q; });
        else vpass(p.x + (size_t)plane * S * S, [](const float4* q) { return __ldg(q); });
        __syncthreads();

        // ---- horizontal pass: out[r][j] = sum_b wx[j][b] * sTmp[r][pos(ix[j][b])]
        float* oplane = p.out + (size_t)plane * S * S + (size_t)r0 * S;
        const AxisTap* cT = colT + s_img[buf] * S;
        for (int j = threadIdx.x; j < S; j += kScaleThreads) {
            const AxisTap t = cT[j];
            if (THT && th == THT) {            // full band: fully unrolled, every address is base + immediate
#pragma unroll
                for (int r = 0; r < (THT ? THT : 1); ++r) __stcs(oplane + (size_t)r * S + j, scale_hgather(sTmp + (size_t)r * TP, t));
            } else {
#pragma unroll 4
                for (int r = 0; r < th; ++r) __stcs(oplane + (size_t)r * S + j, scale_hgather(sTmp + (size_t)r * TP, t));
            }
        }
        __syncthreads();                                        // sTmp and the staging buffer are free again
    }
}


// ------------------------------------------------------------------ rows kernel (round 2)
// One CTA = one band of TH output rows of ALL C planes of one image (the taps of an image are shared by its planes).
// Nothing is staged by copies: the vertical 4-tap pass reads its source rows straight from global memory with 128-bit
// read-only loads (the taps of neighbouring output rows overlap, so about half of them hit L1), 8 independent loads in
// flight per thread; its result goes to a shared intermediate (double-buffered across planes: one barrier per plane)
// from which the horizontal 4-tap gather produces coalesced streaming stores.  ~45 KB of shared memory and no mbarrier
// round trips: four CTAs (1024 threads) per SM hide the gather latency that bound the staged kernel of round 1
// (ncu: issue slots 35 %, shared pipe 65 %, occupancy 24 %; 75 -> 59 us).
// The intermediate keeps round 1's even / odd column split (scale_tmp_pos).  A simulation of the gather's bank
// pattern over the reference's rates and random centres (32 lanes read 32 words spread over 43 - 64 columns) gives 1.24
// wavefronts per instruction at rate 0.5 and 1.87 at 0.75 for it; no layout of the families (c mod m) * Q + c / m,
// m = 1, 2, 4, 8, does better on both (a 4-way split with Q = 8 mod 32 was built and measured: 1.51 / 1.89, same time).  A sliding-window vertical pass (only the source rows that enter the 4-row window
// of the next output row are loaded: 1.5 - 2.2 loads per output vector instead of 4) was built and measured as well:
// it halves the L1 traffic of the pass but its serial walk and register rotation cost more issue slots than the loads
// saved (71 us against 59 us), so the independent-loads form stayed.
struct ScaleRowsParams {
    const float* x;
    float* out;
    const float* rate;
    const float* center;
    int C, S, nbands;
    float two_over_S;
};

template <int ST, int TH>
__global__ void __launch_bounds__(kScaleThreads, 4) scale_rows_kernel(const __grid_constant__ ScaleRowsParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = ST ? ST : p.S;
    float* sV = reinterpret_cast<float*>(smem_raw);                              // [2][TH][TP]
    AxisTap* colT = reinterpret_cast<AxisTap*>(sV + 2 * TH * scale_tmp_pitch(S));   // [S]  (idx = position in a V row)
    AxisTap* rowT = colT + S;                                                    // [TH] (idx = row offset in elements)

    const int band = blockIdx.x % p.nbands;
    const int b = blockIdx.x / p.nbands;
    const int r0 = band * TH, th = min(TH, S - r0);
    const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
    const float cx = __ldg(p.center + 2 * b), cy = __ldg(p.center + 2 * b + 1);
    const int TP = scale_tmp_pitch(S), HALF = scale_tmp_half(S);   // compile-time for the specialised sizes

    for (int j = threadIdx.x; j < S + TH; j += kScaleThreads) {
        AxisTap t;
        if (j < S) {
            scale_axis_tap(j, S, p.two_over_S, inv_rate, cx, t);
#pragma unroll
            for (int a = 0; a < 4; ++a) t.idx[a] = scale_tmp_pos(t.idx[a], S);
            colT[j] = t;
        } else if (j - S < th) {
            scale_axis_tap(r0 + j - S, S, p.two_over_S, inv_rate, cy, t);
#pragma unroll
            for (int a = 0; a < 4; ++a) t.idx[a] *= S;
            rowT[j - S] = t;
        }
    }
    __syncthreads();

    const int CW = S >> 2;
    const size_t plane_elems = (size_t)S * S;
    for (int c = 0; c < p.C; ++c) {
        const float* __restrict__ xp = p.x + ((size_t)b * p.C + c) * plane_elems;
        float* __restrict__ V = sV + (c & 1) * TH * TP;
        // ---- vertical pass: V[r][pos(col)] = sum_a wy[r][a] * x[iy[r][a]][col]
        constexpr int kItems = 2;                      // work items per thread whose loads are issued together
        const int nitems = th * CW;
        for (int base = threadIdx.x; base < nitems; base += kItems * kScaleThreads) {
            float4 v[kItems][4];
            float w[kItems][4];
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const int item = base + k * kScaleThreads;
                if (item < nitems) {
                    const int r = item / CW, c4 = item - r * CW;
                    const AxisTap t = rowT[r];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        v[k][a] = __ldg(reinterpret_cast<const float4*>(xp + t.idx[a] + 4 * c4));
                        w[k][a] = t.w[a];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const int item = base + k * kScaleThreads;
                if (item < nitems) {
                    const int r = item / CW, c4 = item - r * CW;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        acc.x = fmaf(w[k][a], v[k][a].x, acc.x); acc.y = fmaf(w[k][a], v[k][a].y, acc.y);
                        acc.z = fmaf(w[k][a], v[k][a].z, acc.z); acc.w = fmaf(w[k][a], v[k][a].w, acc.w);
                    }
                    float* d = V + r * TP + 2 * c4;
                    *reinterpret_cast<float2*>(d) = make_float2(acc.x, acc.z);             // columns 4c4, 4c4+2
                    *reinterpret_cast<float2*>(d + HALF) = make_float2(acc.y, acc.w);      // columns 4c4+1, 4c4+3
                }
            }
        }
        __syncthreads();
        // ---- horizontal pass: out[r][j] = sum_b wx[j][b] * V[r][pos(ix[j][b])]
        float* __restrict__ oplane = p.out + ((size_t)b * p.C + c) * plane_elems + (size_t)r0 * S;
        for (int j = threadIdx.x; j < S; j += kScaleThreads) {
            const AxisTap t = colT[j];
            if (th == TH) {
#pragma unroll
                for (int r = 0; r < TH; ++r) __stcs(oplane + r * S + j, scale_hgather(V + r * TP, t));
            } else {
                for (int r = 0; r < th; ++r) __stcs(oplane + r * S + j, scale_hgather(V + r * TP, t));
            }
        }
        // no barrier here: the next plane's vertical pass writes the other buffer, and the barrier after it orders
        // this plane's reads before the writes of the plane after next
    }
}

static size_t scale_rows_smem(int th, int S)
{
    return (size_t)2 * th * scale_tmp_pitch(S) * 4 + (size_t)(S + th) * sizeof(AxisTap);
}

struct ScaleDirectParams {
    const float* x;
    float* out;
    const float* rate;
    const float* center;
    int C, S;
    int Ssrc;            // source size (== S except behind the anti-aliasing pre-filter, which shrinks the source)
    float two_over_S;
    long long total;
};

__global__ void __launch_bounds__(256) scale_direct_kernel(const __grid_constant__ ScaleDirectParams p)
{
    const int S = p.S, Ss = p.Ssrc;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % S);
        const long long t = idx / S;
        const int i = (int)(t % S);
        const long long plane = t / S;
        const int b = (int)(plane / p.C);
        const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        AxisTap ty, tx;
        scale_axis_tap(i, Ss, p.two_over_S, inv_rate, __ldg(p.center + 2 * b + 1), ty);
        scale_axis_tap(j, Ss, p.two_over_S, inv_rate, __ldg(p.center + 2 * b), tx);
        const float* xp = p.x + plane * (long long)Ss * Ss;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float* row = xp + (size_t)ty.idx[a] * Ss;
            float h = __ldg(row + tx.idx[0]) * tx.w[0];
            h = fmaf(__ldg(row + tx.idx[1]), tx.w[1], h);
            h = fmaf(__ldg(row + tx.idx[2]), tx.w[2], h);
            h = fmaf(__ldg(row + tx.idx[3]), tx.w[3], h);
            acc = fmaf(h, ty.w[a], acc);
        }
        p.out[idx] = acc;
    }
}

// transpose of the transform w.r.t. the image (autograd through grid_sample when stop_gradient is off, reference
// src/losses/__init__.py:117-122 with no_grad=False): every output gradient is scattered to its 4x4 taps with
// fp32 atomics into a zeroed buffer.  Not on the default path (stop_gradient=True), so a direct kernel.
__global__ void __launch_bounds__(256) scale_backward_kernel(const __grid_constant__ ScaleDirectParams p)
{
    const int S = p.S, Ss = p.Ssrc;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % S);
        const long long t = idx / S;
        const int i = (int)(t % S);
        const long long plane = t / S;
        const int b = (int)(plane / p.C);
        const float inv_rate = __fdiv_rn(1.0f, __ldg(p.rate + b));
        AxisTap ty, tx;
        scale_axis_tap(i, Ss, p.two_over_S, inv_rate, __ldg(p.center + 2 * b + 1), ty);
        scale_axis_tap(j, Ss, p.two_over_S, inv_rate, __ldg(p.center + 2 * b), tx);
        float* gx = p.out + plane * (long long)Ss * Ss;
        const float g = __ldg(p.x + idx);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float ga = g * ty.w[a];
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(gx + (size_t)ty.idx[a] * Ss + tx.idx[c], ga * tx.w[c]);
        }
    }
}

__global__ void scale_params_kernel(const float* u_rate, const float* u_center, int B, int n_rates,
                                    float r0, float r1, float r2, float r3, float* rate, float* center)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // sample_from (src/transforms.py:5-11): values[floor(N * U)]
    int k = (int)floorf(__fmul_rn((float)n_rates, u_rate[b]));
    k = min(max(k, 0), n_rates - 1);
    rate[b] = k == 0 ? r0 : (k == 1 ? r1 : (k == 2 ? r2 : r3));
    // center = 2 * U - 1 (:20-22)
    center[2 * b] = __fsub_rn(__fmul_rn(2.0f, u_center[2 * b]), 1.0f);
    center[2 * b + 1] = __fsub_rn(__fmul_rn(2.0f, u_center[2 * b + 1]), 1.0f);
}

// staging rows for a band of th output rows at the smallest rate the reference samples (0.5)

// SEI_SCALE_NOSTAGE=1 (experiment): no staging buffers, the vertical pass reads its taps through L1
static bool scale_nostage()
{
    const char* e = getenv("SEI_SCALE_NOSTAGE");
    return e && *e == '1';
}
static int scale_src_rows(int th) { return scale_nostage() ? 0 : 2 * th + 6; }

static size_t scale_band_smem(int th, int S)
{
    return ((size_t)2 * scale_src_rows(th) * S + (size_t)th * scale_tmp_pitch(S)) * 4 + (size_t)(2 * S + 2 * th) * sizeof(AxisTap);
}

// band height: 16 rows (two staging buffers of 38 rows + taps = 111 KB at S = 256: two CTAs per SM); SEI_SCALE_TH overrides
int scale_pick_band_rows(int S, int smem_optin, size_t* smem_out)
{
    const size_t budget = std::min((size_t)smem_optin, (size_t)112 * 1024);
    int best = 0;
    for (int th = 8; th <= 16; th += 8)
        if (scale_band_smem(th, S) <= budget) best = th;
    if (best == 0 && scale_band_smem(8, S) <= (size_t)smem_optin) best = 8;
    if (const char* e = getenv("SEI_SCALE_TH")) {     // tuning override
        const int f = atoi(e);
        if (f >= 8 && f % 8 == 0 && f <= 32 && scale_band_smem(f, S) <= (size_t)smem_optin) best = f;
    }
    if (best == 0) return 0;
    best = std::min(best, ((S + 7) / 8) * 8);
    *smem_out = scale_band_smem(best, S);
    return best;
}

}  // namespace sei

using namespace sei;

extern "C" int sei_scale_transform_f32(const float* x, float* out, int B, int C, int S,
                                       const float* rate, const float* center, int path, void* stream)
{
    SEI_REQUIRE(x && out && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0, "bad shape B=%d C=%d S=%d", B, C, S);
    SEI_REQUIRE(path >= SEI_PATH_AUTO && path <= SEI_PATH_TILED, "bad path %d", path);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const long long planes = (long long)B * C;
    size_t smem = 0;
    const int TH = (S % 4 == 0 && S >= 8 && aligned16(x) && aligned16(out)) ? scale_pick_band_rows(S, dp.smem_optin, &smem) : 0;
    const bool tiled_ok = TH > 0 && TH <= kScaleThreads && planes * ((S + TH - 1) / TH) < (1ll << 31);
    const float two_over_S = (float)(2.0 / (double)S);
    const char* v1 = getenv("SEI_SCALE_V1");              // A/B switch: the staged (TMA) band kernel of round 1
    const bool rows_ok = S % 4 == 0 && S >= 8 && S <= 2048 && aligned16(x) && aligned16(out) && (long long)B * ((S + 7) / 8) < (1ll << 31);
    SEI_REQUIRE(path != SEI_PATH_TILED || tiled_ok || rows_ok, "tiled scale-transform path not available for S=%d", S);
    if (rows_ok && path != SEI_PATH_DIRECT && !(v1 && *v1 == '1' && tiled_ok)) {
        ScaleRowsParams q;
        q.x = x; q.out = out; q.rate = rate; q.center = center; q.C = C; q.S = S; q.two_over_S = two_over_S;
        const int th_env = getenv("SEI_SCALE_ROWS_TH") ? atoi(getenv("SEI_SCALE_ROWS_TH")) : 0;
        const int THr = th_env == 8 || th_env == 16 ? th_env : (S <= 256 ? 16 : 8);
        q.nbands = (S + THr - 1) / THr;
        const size_t sm = scale_rows_smem(THr, S);
        SEI_REQUIRE(sm <= (size_t)dp.smem_optin, "scale transform: S=%d needs %zu bytes of shared memory", S, sm);
        const unsigned grid = (unsigned)((long long)B * q.nbands);
        if (S == 256 && THr == 16) {
            SEI_CUDA(allow_smem(scale_rows_kernel<256, 16>, sm));
            scale_rows_kernel<256, 16><<<grid, kScaleThreads, sm, st>>>(q);
        } else if (S == 256 && THr == 8) {
            SEI_CUDA(allow_smem(scale_rows_kernel<256, 8>, sm));
            scale_rows_kernel<256, 8><<<grid, kScaleThreads, sm, st>>>(q);
        } else if (S == 512 && THr == 8) {
            SEI_CUDA(allow_smem(scale_rows_kernel<512, 8>, sm));
            scale_rows_kernel<512, 8><<<grid, kScaleThreads, sm, st>>>(q);
        } else if (THr == 16) {
            SEI_CUDA(allow_smem(scale_rows_kernel<0, 16>, sm));
            scale_rows_kernel<0, 16><<<grid, kScaleThreads, sm, st>>>(q);
        } else {
            SEI_CUDA(allow_smem(scale_rows_kernel<0, 8>, sm));
            scale_rows_kernel<0, 8><<<grid, kScaleThreads, sm, st>>>(q);
        }
        return finish_launch("scale_rows_kernel");
    }
    if (tiled_ok && path != SEI_PATH_DIRECT) {
        ScaleParams p;
        p.x = x; p.out = out; p.rate = rate; p.center = center;
        p.C = C; p.S = S; p.TH = TH; p.nbands = (S + TH - 1) / TH; p.SRC_MAX = scale_src_rows(TH);
        p.two_over_S = two_over_S;
        p.total_bands = planes * p.nbands;
        const int ctas_per_sm = std::max(1, std::min(scale_nostage() ? 3 : 2, (int)((size_t)227 * 1024 / (smem + 1024))));
        const unsigned grid = (unsigned)std::min<long long>(p.total_bands, (long long)dp.sm_count * ctas_per_sm);
        if (S == 256 && TH == 16) {
            SEI_CUDA(allow_smem(scale_band_kernel<256, 16>, smem));
            scale_band_kernel<256, 16><<<grid, kScaleThreads, smem, st>>>(p);
        } else if (S == 512 && TH == 8) {
            SEI_CUDA(allow_smem(scale_band_kernel<512, 8>, smem));
            scale_band_kernel<512, 8><<<grid, kScaleThreads, smem, st>>>(p);
        } else {
            SEI_CUDA(allow_smem(scale_band_kernel<0, 0>, smem));
            scale_band_kernel<0, 0><<<grid, kScaleThreads, smem, st>>>(p);
        }
        return finish_launch("scale_band_kernel");
    }
    ScaleDirectParams d;
    d.x = x; d.out = out; d.rate = rate; d.center = center; d.C = C; d.S = S; d.Ssrc = S;
    d.two_over_S = two_over_S;
    d.total = planes * (long long)S * S;
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_direct_kernel<<<grid, 256, 0, st>>>(d);
    return finish_launch("scale_direct_kernel");
}

// The same resampling from a source of another size: x is [B, C, Ssrc, Ssrc], out [B, C, S, S]; the grid is that of an
// S x S image (reference src/transforms.py:63-83 with antialiased=True: the grid is built for the ORIGINAL shape and
// grid_sample reads the pre-filtered, smaller image through normalised coordinates).
extern "C" int sei_scale_transform_src_f32(const float* x, float* out, int B, int C, int Ssrc, int S,
                                           const float* rate, const float* center, void* stream)
{
    SEI_REQUIRE(x && out && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0 && Ssrc > 0, "bad shape B=%d C=%d Ssrc=%d S=%d", B, C, Ssrc, S);
    if (B == 0) return 0;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    ScaleDirectParams d;
    d.x = x; d.out = out; d.rate = rate; d.center = center; d.C = C; d.S = S; d.Ssrc = Ssrc;
    d.two_over_S = (float)(2.0 / (double)S);
    d.total = (long long)B * C * S * S;
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_direct_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d);
    return finish_launch("scale_direct_kernel");
}

extern "C" int sei_scale_transform_backward_f32(const float* gout, float* gx, int B, int C, int S,
                                                const float* rate, const float* center, void* stream)
{
    SEI_REQUIRE(gout && gx && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0, "bad shape B=%d C=%d S=%d", B, C, S);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    ScaleDirectParams d;
    d.x = gout; d.out = gx; d.rate = rate; d.center = center; d.C = C; d.S = S; d.Ssrc = S;
    d.two_over_S = (float)(2.0 / (double)S);
    d.total = (long long)B * C * S * S;
    SEI_CUDA(cudaMemsetAsync(gx, 0, (size_t)d.total * sizeof(float), st));
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_backward_kernel<<<grid, 256, 0, st>>>(d);
    return finish_launch("scale_backward_kernel");
}

// transpose of sei_scale_transform_src_f32: gout [B, C, S, S] -> gx [B, C, Ssrc, Ssrc]
extern "C" int sei_scale_transform_src_backward_f32(const float* gout, float* gx, int B, int C, int Ssrc, int S,
                                                    const float* rate, const float* center, void* stream)
{
    SEI_REQUIRE(gout && gx && rate && center, "null pointer argument");
    SEI_REQUIRE(B >= 0 && C > 0 && S > 0 && Ssrc > 0, "bad shape B=%d C=%d Ssrc=%d S=%d", B, C, Ssrc, S);
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    ScaleDirectParams d;
    d.x = gout; d.out = gx; d.rate = rate; d.center = center; d.C = C; d.S = S; d.Ssrc = Ssrc;
    d.two_over_S = (float)(2.0 / (double)S);
    d.total = (long long)B * C * S * S;
    SEI_CUDA(cudaMemsetAsync(gx, 0, (size_t)B * C * Ssrc * Ssrc * sizeof(float), st));
    const unsigned grid = (unsigned)std::min<long long>((d.total + 255) / 256, (long long)dp.sm_count * 32);
    scale_backward_kernel<<<grid, 256, 0, st>>>(d);
    return finish_launch("scale_backward_kernel");
}

extern "C" int sei_scale_params_f32(const float* u_rate, const float* u_center, int B,
                                    const float* rates_host, int n_rates, float* rate, float* center,
                                    void* stream)
{
    SEI_REQUIRE(u_rate && u_center && rates_host && rate && center, "null pointer argument");
    SEI_REQUIRE(n_rates >= 1 && n_rates <= 4, "n_rates %d unsupported (1..4)", n_rates);
    if (B <= 0) return 0;
    float r[4] = {rates_host[0], rates_host[0], rates_host[0], rates_host[0]};
    for (int i = 0; i < n_rates; ++i) r[i] = rates_host[i];
    scale_params_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        u_rate, u_center, B, n_rates, r[0], r[1], r[2], r[3], rate, center);
    return finish_launch("scale_params_kernel");
}
