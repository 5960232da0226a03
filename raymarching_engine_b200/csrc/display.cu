// display.cu -- the present pass as a CUDA kernel (statically compiled for sm_100a).
//
// Restates /root/reference/client/public/shader/display.frag:20-61 as driven by
// /root/reference/client/src/index.tsx:25-59: per pixel a variable-radius Gaussian blur of the
// colour accumulator (radius from the accumulated depth-of-field radius), times brightness,
// gamma 1/2.2, alpha 1, converted to RGBA8.  Textures are NEAREST + REPEAT
// (/root/reference/client/src/renderer/LoadRenderJobContext.tsx:43-48).
//
// Arithmetic uses the exact policy of glsl_rt.h (single IEEE operations ptxas cannot fuse and
// the shared rm_math.h exp/pow) so the bytes are identical to the CPU oracle's.
//
// Layout: a thread presents one pixel; four neighbouring lanes' bytes leave as one 128-bit store.
// Rows are this rank's LOCAL rows; the texcoord uses the global row.  The blur reads neighbours from the local buffer, which is only correct when this
// rank owns the whole frame (n_ranks == 1) or the blur radius is 0 (preview mode);
// the host gathers the colour plane to one rank before presenting a blurred multi-GPU frame.
#include <cuda_runtime.h>

#define GLSL_NS xg
#define GLSL_FAST 0
#include "device_src/glsl_rt.h"

namespace xg {
namespace disp {

__device__ __forceinline__ float h2f(unsigned short h) {
    float f;
    asm("{ .reg .b16 t; mov.b16 t, %1; cvt.f32.f16 %0, t; }" : "=f"(f) : "h"(h));
    return f;
}

// gamma + RGBA8 quantisation of display.frag:54,61 as a table search: byte = #{k : x >= T[k]}, T = the 255 thresholds
// of device_src/display_gamma_table.inc - derived from, and verified for every float against, the binary64-evaluated
// pow(x, 1/2.2) it replaces (tools/gen_gamma_table.cpp; three such pow() per pixel were ~80 % of this kernel's
// instructions).  NaN and negative inputs compare false everywhere -> 0, like clamp(NaN -> 0).
__device__ const unsigned int rm_gamma_bits[255] = {
#include "device_src/display_gamma_table.inc"
};
__device__ __forceinline__ unsigned int gammaByte(const float* __restrict__ T, float x) {     // T[0] = -inf, T[1..255]
    int k = 0;
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1) k += (x >= T[k + step]) ? step : 0;
    return (unsigned int)k;
}

// one pixel of the present pass -> packed RGBA8 (r in the low byte)
__device__ __forceinline__ unsigned int displayPixel(const float4* __restrict__ color, const ushort4* __restrict__ nd, const float* __restrict__ T,
                                                     int x, int ly, int gy, int W, int H, int tileRows, int nRanks, int rank, float brightness) {
    const size_t idx = (size_t)ly * (size_t)W + (size_t)x;
    const vec2 texcoord(g_div(g_add((float)x, 0.5f), (float)W), g_div(g_add((float)gy, 0.5f), (float)H));
    const float PI_D = 3.1415926535f;                                    // display.frag:9
    const float ndw = g_mul(h2f(nd[idx].w), brightness);                  // display.frag:18
    const float kernelSize = clamp(g_mul(ndw, 200.0f), 0.0f, 16.0f);     // :20
    vec4 avg(0.0f);
    float sampleCount = 0.0f;
    const float sigma = g_mul(max(kernelSize, 1.0f), 0.3f);
    const float norm = g_div(1.0f, g_mul(g_mul(g_mul(2.0f, PI_D), sigma), sigma));
    const float twoSigma2 = g_mul(g_mul(2.0f, sigma), sigma);
    // NEAREST + REPEAT: texel = floor(uv * size) mod size.  The wrap is a compare-and-add when the index
    // is within one period of the range (always, for a +-16 pixel kernel on frames >= 16 pixels), an
    // integer remainder otherwise - same result.
    auto wrap = [](float f, int n) -> long long {
        if (f >= -(float)n && f < 2.0f * (float)n) {
            int v = (int)f;
            if (v < 0) v += n;
            else if (v >= n) v -= n;
            return v;
        }
        if (fabsf(f) < 1.0e9f) { int v = (int)f % n; return v < 0 ? v + n : v; }
        long long v = (long long)f % n;
        return v < 0 ? v + n : v;
    };
    if (kernelSize == 0.0f) {
        // One tap (every preview-mode frame; full mode without depth of field): y = xo = -+0, so both
        // texture offsets are +-0, uv = texcoord, and floor(fl(fl((g + 0.5) / N) * N)) == g for every g + 0.5 < 2^22
        // (two roundings of relative size 2^-24 move g + 0.5 by less than 0.5): the tap is the pixel itself.
        // d2 = 0 -> factor = norm (see below).  The arithmetic on the tap is the loop's, operation for operation.
        sampleCount = g_add(sampleCount, norm);
        const float4 c = color[idx];
        avg += vec4(c.x, c.y, c.z, c.w) * norm;
    } else {
        for (float y = -kernelSize; y <= kernelSize; y = g_add(y, 1.0f)) {   // :42-50
            // everything that depends on the row only, once per row (the same operations the shader repeats
            // for every tap of the row)
            const float texOffsetY = g_div(y, (float)H);
            const float uvY = g_add(texcoord.y, texOffsetY);
            const long long j = wrap(floor(g_mul(uvY, (float)H)), H);
            // global row j -> local row (identity when this rank owns every row)
            long long lj = j;
            if (nRanks > 1) {
                const long long tj = j / tileRows;
                lj = (tj / nRanks) * tileRows + (j - tj * tileRows);
                if (tj % nRanks != rank) lj = ly;   // not resident here (see header note)
            }
            const float4* __restrict__ row = color + (size_t)lj * (size_t)W;
            for (float xo = -kernelSize; xo <= kernelSize; xo = g_add(xo, 1.0f)) {
                const vec2 offset(xo, y);
                const float texOffsetX = g_div(xo, (float)W);
                // the centre tap (the only one when kernelSize is 0: every preview-mode frame): exp(-0) is exactly 1 in the
                // shared exp (rm_math.h dexp2_k: n = 0, f = 0, polynomial 1) and norm * 1 == norm, so the call is skipped
                const float d2 = dot(offset, offset);
                const float factor = (d2 == 0.0f) ? norm : g_mul(norm, rmx::exp_ft(-g_div(d2, twoSigma2)));   // exp of the exact policy, table-driven coefficients
                sampleCount = g_add(sampleCount, factor);
                const float uvX = g_add(texcoord.x, texOffsetX);
                const long long i = wrap(floor(g_mul(uvX, (float)W)), W);
                const float4 c = row[i];
                avg += vec4(c.x, c.y, c.z, c.w) * factor;
            }
        }
    }
    avg /= sampleCount;
    const vec3 base = vec3(avg.x, avg.y, avg.z) * brightness;
    // :54 pow(vec4(base, 1), 1/2.2) and the RGBA8 conversion: alpha = pow(1, y) = 1 -> 255
    return gammaByte(T, base.x) | (gammaByte(T, base.y) << 8) | (gammaByte(T, base.z) << 16) | 0xff000000u;
}

// Layout: one pixel per thread, a warp per 32-pixel row segment (a warp reads 32 consecutive float4 texels - fully
// coalesced - and the lanes of a blurred neighbourhood diverge as little as the scene allows); the RGBA8 results of four neighbouring lanes are collected with
// shuffles and written by every fourth lane as ONE 128-bit store (a warp: one full 128-byte line).  Rows whose width is
// not a multiple of 4 use scalar stores.
__global__ void __launch_bounds__(256) rm_display_kernel(const float4* __restrict__ color, const ushort4* __restrict__ nd,
                                                         uchar4* __restrict__ out, uchar4* __restrict__ gather, int W, int localRows, int H,
                                                         int tileRows, int nRanks, int rank, float brightness) {
    __shared__ float T[256];
    T[threadIdx.x] = threadIdx.x ? __uint_as_float(rm_gamma_bits[threadIdx.x - 1]) : __int_as_float(0xff800000);
    __syncthreads();
    // a block presents a 32 x 8 pixel tile, one row segment per warp: the +-16-row taps of a blurred neighbourhood are
    // shared between the rows of the tile through L1
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ly = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ly >= localRows) return;                       // (uniform per warp)
    const int t = ly / tileRows;
    const int gy = (t * nRanks + rank) * tileRows + (ly - t * tileRows);
    const unsigned int px = x < W ? displayPixel(color, nd, T, x, ly, gy, W, H, tileRows, nRanks, rank, brightness) : 0u;
    const int lane = threadIdx.x & 31, q = lane & ~3;
    const unsigned int p0 = __shfl_sync(0xffffffffu, px, q), p1 = __shfl_sync(0xffffffffu, px, q + 1),
                       p2 = __shfl_sync(0xffffffffu, px, q + 2), p3 = __shfl_sync(0xffffffffu, px, q + 3);
    if (x >= W) return;
    const size_t idx = (size_t)ly * (size_t)W + (size_t)x;
    // fused tile gather (multi-GPU row-tile sharding): the same pixels go straight into the assembled
    // full frame - usually rank 0's memory mapped over NVLink (CUDA IPC / peer access) - at their GLOBAL row.  Full
    // 128-byte lines per warp, so the peer stores are fully coalesced; no separate collective moves pixels.
    const size_t gidx = (size_t)gy * (size_t)W + (size_t)x;
    if ((W & 3) == 0) {
        if ((lane & 3) == 0) {                         // x is a multiple of 4 and x + 3 < W
            const uint4 v = make_uint4(p0, p1, p2, p3);
            *reinterpret_cast<uint4*>(out + idx) = v;
            if (gather) *reinterpret_cast<uint4*>(gather + gidx) = v;
        }
    } else {
        reinterpret_cast<unsigned int*>(out)[idx] = px;
        if (gather) reinterpret_cast<unsigned int*>(gather)[gidx] = px;
    }
}

}  // namespace disp
}  // namespace xg

// Row scatter for multi-GPU frames that need the display blur (full mode with depth of field): a rank
// copies its local rows of an accumulator plane into a FULL-FRAME plane (usually rank 0's memory over
// NVLink) at their global rows; rank 0 then runs the display pass over the assembled planes
// (SURVEY.md 8e caveat: the blur reads +-16 neighbour rows with REPEAT wrap).  16-byte chunks.
__global__ void __launch_bounds__(256) rm_scatter_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int rowVec,
                                                              int localRows, int tileRows, int nRanks, int rank) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly = blockIdx.y;
    if (v >= rowVec || ly >= localRows) return;
    const int t = ly / tileRows;
    const int gy = (t * nRanks + rank) * tileRows + (ly - t * tileRows);
    dst[(size_t)gy * (size_t)rowVec + (size_t)v] = src[(size_t)ly * (size_t)rowVec + (size_t)v];
}
extern "C" cudaError_t rmb_launch_scatter_rows(const void* src, void* dst, int row_bytes, int local_rows, int tile_rows, int n_ranks,
                                               int rank, cudaStream_t stream) {
    const int rowVec = row_bytes / 16;
    dim3 block(256, 1, 1), grid((unsigned)((rowVec + 255) / 256), (unsigned)local_rows, 1);
    rm_scatter_rows_kernel<<<grid, block, 0, stream>>>((const uint4*)src, (uint4*)dst, rowVec, local_rows, tile_rows, n_ranks, rank);
    return cudaGetLastError();
}

// FP32 FMA throughput probe: the roofline denominator for this FP32-bound path is not in
// MEASURED_PEAKS.json (which has HBM and bf16 only), so bench.py measures it live.  Each thread
// runs 8 independent FFMA chains; 2 flop per FFMA.
__global__ void __launch_bounds__(256) rm_fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// The same probe with Blackwell's packed FFMA2 (fma.rn.f32x2: two FMAs per lane per instruction): shows
// whether packing raises the FP32 ceiling itself or only saves issue slots.  4 flop per FFMA2.
__global__ void __launch_bounds__(256) rm_fp32x2_peak_kernel(float* out, int iters, float a, float b) {
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int k = 0; k < 8; k++) { float v = threadIdx.x + k; asm("mov.b64 %0, {%1, %1};" : "=l"(x[k]) : "f"(v)); }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[j]) : "l"(aa), "l"(bb));
        }
    }
    float s = 0.0f;
    for (int k = 0; k < 8; k++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" cudaError_t rmb_launch_fp32x2_peak(float* scratch, int blocks, int iters, cudaStream_t stream) {
    rm_fp32x2_peak_kernel<<<blocks, 256, 0, stream>>>(scratch, iters, 0.999999f, 1e-7f);
    return cudaGetLastError();
}

// Launches the probe with `blocks` blocks of 256 threads; flops = blocks*256*iters*16*8*2.
extern "C" cudaError_t rmb_launch_fp32_peak(float* scratch, int blocks, int iters, cudaStream_t stream) {
    rm_fp32_peak_kernel<<<blocks, 256, 0, stream>>>(scratch, iters, 0.999999f, 1e-7f);
    return cudaGetLastError();
}

extern "C" cudaError_t rmb_launch_display(const void* color, const void* nd, void* rgba8, void* gather, int W, int local_rows, int H,
                                          int tile_rows, int n_ranks, int rank, float brightness, cudaStream_t stream) {
    dim3 block(256, 1, 1), grid((unsigned)((W + 31) / 32), (unsigned)((local_rows + 7) / 8), 1);
    xg::disp::rm_display_kernel<<<grid, block, 0, stream>>>((const float4*)color, (const ushort4*)nd, (uchar4*)rgba8, (uchar4*)gather, W,
                                                            local_rows, H, tile_rows, n_ranks, rank, brightness);
    return cudaGetLastError();
}
