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
// Layout: one thread per pixel, consecutive threads along x; each thread writes one uchar4
// (a warp writes one full 128-byte line).  Rows are this rank's LOCAL rows; the texcoord uses the
// global row.  The blur reads neighbours from the local buffer, which is only correct when this
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

__global__ void __launch_bounds__(256) rm_display_kernel(const float4* __restrict__ color, const ushort4* __restrict__ nd,
                                                         uchar4* __restrict__ out, uchar4* __restrict__ gather, int W, int localRows, int H,
                                                         int tileRows, int nRanks, int rank, float brightness) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly = blockIdx.y;
    if (x >= W || ly >= localRows) return;
    const int t = ly / tileRows;
    const int gy = (t * nRanks + rank) * tileRows + (ly - t * tileRows);
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
            const float factor = g_mul(norm, rmx::exp_ft(-g_div(dot(offset, offset), twoSigma2)));   // exp of the exact policy, table-driven coefficients
            sampleCount = g_add(sampleCount, factor);
            const float uvX = g_add(texcoord.x, texOffsetX);
            const long long i = wrap(floor(g_mul(uvX, (float)W)), W);
            const float4 c = row[i];
            avg += vec4(c.x, c.y, c.z, c.w) * factor;
        }
    }
    avg /= sampleCount;
    const vec4 base(vec3(avg.x, avg.y, avg.z) * brightness, 1.0f);
    const float ig = g_div(1.0f, 2.2f);
    const vec4 frag(rmx::pow_ft(base.x, ig), rmx::pow_ft(base.y, ig), rmx::pow_ft(base.z, ig), rmx::pow_ft(base.w, ig));   // :54
    unsigned char b[4];
    for (int cpt = 0; cpt < 4; cpt++) {
        float v = frag[cpt];
        if (isnan(v)) v = 0.0f;
        v = clamp(v, 0.0f, 1.0f);
        b[cpt] = (unsigned char)(int)floor(g_add(g_mul(v, 255.0f), 0.5f));
    }
    out[idx] = make_uchar4(b[0], b[1], b[2], b[3]);
    // fused tile gather (multi-GPU row-tile sharding): the same pixel goes straight into the assembled
    // full frame - usually rank 0's memory mapped over NVLink (CUDA IPC) - at its GLOBAL row.  A warp
    // writes one 128-byte line, so the peer stores are fully coalesced; no separate collective moves pixels.
    if (gather) gather[(size_t)gy * (size_t)W + (size_t)x] = make_uchar4(b[0], b[1], b[2], b[3]);
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
    dim3 block(256, 1, 1), grid((unsigned)((W + 255) / 256), (unsigned)local_rows, 1);
    xg::disp::rm_display_kernel<<<grid, block, 0, stream>>>((const float4*)color, (const ushort4*)nd, (uchar4*)rgba8, (uchar4*)gather, W,
                                                            local_rows, H, tile_rows, n_ranks, rank, brightness);
    return cudaGetLastError();
}
