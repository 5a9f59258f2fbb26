// Screen post-processing on the device: the step renderRayTracing ends with (src/main.cpp:397-398).
//
// Replaces Screen::postprocessImage / applyBloomEffect / filterLightPixels / applyKernel / boxKernel / gaussianKernel /
// addImages / clamp / reinhardToneMap / exposureToneMap / gammaCorrection and the 8-bit conversion of writeBitmapToFile
// (src/screen.cpp:40-69, 226-395).  Images are float4 in the Screen layout (row 0 = top); .w is carried along untouched.
//
// Exactness: the bloom filter is a (2f+1)^2-tap sum per pixel that the reference accumulates in a fixed order (dx outer,
// dy inner, src/screen.cpp:301-322) with out-of-image taps reading black; the kernels keep that order and the single
// rounding per operation (rt_math.cuh), so box / Gaussian bloom, the clamp and the Reinhard map are bit-identical to the
// reference.  The Gaussian weights are evaluated on the host with the same libm call the reference makes (rt_capi.cu).
// exp / pow of the exposure map and the gamma curve come from CUDA's double-precision exp / pow rounded to float, which
// agrees with glibc's expf / powf to 1 ulp.
#include "rt_kernels.h"
#include "rt_math.cuh"

namespace rtb {
namespace {

constexpr int kPostBlockX = 32, kPostBlockY = 8;

__device__ __forceinline__ float4 ld4(const float4* p) { return __ldg(p); }

// filterLightPixels (src/screen.cpp:271-283) with convertToGrayscale (385-387): dot(pixel, (0.2126, 0.7152, 0.0722))
__global__ void k_post_light(const float4* __restrict__ img, float4* __restrict__ light, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 p = img[i];
        const float brightness = xdot(mk3(p), mk3((float)0.2126, (float)0.7152, (float)0.0722));
        light[i] = brightness >= 1.0f ? p : make_float4(0.0f, 0.0f, 0.0f, p.w);
    }
}

// applyKernel (src/screen.cpp:285-299) with boxKernel (301-313) or gaussianKernel (316-322).  A block of 32 x 8 threads
// produces 32 x (8 * R) pixels from a (32 + 2f) x (8R + 2f) neighbourhood staged in shared memory (zero outside the image,
// like getPixel, 390-396).  Each thread owns R vertically adjacent pixels: walking down a tile column it feeds every loaded
// texel to the (up to R) accumulators whose window contains it, so a texel is read from shared memory once per R outputs
// while each accumulator still receives its taps in the reference's order (dx outer, dy inner).  Measured at 4K, box
// filter of size 5 (121 taps): 0.48 ms with one pixel per thread (shared-memory bandwidth), see profiles/README.md.
constexpr int kPostRows = 8;

// One texel of the column walk feeds the accumulators k in [k_lo, k_hi].  The bounds are compile-time constants at every
// call site (the loops around it are fully unrolled), so no predicated instructions are left over.
template <bool GAUSS>
__device__ __forceinline__ void post_feed(f3 (&sum)[kPostRows], const float4& p, const float (&g)[kPostRows], int k_lo, int k_hi)
{
#pragma unroll
    for (int k = 0; k < kPostRows; k++) {
        if (k < k_lo || k > k_hi)
            continue;
        if (GAUSS)
            sum[k] = xadd(sum[k], mk3(xmul(g[k], p.x), xmul(g[k], p.y), xmul(g[k], p.z)));
        else
            sum[k] = xadd(sum[k], mk3(p));
    }
}

template <bool GAUSS>
__global__ void __launch_bounds__(kPostBlockX* kPostBlockY) k_post_blur_staged(const float4* __restrict__ src, float4* __restrict__ dst, int w, int h, int f,
    const float* __restrict__ weights)
{
    constexpr int R = kPostRows;
    extern __shared__ float4 tile[];
    const int tw = kPostBlockX + 2 * f, th = kPostBlockY * R + 2 * f;
    const int x0 = blockIdx.x * kPostBlockX - f, y0 = blockIdx.y * (kPostBlockY * R) - f;
    for (int idx = threadIdx.y * kPostBlockX + threadIdx.x; idx < tw * th; idx += kPostBlockX * kPostBlockY) {
        const int gx = x0 + idx % tw, gy = y0 + idx / tw;
        tile[idx] = (gx >= 0 && gy >= 0 && gx < w && gy < h) ? ld4(&src[(size_t)gy * w + gx]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();
    const int x = blockIdx.x * kPostBlockX + threadIdx.x, y = blockIdx.y * (kPostBlockY * R) + threadIdx.y * R;
    if (x >= w || y >= h)
        return;
    const int side = 2 * f + 1;
    f3 sum[R];
#pragma unroll
    for (int k = 0; k < R; k++)
        sum[k] = mk3(0.0f, 0.0f, 0.0f);
    for (int i = -f; i < f + 1; i++) {
        const float4* col = tile + (threadIdx.y * R) * tw + (threadIdx.x + f + i);
        const float* wcol = weights + (i + f) * side;
        float g[R]; // g[k] = weight of tap jj - k: a window sliding down the weight column with the walk
#pragma unroll
        for (int k = 0; k < R; k++)
            g[k] = 0.0f;
        auto advance = [&](int jj) {
            if (GAUSS) {
#pragma unroll
                for (int k = R - 1; k > 0; k--)
                    g[k] = g[k - 1];
                g[0] = jj < side ? __ldg(&wcol[jj]) : 0.0f;
            }
        };
        if (side >= R - 1) {
            // ramp up: texel m reaches pixels 0..m; steady: every pixel; ramp down: texel side-1+m reaches pixels m..R-1
#pragma unroll
            for (int m = 0; m < R - 1; m++) {
                advance(m);
                post_feed<GAUSS>(sum, col[m * tw], g, 0, m);
            }
            for (int jj = R - 1; jj < side; jj++) {
                advance(jj);
                post_feed<GAUSS>(sum, col[jj * tw], g, 0, R - 1);
            }
#pragma unroll
            for (int m = 1; m < R; m++) {
                advance(side - 1 + m);
                post_feed<GAUSS>(sum, col[(side - 1 + m) * tw], g, m, R - 1);
            }
        } else { // filter narrower than the pixel run (f < 3): windows do not all overlap, test each pair
            for (int jj = 0; jj < side + R - 1; jj++) {
                advance(jj);
                const float4 p = col[jj * tw];
#pragma unroll
                for (int k = 0; k < R; k++)
                    if (jj - k >= 0 && jj - k < side)
                        post_feed<GAUSS>(sum, p, g, k, k);
            }
        }
    }
    const float n = (float)(side * side);
#pragma unroll
    for (int k = 0; k < R; k++) {
        if (y + k >= h)
            break;
        f3 c = sum[k];
        if (!GAUSS) // sum /= (2f+1)^2: glm converts the int to float and divides each component
            c = mk3(xdiv(c.x, n), xdiv(c.y, n), xdiv(c.z, n));
        dst[(size_t)(y + k) * w + x] = make_float4(c.x, c.y, c.z, tile[(threadIdx.y * R + k + f) * tw + threadIdx.x + f].w);
    }
}

// The same filter without staging, for filter sizes whose neighbourhood does not fit in shared memory (or negative ones,
// whose loops are empty): every tap is a bounds-checked global load.
template <bool GAUSS>
__global__ void __launch_bounds__(kPostBlockX* kPostBlockY) k_post_blur_direct(const float4* __restrict__ src, float4* __restrict__ dst, int w, int h, int f,
    const float* __restrict__ weights)
{
    const int x = blockIdx.x * kPostBlockX + threadIdx.x, y = blockIdx.y * kPostBlockY + threadIdx.y;
    if (x >= w || y >= h)
        return;
    const int side = 2 * f + 1;
    f3 sum = mk3(0.0f, 0.0f, 0.0f);
    for (int i = -f; i < f + 1; i++) {
        for (int j = -f; j < f + 1; j++) {
            const int gx = x + i, gy = y + j;
            const float4 p = (gx >= 0 && gy >= 0 && gx < w && gy < h) ? ld4(&src[(size_t)gy * w + gx]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (GAUSS) {
                const float g = __ldg(&weights[(i + f) * side + (j + f)]);
                sum = xadd(sum, mk3(xmul(g, p.x), xmul(g, p.y), xmul(g, p.z)));
            } else {
                sum = xadd(sum, mk3(p));
            }
        }
    }
    if (!GAUSS) {
        const float n = (float)(side * side);
        sum = mk3(xdiv(sum.x, n), xdiv(sum.y, n), xdiv(sum.z, n));
    }
    const size_t o = (size_t)y * w + x;
    dst[o] = make_float4(sum.x, sum.y, sum.z, ld4(&src[o]).w);
}

__device__ __forceinline__ float clamp01(float v)
{
    const float lo = (v < 0.0f) ? 0.0f : v; // glm::max(v, 0) = (v < 0) ? 0 : v
    return (1.0f < lo) ? 1.0f : lo;          // glm::min(., 1) = (1 < .) ? 1 : .
}

// addImages + tone mapping (src/screen.cpp:252-267, 340-383).  option: rt_b200.h RT_FILTER_*
__global__ void k_post_combine(float4* img, const float4* __restrict__ light, size_t n, int option, float exposure)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = img[i], b = light[i];
        f3 c = xadd(mk3(a), mk3(b));
        if (option == 1) {
            c = mk3(clamp01(c.x), clamp01(c.y), clamp01(c.z));
        } else if (option == 2) { // pixel / (pixel + 1)
            c = mk3(xdiv(c.x, xadd(c.x, 1.0f)), xdiv(c.y, xadd(c.y, 1.0f)), xdiv(c.z, xadd(c.z, 1.0f)));
        } else if (option == 3) { // 1 - exp(-pixel * exposure)
            const f3 e = xmul(xneg(c), exposure);
            c = mk3(xsub(1.0f, (float)exp((double)e.x)), xsub(1.0f, (float)exp((double)e.y)), xsub(1.0f, (float)exp((double)e.z)));
        }
        img[i] = make_float4(c.x, c.y, c.z, a.w);
    }
}

// gammaCorrection (src/screen.cpp:380-382): pow(pixel, 1 / gamma) per component
__global__ void k_post_gamma(float4* img, size_t n, float e)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = img[i];
        img[i] = make_float4((float)pow((double)a.x, (double)e), (float)pow((double)a.y, (double)e), (float)pow((double)a.z, (double)e), a.w);
    }
}

// writeBitmapToFile's conversion (src/screen.cpp:44-49): clamp to [0, 1], times 255, truncate; alpha 255
__global__ void k_post_rgba8(const float4* __restrict__ img, uchar4* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = img[i];
        out[i] = make_uchar4((unsigned char)__float2int_rz(xmul(clamp01(a.x), 255.0f)), (unsigned char)__float2int_rz(xmul(clamp01(a.y), 255.0f)),
            (unsigned char)__float2int_rz(xmul(clamp01(a.z), 255.0f)), 255);
    }
}

__global__ void k_unpack_rgb(const float* __restrict__ in, float4* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 1.0f);
}

int stream_grid(size_t n, int sm_count) { return (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count * 8); }

} // namespace

void launch_post_light(cudaStream_t st, int sm_count, const float4* img, float4* light, size_t n)
{
    if (n)
        k_post_light<<<stream_grid(n, sm_count), 256, 0, st>>>(img, light, n);
}

constexpr int kPostMaxStagedFilter = 16; // (32 + 32) x (64 + 32) texels = 96 KB of shared memory

void launch_post_blur(cudaStream_t st, const float4* src, float4* dst, int w, int h, int f, bool gauss, const float* weights)
{
    if (w <= 0 || h <= 0)
        return;
    const dim3 block(kPostBlockX, kPostBlockY);
    if (f >= 0 && f <= kPostMaxStagedFilter) {
        const dim3 grid((w + kPostBlockX - 1) / kPostBlockX, (h + kPostBlockY * kPostRows - 1) / (kPostBlockY * kPostRows));
        const size_t smem = (size_t)(kPostBlockX + 2 * f) * (kPostBlockY * kPostRows + 2 * f) * sizeof(float4);
        static bool opted_in = false;
        if (!opted_in) {
            const int max_smem = (kPostBlockX + 2 * kPostMaxStagedFilter) * (kPostBlockY * kPostRows + 2 * kPostMaxStagedFilter) * (int)sizeof(float4);
            cudaFuncSetAttribute(k_post_blur_staged<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
            cudaFuncSetAttribute(k_post_blur_staged<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
            opted_in = true;
        }
        if (gauss)
            k_post_blur_staged<true><<<grid, block, smem, st>>>(src, dst, w, h, f, weights);
        else
            k_post_blur_staged<false><<<grid, block, smem, st>>>(src, dst, w, h, f, weights);
    } else {
        const dim3 grid((w + kPostBlockX - 1) / kPostBlockX, (h + kPostBlockY - 1) / kPostBlockY);
        if (gauss)
            k_post_blur_direct<true><<<grid, block, 0, st>>>(src, dst, w, h, f, weights);
        else
            k_post_blur_direct<false><<<grid, block, 0, st>>>(src, dst, w, h, f, weights);
    }
}

void launch_post_combine(cudaStream_t st, int sm_count, float4* img, const float4* light, size_t n, int option, float exposure)
{
    if (n)
        k_post_combine<<<stream_grid(n, sm_count), 256, 0, st>>>(img, light, n, option, exposure);
}

void launch_post_gamma(cudaStream_t st, int sm_count, float4* img, size_t n, float e)
{
    if (n)
        k_post_gamma<<<stream_grid(n, sm_count), 256, 0, st>>>(img, n, e);
}

void launch_post_rgba8(cudaStream_t st, int sm_count, const float4* img, unsigned char* out, size_t n)
{
    if (n)
        k_post_rgba8<<<stream_grid(n, sm_count), 256, 0, st>>>(img, reinterpret_cast<uchar4*>(out), n);
}

void launch_unpack_rgb(cudaStream_t st, int sm_count, const float* in, float4* out, size_t n)
{
    if (n)
        k_unpack_rgb<<<stream_grid(n, sm_count), 256, 0, st>>>(in, out, n);
}

} // namespace rtb
