// HBM-bound companions of the tensor-core kernels: GroupNorm statistics / apply (+SiLU, +resample),
// row softmax, timestep embedding, small dense layers, input preparation.  All NHWC, 128-bit accesses.
#include "ops.h"
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdlib.h>

namespace dsb {

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ void split_h16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// ------------------------------------------------------------------------------------------ GN stats
// grid (chunks, B); block (ncol4 <= 384, rows).  Thread (tx, ty) owns float4 column tx.
__global__ void gn_stats_kernel(ds_gn_stats_desc d, int pix_per_cta) {
    __shared__ double s_sum[64];
    __shared__ double s_sq[64];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < 64) { s_sum[tid] = 0.0; s_sq[tid] = 0.0; }
    __syncthreads();
    const int C = d.C0 + d.C1;
    const int cpg = C / d.groups;
    const int n = blockIdx.y;
    const int ncol4 = C / 4;
    const int p_begin = blockIdx.x * pix_per_cta;
    int p_end = p_begin + pix_per_cta;
    if (p_end > d.HW) p_end = d.HW;
    for (int col = threadIdx.x; col < ncol4; col += blockDim.x) {
        const int c = col * 4;
        const float* base;
        int pitch, cc;
        if (c < d.C0) { base = d.src0; pitch = d.C0; cc = c; }
        else { base = d.src1; pitch = d.C1; cc = c - d.C0; }
        const int gA = c / cpg;
        const int gB = (c + 3) / cpg;
        const int nA = (gA == gB) ? 4 : ((gA + 1) * cpg - c);   // channels of this float4 that belong to group A
        float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
#pragma unroll 8
        for (int p = p_begin + threadIdx.y; p < p_end; p += blockDim.y) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(base + ((long long)n * d.HW + p) * pitch + cc));
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nA) { sA += e[j]; qA += e[j] * e[j]; }
                else { sB += e[j]; qB += e[j] * e[j]; }
            }
        }
        atomicAdd(&s_sum[gA], (double)sA);
        atomicAdd(&s_sq[gA], (double)qA);
        if (gB != gA) {
            atomicAdd(&s_sum[gB], (double)sB);
            atomicAdd(&s_sq[gB], (double)qB);
        }
    }
    __syncthreads();
    if (tid < d.groups) {
        atomicAdd(&d.sums[((long long)n * d.groups + tid) * 2 + 0], s_sum[tid]);
        atomicAdd(&d.sums[((long long)n * d.groups + tid) * 2 + 1], s_sq[tid]);
    }
}

// ------------------------------------------------------------------------------------------ GN statistics from quad partials
// One CTA per sample.  Step 1: thread t owns the (quad, sum|sumsq) column t of the concatenated partial row (C/2 columns) and adds
// it over the sample's slabs in fp64 (coalesced along t).  Step 2: thread j < 2*groups adds the cpg/4 quads of its group.
__global__ void __launch_bounds__(1024) gn_finalize_kernel(ds_gn_finalize_desc d) {
    extern __shared__ double s_cols[];                    // [cols] column sums, then [SG][cols] partials
    __shared__ float s_mu[64], s_rstd[64];
    const int n = blockIdx.x;
    const int C = d.C0 + d.C1;
    if (d.quads0) {
        // partial rows: {sum, sumsq} per unit of u0 / u1 channels (4 = quads, 2 = pairs) -> w floats per 32-row slab.
        // Thread (column, slab group): the 1024 threads split the sample's slabs SG ways (round 2: one thread per column walked all
        // HW / 32 slabs alone, 35 us per launch at 64 x 64); the SG partials of a column are added in a fixed order (deterministic).
        const int u0 = d.unit0 == 2 ? 2 : 4, u1 = d.unit1 == 2 ? 2 : 4;
        const int w0 = d.C0 / u0 * 2, w1 = d.C1 / u1 * 2;
        const int cols = w0 + w1;
        int SG = blockDim.x / cols;
        if (SG < 1) SG = 1;
        if (SG > d.slabs_per_sample) SG = d.slabs_per_sample;
        double* part = s_cols + cols;
        for (int idx = threadIdx.x; idx < cols * SG; idx += blockDim.x) {
            const int t = idx % cols, sg = idx / cols;
            const float* src = (t < w0) ? d.quads0 + (long long)n * d.slabs_per_sample * w0 + t
                                        : d.quads1 + (long long)n * d.slabs_per_sample * w1 + (t - w0);
            const int pitch = (t < w0) ? w0 : w1;
            double acc = 0.0;
#pragma unroll 4
            for (int sl = sg; sl < d.slabs_per_sample; sl += SG) acc += (double)__ldg(src + (long long)sl * pitch);
            part[(long long)sg * cols + t] = acc;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < cols; t += blockDim.x) {
            double acc = 0.0;
            for (int sg = 0; sg < SG; ++sg) acc += part[(long long)sg * cols + t];
            s_cols[t] = acc;
        }
        __syncthreads();
        const int cpg = C / d.groups;
        for (int j = threadIdx.x; j < 2 * d.groups; j += blockDim.x) {
            const int g = j >> 1, k = j & 1;
            // channels [g * cpg, (g + 1) * cpg): whole units of source 0, then of source 1 (group boundaries fall on unit boundaries,
            // and C0 is a multiple of cpg or the group straddles the two sources at a unit boundary of each)
            double acc = 0.0;
            int c = g * cpg;
            const int c_end = c + cpg;
            while (c < c_end) {
                if (c < d.C0) { acc += s_cols[(c / u0) * 2 + k]; c += u0; }
                else { acc += s_cols[w0 + ((c - d.C0) / u1) * 2 + k]; c += u1; }
            }
            d.sums[((long long)n * d.groups + g) * 2 + k] = acc;
        }
        __syncthreads();
    }
    if (!d.coef) return;
    // second product: y = x * a + b per (sample, channel), so that gn_apply starts streaming without an fp64 prologue per thread
    const double cnt = (double)(C / d.groups) * d.HW;
    for (int g = threadIdx.x; g < d.groups; g += blockDim.x) {
        const double s = d.sums[((long long)n * d.groups + g) * 2 + 0];
        const double q = d.sums[((long long)n * d.groups + g) * 2 + 1];
        const double mu = s / cnt;
        double var = q / cnt - mu * mu;
        if (var < 0.0) var = 0.0;
        s_mu[g] = (float)mu;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)d.eps));
    }
    __syncthreads();
    const int cpg = C / d.groups;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float aa = s_rstd[g] * __ldg(d.gamma + c);
        float bb = __ldg(d.beta + c);
        if (d.ada) {
            const float sc = d.ada[(long long)n * d.ada_stride + c] + 1.0f;
            const float sh = d.ada[(long long)n * d.ada_stride + C + c];
            aa *= sc;
            bb = bb * sc + sh;
        }
        reinterpret_cast<float2*>(d.coef)[(long long)n * C + c] = make_float2(aa, fmaf(-s_mu[g], aa, bb));
    }
}

// ------------------------------------------------------------------------------------------ GN apply
// grid (chunks, B); block = nc8 * rows threads.  Thread (c8, prow) owns 8 fixed channels: its normalisation coefficients
// live in registers (mean, a = rstd*gamma*(1+ada_scale), b = beta*(1+ada_scale)+ada_shift) and it streams over output pixels.
// fmt 1 (operand of an f8 GEMM, csrc/ops.h): fp16 plane of v * 2^A16 (saturating) followed by the two e4m3 byte planes
// (v - hi) * 2^LO8 and hi * 2^HI8, where hi is the value the fp16 plane represents.  Powers of two: the roundings are those of v.
// Packed arithmetic (two values per instruction wherever the ISA has it): v * 2^A16 -> f16x2 convert -> clamp as half2 (a value beyond the fp16
// range converts to inf and is clamped back: the same result as clamping first) -> hi byte plane straight from the half2
// (cvt.e4m3x2.f16x2 of hi * 2^(HI8 - A16), exact: a power of two) -> lo = fma(hi, -2^(LO8 - A16), v * 2^LO8) = (v - hi / 2^A16) * 2^LO8 with
// one rounding, as before.  7 instructions per value instead of 11 (gn_apply with this store was issue-bound: 41 instructions per element
// with both outputs, profiles/r02/ncu_gn_apply_v3_cifar_r02m.txt).
__device__ __forceinline__ void f8_image_pair(float v0, float v1, uint32_t& hi16, unsigned short& lo8, unsigned short& hi8) {
    constexpr float kA16 = (float)(1 << DS_F8_SH_A16), kLo8 = (float)(1 << DS_F8_SH_LO8);
    constexpr float kLoA = (float)(1 << (DS_F8_SH_LO8 - DS_F8_SH_A16));
    static_assert(DS_F8_SH_A16 >= DS_F8_SH_HI8 && DS_F8_SH_LO8 >= DS_F8_SH_A16, "operand scales");
    const __half2 lim = __float2half2_rn(65504.f);
    __half2 h = __floats2half2_rn(v0 * kA16, v1 * kA16);
    h = __hmin2(__hmax2(h, __hneg2(lim)), lim);
    hi16 = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = __half22float2(h);
    lo8 = __nv_cvt_float2_to_fp8x2(make_float2(fmaf(hf.x, -kLoA, v0 * kLo8), fmaf(hf.y, -kLoA, v1 * kLo8)), __NV_SATFINITE, __NV_E4M3);
    const __half2 h8 = __hmul2(h, __float2half2_rn(1.0f / (float)(1 << (DS_F8_SH_A16 - DS_F8_SH_HI8))));
    hi8 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&h8), __NV_SATFINITE, __NV_E4M3);
}

__device__ __forceinline__ void gn_store_f8(__half* base, long long plane, long long o, const float* v) {
    __align__(16) uint32_t hi[4];
    __align__(8) unsigned short lo8[4];
    __align__(8) unsigned short hi8[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) f8_image_pair(v[2 * j], v[2 * j + 1], hi[j], lo8[j], hi8[j]);
    *reinterpret_cast<uint4*>(base + o) = *reinterpret_cast<const uint4*>(hi);
    unsigned char* b8 = reinterpret_cast<unsigned char*>(base + plane);
    *reinterpret_cast<uint2*>(b8 + o) = *reinterpret_cast<const uint2*>(lo8);
    *reinterpret_cast<uint2*>(b8 + plane + o) = *reinterpret_cast<const uint2*>(hi8);
}

// fp16 hi / lo planes of two values: one packed convert each way (hi = rn(v), lo = rn(v - hi), as split_h16)
__device__ __forceinline__ void split_h16_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void gn_store_planes(__half* base, long long plane, long long o, const float* v, int nplanes, int fmt = 0) {
    if (fmt == 1) { gn_store_f8(base, plane, o, v); return; }
    __align__(16) uint32_t hi[4];
    __align__(16) uint32_t lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_h16_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    *reinterpret_cast<uint4*>(base + o) = *reinterpret_cast<const uint4*>(hi);
    if (nplanes > 1) *reinterpret_cast<uint4*>(base + plane + o) = *reinterpret_cast<const uint4*>(lo);
}

template <int RESAMPLE>
__global__ void __launch_bounds__(512) gn_apply_kernel(ds_gn_apply_desc d, int pix_per_cta, int nc8, int rows) {
    const int C = d.C0 + d.C1;
    const int n = blockIdx.y;
    const int c8 = threadIdx.x % nc8;
    const int prow = threadIdx.x / nc8;
    const int c = c8 * 8;
    const bool norm = d.sums != nullptr;
    float mean[8], a[8], b[8];
    if (norm) {
        const int cpg = C / d.groups;
        const double cnt = (double)cpg * d.H * d.W;
        int g_prev = -1;
        float mu_f = 0.f, rstd = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c + j) / cpg;
            if (g != g_prev) {                 // at most a few distinct groups per 8 channels
                const double s = d.sums[((long long)n * d.groups + g) * 2 + 0];
                const double q = d.sums[((long long)n * d.groups + g) * 2 + 1];
                const double mu = s / cnt;
                double var = q / cnt - mu * mu;
                if (var < 0.0) var = 0.0;
                rstd = (float)(1.0 / sqrt(var + (double)d.eps));
                mu_f = (float)mu;
                g_prev = g;
            }
            float aa = rstd * __ldg(d.gamma + c + j);
            float bb = __ldg(d.beta + c + j);
            if (d.ada) {
                const float sc = d.ada[(long long)n * d.ada_stride + c + j] + 1.0f;
                const float sh = d.ada[(long long)n * d.ada_stride + C + c + j];
                aa *= sc;
                bb = bb * sc + sh;
            }
            mean[j] = mu_f; a[j] = aa; b[j] = bb;
        }
    }
    const int Ho = RESAMPLE == 1 ? d.H / 2 : (RESAMPLE == 2 ? d.H * 2 : d.H);
    const int Wo = RESAMPLE == 1 ? d.W / 2 : (RESAMPLE == 2 ? d.W * 2 : d.W);
    const int npix = Ho * Wo;              // RESAMPLE == 3 iterates over INPUT pixels and scatters them into the phase layout
    const long long plane = (long long)d.B * npix * C;
    const float* base;
    int pitch, cc;
    if (c < d.C0) { base = d.src0; pitch = d.C0; cc = c; }
    else { base = d.src1; pitch = d.C1; cc = c - d.C0; }
    base += (long long)n * d.H * d.W * pitch + cc;
    __half* oact = reinterpret_cast<__half*>(d.out_act);
    __half* oraw = reinterpret_cast<__half*>(d.out_raw);
    const int p_begin = blockIdx.x * pix_per_cta;
    int p_end = p_begin + pix_per_cta;
    if (p_end > npix) p_end = npix;
    int p_begin_tail = p_begin + prow;
    if (RESAMPLE == 0) {
        // plain path (the common case): two pixels per iteration so that four 16-byte loads are in flight per thread
        int po = p_begin + prow;
        for (; po + rows < p_end; po += 2 * rows) {
            const float* s0 = base + (long long)po * pitch;
            const float* s1 = base + (long long)(po + rows) * pitch;
            const float4 a0 = __ldcs(reinterpret_cast<const float4*>(s0));
            const float4 a1 = __ldcs(reinterpret_cast<const float4*>(s0) + 1);
            const float4 b0 = __ldcs(reinterpret_cast<const float4*>(s1));
            const float4 b1 = __ldcs(reinterpret_cast<const float4*>(s1) + 1);
            const float ea[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float eb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float ya[8], yb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float u = 0.f, w = 0.f;
                if (norm) {
                    u = (ea[j] - mean[j]) * a[j] + b[j];
                    w = (eb[j] - mean[j]) * a[j] + b[j];
                    if (d.silu) { u = silu_f(u); w = silu_f(w); }
                }
                ya[j] = u; yb[j] = w;
            }
            const long long oa = ((long long)n * npix + po) * C + c;
            const long long ob = ((long long)n * npix + po + rows) * C + c;
            if (oact) { gn_store_planes(oact, plane, oa, ya, d.nplanes, d.fmt); gn_store_planes(oact, plane, ob, yb, d.nplanes, d.fmt); }
            if (oraw) { gn_store_planes(oraw, plane, oa, ea, d.nplanes, d.fmt); gn_store_planes(oraw, plane, ob, eb, d.nplanes, d.fmt); }
            if (d.out_raw_f32) {
                *reinterpret_cast<float4*>(d.out_raw_f32 + oa) = a0; *reinterpret_cast<float4*>(d.out_raw_f32 + oa + 4) = a1;
                *reinterpret_cast<float4*>(d.out_raw_f32 + ob) = b0; *reinterpret_cast<float4*>(d.out_raw_f32 + ob + 4) = b1;
            }
        }
        p_begin_tail = po;
    }
    for (int po = (RESAMPLE == 0 ? p_begin_tail : p_begin + prow); po < p_end; po += rows) {
        const int ho = po / Wo, wo = po - ho * Wo;
        float act[8], raw[8];
        if (RESAMPLE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { act[j] = 0.f; raw[j] = 0.f; }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float* src = base + (long long)((ho * 2 + (t >> 1)) * d.W + wo * 2 + (t & 1)) * pitch;
                const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src));
                const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
                const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    raw[j] += 0.25f * e[j];
                    if (norm) {
                        float y = (e[j] - mean[j]) * a[j] + b[j];
                        if (d.silu) y = silu_f(y);
                        act[j] += 0.25f * y;
                    }
                }
            }
        } else {
            const int hi_ = RESAMPLE == 2 ? (ho >> 1) : ho;
            const int wi_ = RESAMPLE == 2 ? (wo >> 1) : wo;
            const float* src = base + (long long)(hi_ * d.W + wi_) * pitch;
            const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src));
            const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
            const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                raw[j] = e[j];
                float y = 0.f;
                if (norm) {
                    y = (e[j] - mean[j]) * a[j] + b[j];
                    if (d.silu) y = silu_f(y);
                }
                act[j] = y;
            }
        }
        long long o = ((long long)n * npix + po) * C + c;
        if (RESAMPLE == 3) {
            // space-to-depth: input pixel (ho, wo) -> output pixel (ho/2, wo/2), channel block ((ho&1)*2 + (wo&1))*C
            const int h2 = ho >> 1, w2 = wo >> 1, ph = ((ho & 1) << 1) | (wo & 1);
            o = (((long long)n * (d.H / 2) + h2) * (d.W / 2) + w2) * (4LL * C) + (long long)ph * C + c;
        }
        if (oact) gn_store_planes(oact, plane, o, act, d.nplanes, d.fmt);
        if (oraw) gn_store_planes(oraw, plane, o, raw, d.nplanes, d.fmt);
        if (d.out_raw_f32) {
            *reinterpret_cast<float4*>(d.out_raw_f32 + o) = make_float4(raw[0], raw[1], raw[2], raw[3]);
            *reinterpret_cast<float4*>(d.out_raw_f32 + o + 4) = make_float4(raw[4], raw[5], raw[6], raw[7]);
        }
    }
}

// The plain (RESAMPLE == 0) path, default since round 1 (DSB_GN_APPLY_V2=0 selects the older loop inside gn_apply_kernel<0> for A/B): the
// normalisation is folded to one FMA per element, y = x * a + b' with b' = b - mean * a (16 instead of 24 live coefficient registers), and
// four pixels are processed per iteration so that eight 16-byte loads are in flight per thread (the two-pixel loop keeps ~49 KB per SM in
// flight, about the minimum HBM needs).  Measured on the CIFAR-10 forward at batch 512: 10.7 -> 9.55 ms (profiles/r01d).
template <int NP>
__device__ __forceinline__ void gn_v2_pixels(const ds_gn_apply_desc& d, const float* base, int pitch, long long n, int npix, int C, int c,
                                             long long plane, int po, int rows, bool norm, const float* a, const float* b, __half* oact,
                                             __half* oraw) {
    float4 v[NP][2];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const float4* s = reinterpret_cast<const float4*>(base + (long long)(po + k * rows) * pitch);
        v[k][0] = __ldcs(s);
        v[k][1] = __ldcs(s + 1);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const float e[8] = {v[k][0].x, v[k][0].y, v[k][0].z, v[k][0].w, v[k][1].x, v[k][1].y, v[k][1].z, v[k][1].w};
        const long long o = ((long long)n * npix + po + k * rows) * C + c;
        if (oact) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float u = 0.f;
                if (norm) {
                    u = fmaf(e[j], a[j], b[j]);
                    if (d.silu) u = silu_f(u);
                }
                y[j] = u;
            }
            gn_store_planes(oact, plane, o, y, d.nplanes, d.fmt);
        }
        if (oraw) gn_store_planes(oraw, plane, o, e, d.nplanes, d.fmt);
        if (d.out_raw_f32) {
            *reinterpret_cast<float4*>(d.out_raw_f32 + o) = v[k][0];
            *reinterpret_cast<float4*>(d.out_raw_f32 + o + 4) = v[k][1];
        }
    }
}

__global__ void __launch_bounds__(512) gn_apply_v2_kernel(ds_gn_apply_desc d, int pix_per_cta, int nc8, int rows) {
    const int C = d.C0 + d.C1;
    const int n = blockIdx.y;
    const int c8 = threadIdx.x % nc8;
    const int prow = threadIdx.x / nc8;
    const int c = c8 * 8;
    const bool norm = d.sums != nullptr;
    float a[8], b[8];
    if (norm) {
        const int cpg = C / d.groups;
        const double cnt = (double)cpg * d.H * d.W;
        int g_prev = -1;
        float mu_f = 0.f, rstd = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c + j) / cpg;
            if (g != g_prev) {
                const double s = d.sums[((long long)n * d.groups + g) * 2 + 0];
                const double q = d.sums[((long long)n * d.groups + g) * 2 + 1];
                const double mu = s / cnt;
                double var = q / cnt - mu * mu;
                if (var < 0.0) var = 0.0;
                rstd = (float)(1.0 / sqrt(var + (double)d.eps));
                mu_f = (float)mu;
                g_prev = g;
            }
            float aa = rstd * __ldg(d.gamma + c + j);
            float bb = __ldg(d.beta + c + j);
            if (d.ada) {
                const float sc = d.ada[(long long)n * d.ada_stride + c + j] + 1.0f;
                const float sh = d.ada[(long long)n * d.ada_stride + C + c + j];
                aa *= sc;
                bb = bb * sc + sh;
            }
            a[j] = aa;
            b[j] = fmaf(-mu_f, aa, bb);
        }
    }
    const int npix = d.H * d.W;
    const long long plane = (long long)d.B * npix * C;
    const float* base;
    int pitch, cc;
    if (c < d.C0) { base = d.src0; pitch = d.C0; cc = c; }
    else { base = d.src1; pitch = d.C1; cc = c - d.C0; }
    base += (long long)n * npix * pitch + cc;
    __half* oact = reinterpret_cast<__half*>(d.out_act);
    __half* oraw = reinterpret_cast<__half*>(d.out_raw);
    const int p_begin = blockIdx.x * pix_per_cta;
    int p_end = p_begin + pix_per_cta;
    if (p_end > npix) p_end = npix;
    int po = p_begin + prow;
    for (; po + 3 * rows < p_end; po += 4 * rows) gn_v2_pixels<4>(d, base, pitch, n, npix, C, c, plane, po, rows, norm, a, b, oact, oraw);
    for (; po < p_end; po += rows) gn_v2_pixels<1>(d, base, pitch, n, npix, C, c, plane, po, rows, norm, a, b, oact, oraw);
}

// Round 2 (resample == 0 with precomputed coefficients, ds_gn_apply_desc.coef): the v2 inner loop inside a PERSISTENT, EVENLY SPLIT grid.
// v2 launched one CTA per (sample, 16-pixels-per-thread chunk): 4096 CTAs over 148 x 8 slots = 3.46 waves (a 13 % tail), each thread
// paying an fp64 mean / rsqrt prologue for 16 pixels of work -- 76 % of the copy bandwidth.  Here the (sample, pixel
// row) space is cut into gridDim.x equal ranges (+-1 row), the grid is exactly the number of co-resident CTAs, and a thread fetches its
// 16 coefficients (4 x 16 B, L2-resident table written by gn_finalize) only when its range crosses into another sample.
__global__ void __launch_bounds__(256, 3) gn_apply_v3_kernel(ds_gn_apply_desc d, int nc8, int rows, int units_per_sample, long long total_units) {
    const int C = d.C0 + d.C1;
    const int c8 = threadIdx.x % nc8;
    const int prow = threadIdx.x / nc8;
    const int c = c8 * 8;
    const int npix = d.H * d.W;
    const long long plane = (long long)d.B * npix * C;
    const float* base0;
    int pitch, cc;
    if (c < d.C0) { base0 = d.src0; pitch = d.C0; cc = c; }
    else { base0 = d.src1; pitch = d.C1; cc = c - d.C0; }
    __half* oact = reinterpret_cast<__half*>(d.out_act);
    __half* oraw = reinterpret_cast<__half*>(d.out_raw);
    // r02s: with one contiguous range per CTA (round-2 v3) the 444 CTAs streamed through 444 x (2 sources + up to 5 output planes) distant
    // 2 MB pages at once and the 3.2 GB concat layers (512 channels, 32x32, act + raw outputs) ran at 2.8 TB/s against 5.3 TB/s for the
    // 1 GB layers.  Now the grid sweeps the tensor together: chunks of kChunk units (>= 8 pixel rows, never crossing a sample) are dealt
    // round-robin, so at any time all CTAs work inside one ~30 MB window per stream; the coefficients are refetched only when the sample
    // changes.
    constexpr int kChunk = 8;
    const int chunks_per_sample = (units_per_sample + kChunk - 1) / kChunk;
    const long long total_chunks = (long long)chunks_per_sample * d.B;
    float a[8], b[8];
    int n_prev = -1;
    for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
        const int n = (int)(ch / chunks_per_sample);
        const int uin = ((int)(ch - (long long)n * chunks_per_sample)) * kChunk;
        const int nun = min(kChunk, units_per_sample - uin);
        if (n != n_prev) {
            const float4* cf = reinterpret_cast<const float4*>(d.coef + ((long long)n * C + c) * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = __ldg(cf + j);
                a[2 * j] = t.x; b[2 * j] = t.y; a[2 * j + 1] = t.z; b[2 * j + 1] = t.w;
            }
            n_prev = n;
        }
        const float* base = base0 + (long long)n * npix * pitch + cc;
        int po = uin * rows + prow;                      // pixel of this thread in the first unit of the chunk
        int k = 0;
        // whole units only: the last unit of a sample may be partial when rows does not divide H*W
        const int full = ((uin + nun) * rows <= npix) ? nun : nun - 1;
        for (; k + 4 <= full; k += 4, po += 4 * rows) gn_v2_pixels<4>(d, base, pitch, n, npix, C, c, plane, po, rows, true, a, b, oact, oraw);
        for (; k < full; ++k, po += rows) gn_v2_pixels<1>(d, base, pitch, n, npix, C, c, plane, po, rows, true, a, b, oact, oraw);
        if (k < nun && po < npix) gn_v2_pixels<1>(d, base, pitch, n, npix, C, c, plane, po, rows, true, a, b, oact, oraw);
    }
}

// ------------------------------------------------------------------------------------------ softmax
// Single-read softmax: the row is held in registers (NV float4 per thread), one HBM read + one write per score.
//   ROWS_PER_CTA = 8 warps, one warp per row  (L <= 32*4*NV)         -- short rows
//   ROWS_PER_CTA = 1, 256 threads per row      (L <= 256*4*NV)        -- long rows (4096 keys of the 64x64 SD self-attention)
template <int NV, bool CTA_ROW>
__global__ void __launch_bounds__(256) softmax_reg_kernel(ds_softmax_desc d) {
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = CTA_ROW ? (long long)blockIdx.x : (long long)blockIdx.x * 8 + warp;
    if (row >= d.rows) return;                      // whole warp (or CTA) exits together
    const int tid = CTA_ROW ? threadIdx.x : lane;
    const int nthr = CTA_ROW ? 256 : 32;
    const float* s = d.S + row * d.L;
    float4 v[NV];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int j = (tid + k * nthr) * 4;
        if (j < d.L) {
            v[k] = __ldcs(reinterpret_cast<const float4*>(s + j));
            m = fmaxf(m, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (CTA_ROW) {
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        __syncthreads();
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int j = (tid + k * nthr) * 4;
        if (j < d.L) {
            v[k].x = expf(v[k].x - m); v[k].y = expf(v[k].y - m); v[k].z = expf(v[k].z - m); v[k].w = expf(v[k].w - m);
            sum += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (CTA_ROW) {
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w];
    }
    const float inv = 1.0f / sum;
    __half* P = reinterpret_cast<__half*>(d.P);
    const long long plane = d.rows * d.L;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int j = (tid + k * nthr) * 4;
        if (j < d.L) {
            const float e[4] = {v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv};
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_h16(e[q], hi[q], lo[q]);
            *reinterpret_cast<uint2*>(P + row * d.L + j) = *reinterpret_cast<const uint2*>(hi);
            if (d.nplanes > 1) *reinterpret_cast<uint2*>(P + plane + row * d.L + j) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// one warp per row
__global__ void softmax_kernel(ds_softmax_desc d) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= d.rows) return;
    const int lane = threadIdx.x & 31;
    const int pin = d.pitch_in ? d.pitch_in : d.L;
    const int pout = d.pitch_out ? d.pitch_out : d.L;
    const float* s = d.S + row * pin;
    __half* P = reinterpret_cast<__half*>(d.P);
    const long long plane = d.rows * pout;
    if ((d.L & 3) == 0 && (pin & 3) == 0 && (pout & 3) == 0) {
        float m = -INFINITY;
        for (int j = lane * 4; j < d.L; j += 128) {
            const float4 v = *reinterpret_cast<const float4*>(s + j);
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int j = lane * 4; j < d.L; j += 128) {
            const float4 v = *reinterpret_cast<const float4*>(s + j);
            sum += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        for (int j = lane * 4; j < d.L; j += 128) {
            const float4 v = *reinterpret_cast<const float4*>(s + j);
            const float e[4] = {expf(v.x - m) * inv, expf(v.y - m) * inv, expf(v.z - m) * inv, expf(v.w - m) * inv};
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_h16(e[k], hi[k], lo[k]);
            *reinterpret_cast<uint2*>(P + row * pout + j) = *reinterpret_cast<const uint2*>(hi);
            if (d.nplanes > 1) *reinterpret_cast<uint2*>(P + plane + row * pout + j) = *reinterpret_cast<const uint2*>(lo);
        }
        return;
    }
    // generic row length (e.g. 77 context tokens): scalar accesses
    float m = -INFINITY;
    for (int j = lane; j < d.L; j += 32) m = fmaxf(m, s[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < d.L; j += 32) sum += expf(s[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < d.L; j += 32) {
        __half hi, lo;
        split_h16(expf(s[j] - m) * inv, hi, lo);
        P[row * pout + j] = hi;
        if (d.nplanes > 1) P[plane + row * pout + j] = lo;
    }
}

// ------------------------------------------------------------------------------------------ embedding
__global__ void posemb_kernel(ds_posemb_desc d) {
    const int n = blockIdx.x;
    if (d.mode == 1) {
        // LDM timestep_embedding (util.py:151-171): freqs = exp(-ln(10000) * i / half), [cos | sin]
        const float t = d.sigma[n];
        const int half = d.num_channels / 2;
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const float freq = expf(-logf(10000.0f) * (float)i / (float)half);
            const float a = t * freq;
            d.emb[n * d.num_channels + i] = cosf(a);
            d.emb[n * d.num_channels + half + i] = sinf(a);
        }
        return;
    }
    const float sigma = d.sigma[n];
    const float sd = d.sigma_data;
    const float s2 = sigma * sigma + sd * sd;
    const float c_noise = logf(sigma) / 4.0f;
    if (threadIdx.x == 0) {
        d.coef[n * 4 + 0] = sd * sd / s2;
        d.coef[n * 4 + 1] = sigma * sd / sqrtf(s2);
        d.coef[n * 4 + 2] = 1.0f / sqrtf(s2);
        d.coef[n * 4 + 3] = c_noise;
    }
    const int half = d.num_channels / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        // freqs = (1/10000) ** (i / (half - endpoint))   (networks_edm.py:193-195)
        const float fr = (float)i / (float)(half - (d.endpoint ? 1 : 0));
        const float freq = powf(1.0f / 10000.0f, fr);
        const float a = c_noise * freq;
        const float cs = cosf(a), sn = sinf(a);
        // reference layout is [cos | sin]; SongUNet then swaps the halves to [sin | cos]
        if (d.swap_sincos) { d.emb[n * d.num_channels + i] = sn; d.emb[n * d.num_channels + half + i] = cs; }
        else { d.emb[n * d.num_channels + i] = cs; d.emb[n * d.num_channels + half + i] = sn; }
    }
}

// one warp per output feature; loops over rows re-using the weight row held in registers
template <int MAXF>
__global__ void linear_kernel(ds_linear_desc d) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= d.out_f) return;
    const int lane = threadIdx.x & 31;
    float w[MAXF / 32];
#pragma unroll
    for (int k = 0; k < MAXF / 32; ++k) {
        const int i = lane + 32 * k;
        w[k] = (i < d.in_f) ? d.W[(long long)o * d.in_f + i] : 0.f;
    }
    const float bias = d.b ? d.b[o] : 0.f;
    for (int n = blockIdx.y; n < d.n_rows; n += gridDim.y) {
        const float* x = d.in + (long long)n * d.in_stride;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < MAXF / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < d.in_f) acc += w[k] * (d.in_scale * x[i]);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = acc + bias;
            if (d.add) v += d.add[(long long)n * d.add_stride + o];
            if (d.act == 1) v = silu_f(v);
            d.out[(long long)n * d.out_f + o] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ input prep
__global__ void prep_input_kernel(ds_prep_input_desc d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (n, pixel, 8-ch group)
    const long long total = (long long)d.B * d.HW * 8;
    if (idx >= total) return;
    const int c8 = (int)(idx & 7);
    const long long px = idx >> 3;
    const int n = (int)(px / d.HW);
    const int hw = (int)(px - (long long)n * d.HW);
    const int xb = d.x_batch > 0 ? d.x_batch : d.B;
    const int nx = n % xb;
    const float cin = d.coef[nx * d.coef_stride + 2];
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c8 * 8 + j;
        float v = 0.f;
        if (c < d.C) v = cin * d.x[((long long)nx * d.C + c) * d.HW + hw];
        split_h16(v, hi[j], lo[j]);
    }
    __half* o = reinterpret_cast<__half*>(d.out);
    const long long off = px * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(o + off) = *reinterpret_cast<const uint4*>(hi);
    if (d.nplanes > 1) *reinterpret_cast<uint4*>(o + (long long)d.B * d.HW * 64 + off) = *reinterpret_cast<const uint4*>(lo);
}

// one warp per token row; the row lives in registers (C <= 2048), two-pass mean / variance like torch's layer_norm
__global__ void layernorm_kernel(ds_layernorm_desc d) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= d.rows) return;
    const int lane = threadIdx.x & 31;
    const float* x = d.src + row * d.C;
    constexpr int MAXV = 16;                                   // float4 per lane: C <= 32*4*16 = 2048
    float4 v[MAXV];
    const int nv = d.C / 4;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            v[k] = *reinterpret_cast<const float4*>(x + 4 * j);
            sum += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d.C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
            sq += a * a + b * b + c * c + e * e;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)d.C + d.eps);
    __half* out = reinterpret_cast<__half*>(d.out);
    const long long plane = d.rows * d.C;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            const float4 g = *reinterpret_cast<const float4*>(d.gamma + 4 * j);
            const float4 b = *reinterpret_cast<const float4*>(d.beta + 4 * j);
            const float y[4] = {(v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y,
                                (v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w};
            if (d.fmt == 2) {                    // fp32 result (the last LayerNorm of the CLIP text encoder)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) + row * d.C + 4 * j) = make_float4(y[0], y[1], y[2], y[3]);
                continue;
            }
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_h16(y[q], hi[q], lo[q]);
            *reinterpret_cast<uint2*>(out + row * d.C + 4 * j) = *reinterpret_cast<const uint2*>(hi);
            if (d.nplanes > 1) *reinterpret_cast<uint2*>(out + plane + row * d.C + 4 * j) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// quick-GELU (ds_geglu_desc.mode == 1): out = x * sigmoid(1.702 x) on fp32 [rows][I] -> fp16 hi/lo planes (CLIP MLP activation)
__global__ void quick_gelu_kernel(ds_geglu_desc d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 values
    const long long total = d.rows * (d.I / 4);
    if (idx >= total) return;
    const float4 a = *reinterpret_cast<const float4*>(d.src + idx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
    __align__(8) __half hi[4];
    __align__(8) __half lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_h16(av[q] / (1.0f + expf(-1.702f * av[q])), hi[q], lo[q]);
    __half* out = reinterpret_cast<__half*>(d.out);
    *reinterpret_cast<uint2*>(out + idx * 4) = *reinterpret_cast<const uint2*>(hi);
    if (d.nplanes > 1) *reinterpret_cast<uint2*>(out + d.rows * d.I + idx * 4) = *reinterpret_cast<const uint2*>(lo);
}

// token + position embedding (ds_embed_desc): one thread per 4 channels of one row
__global__ void embed_kernel(ds_embed_desc d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = d.C / 4;
    if (idx >= d.rows * c4) return;
    const long long row = idx / c4;
    const int j = (int)(idx - row * c4) * 4;
    int id = d.ids[row];
    id = id < 0 ? 0 : (id >= d.vocab ? d.vocab - 1 : id);
    const float4 t = *reinterpret_cast<const float4*>(d.tok + (long long)id * d.C + j);
    const float4 p = *reinterpret_cast<const float4*>(d.pos + (long long)(row % d.T) * d.C + j);
    *reinterpret_cast<float4*>(d.out + row * d.C + j) = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
}

__global__ void geglu_kernel(ds_geglu_desc d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 outputs
    const int i4 = d.I / 4;
    const long long total = d.rows * i4;
    if (idx >= total) return;
    const long long row = idx / i4;
    const int j = (int)(idx - row * i4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(d.src + row * 2 * d.I + j);
    const float4 g = *reinterpret_cast<const float4*>(d.src + row * 2 * d.I + d.I + j);
    const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
    __align__(8) __half hi[4];
    __align__(8) __half lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float gelu = 0.5f * gv[q] * (1.0f + erff(gv[q] * 0.70710678118654752f));
        split_h16(av[q] * gelu, hi[q], lo[q]);
    }
    __half* out = reinterpret_cast<__half*>(d.out);
    *reinterpret_cast<uint2*>(out + row * d.I + j) = *reinterpret_cast<const uint2*>(hi);
    if (d.nplanes > 1) *reinterpret_cast<uint2*>(out + d.rows * d.I + row * d.I + j) = *reinterpret_cast<const uint2*>(lo);
}

// ---- f8 operand image (csrc/ops.h) of four consecutive values: fp16 (v * 2^A16) | e4m3 ((v - hi) * 2^LO8) | e4m3 (hi * 2^HI8) --------------
// `plane` = elements per plane; o = element offset.  Same arithmetic as gn_store_f8 (which handles eight values).
__device__ __forceinline__ void store4_f8(__half* base, long long plane, long long o, const float* v) {
    __align__(8) uint32_t hi[2];
    __align__(4) unsigned short lo8[2];
    __align__(4) unsigned short hi8[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) f8_image_pair(v[2 * j], v[2 * j + 1], hi[j], lo8[j], hi8[j]);
    *reinterpret_cast<uint2*>(base + o) = *reinterpret_cast<const uint2*>(hi);
    unsigned char* b8 = reinterpret_cast<unsigned char*>(base + plane);
    *reinterpret_cast<unsigned int*>(b8 + o) = *reinterpret_cast<const unsigned int*>(lo8);
    *reinterpret_cast<unsigned int*>(b8 + plane + o) = *reinterpret_cast<const unsigned int*>(hi8);
}

// LayerNorm / GEGLU writing the f8 operand image (fmt == 1; opt-in, feeds an f8 GEMM).  Separate kernels so that the default ones above
// stay byte-identical; the arithmetic before the store is the same.
__global__ void layernorm_f8_kernel(ds_layernorm_desc d) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= d.rows) return;
    const int lane = threadIdx.x & 31;
    const float* x = d.src + row * d.C;
    constexpr int MAXV = 16;
    float4 v[MAXV];
    const int nv = d.C / 4;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            v[k] = *reinterpret_cast<const float4*>(x + 4 * j);
            sum += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d.C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
            sq += a * a + b * b + c * c + e * e;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)d.C + d.eps);
    __half* out = reinterpret_cast<__half*>(d.out);
    const long long plane = d.rows * d.C;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int j = lane + 32 * k;
        if (j < nv) {
            const float4 g = *reinterpret_cast<const float4*>(d.gamma + 4 * j);
            const float4 b = *reinterpret_cast<const float4*>(d.beta + 4 * j);
            const float y[4] = {(v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y,
                                (v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w};
            store4_f8(out, plane, row * d.C + 4 * j, y);
        }
    }
}

__global__ void geglu_f8_kernel(ds_geglu_desc d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i4 = d.I / 4;
    const long long total = d.rows * i4;
    if (idx >= total) return;
    const long long row = idx / i4;
    const int j = (int)(idx - row * i4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(d.src + row * 2 * d.I + j);
    const float4 g = *reinterpret_cast<const float4*>(d.src + row * 2 * d.I + d.I + j);
    const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q] = av[q] * (0.5f * gv[q] * (1.0f + erff(gv[q] * 0.70710678118654752f)));
    store4_f8(reinterpret_cast<__half*>(d.out), d.rows * d.I, row * d.I + j, y);
}

__global__ void chanmean_kernel(ds_chanmean_desc d) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= d.rows) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int c = lane; c < d.C; c += 32) s += d.src[row * d.C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) d.out[row] = s / (float)d.C;
}

static inline int ok() { return cudaGetLastError() == cudaSuccess ? 0 : -1; }

}  // namespace dsb

using namespace dsb;

extern "C" int ds_gn_stats_launch(const ds_gn_stats_desc* d, cudaStream_t stream) {
    const int C = d->C0 + d->C1;
    if (C % 4 || d->groups > 64 || C % d->groups || (d->C0 % 4)) return -2;
    const int ncol4 = C / 4;
    int bx = ncol4 < 256 ? ncol4 : 256;
    // keep bx a divisor-friendly size; columns loop with stride bx anyway
    int by = 256 / bx; if (by < 1) by = 1;
    int pix_per_cta = 8 * by;                  // 8 loads in flight per thread
    if (pix_per_cta < 64) pix_per_cta = 64;
    // small batches (latent diffusion: 16 samples): shorter pixel runs per CTA so that the grid still covers the SMs
    const int min_pix = by > 8 ? by : 8;
    while (pix_per_cta > min_pix && (long long)((d->HW + pix_per_cta - 1) / pix_per_cta) * d->B < 148 * 4) pix_per_cta /= 2;
    const int chunks = (d->HW + pix_per_cta - 1) / pix_per_cta;
    gn_stats_kernel<<<dim3(chunks, d->B), dim3(bx, by), 0, stream>>>(*d, pix_per_cta);
    return ok();
}

extern "C" int ds_gn_finalize_launch(const ds_gn_finalize_desc* d, cudaStream_t stream) {
    const int C = d->C0 + d->C1;
    if (d->groups <= 0 || d->groups > 64 || C % d->groups) return -2;
    if (d->quads0) {
        const int u0 = d->unit0 == 2 ? 2 : 4, u1 = d->unit1 == 2 ? 2 : 4;
        const int cpg = C / d->groups;
        if (d->C0 % u0 || d->C1 % u1 || (d->C1 > 0 && !d->quads1)) return -2;
        // every group must be a union of whole partial units of the sources it covers
        const int rem = d->C0 % cpg;                    // channels of a group that straddles the two sources, on the source-0 side
        if (cpg % u0 || rem % u0) return -2;
        if (d->C1 > 0 && (cpg % u1 || (rem ? (cpg - rem) % u1 : 0))) return -2;
    }
    if (!d->quads0 && !d->coef) return -2;               // nothing to do
    if (d->coef && (!d->gamma || !d->beta || d->HW <= 0)) return -2;
    // threads: enough for (columns x slab groups), at most 1024; shared memory: [cols] + [SG][cols] doubles, cols <= C, SG * cols <= max(cols, threads)
    int threads = 256;
    size_t smem = (size_t)(C + 2) * sizeof(double);
    if (d->quads0) {
        const int u0 = d->unit0 == 2 ? 2 : 4, u1 = d->unit1 == 2 ? 2 : 4;
        const int cols = d->C0 / u0 * 2 + d->C1 / u1 * 2;
        threads = 1024;
        const int sg = cols >= threads ? 1 : threads / cols;
        smem = (size_t)(cols + (size_t)sg * cols + 2) * sizeof(double);
    }
    gn_finalize_kernel<<<d->B, threads, smem, stream>>>(*d);
    return ok();
}

extern "C" int ds_gn_apply_launch(const ds_gn_apply_desc* d, cudaStream_t stream) {
    const int C = d->C0 + d->C1;
    if (C % 8 || (d->C0 % 8)) return -2;
    if (d->fmt != 0 && (d->fmt != 1 || d->resample == 3 || d->nplanes != 2)) return -2;   // the f8 layout reuses the two-plane footprint
    const int nc8 = C / 8;
    if (nc8 > 512) return -2;
    int rows = 256 / nc8;
    if (rows < 1) rows = 1;
    const int threads = nc8 * rows;
    const int Ho = d->resample == 1 ? d->H / 2 : (d->resample == 2 ? d->H * 2 : d->H);
    const int Wo = d->resample == 1 ? d->W / 2 : (d->resample == 2 ? d->W * 2 : d->W);
    const int npix = Ho * Wo;
    // ~16 pixels per thread amortise the per-thread coefficient set-up; keep at least ~4 CTAs per SM in flight overall
    int pix_per_cta = rows * 16;
    while (pix_per_cta > rows && (long long)((npix + pix_per_cta - 1) / pix_per_cta) * d->B < 148 * 4) pix_per_cta /= 2;
    const int chunks = (npix + pix_per_cta - 1) / pix_per_cta;
    dim3 grid(chunks, d->B);
    if (d->coef && d->resample == 0 && d->sums == nullptr && threads > 256) return -2;      // wider than 2048 channels: use the sums path
    if (d->coef && d->resample == 0 && d->sums == nullptr) {
        // persistent variant: grid = co-resident CTAs (occupancy query per block size and device, cached)
        static int occ[64][17] = {};
        static int sms[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        const int slot = threads / 32;
        if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (!occ[dev][slot]) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gn_apply_v3_kernel, threads, 0) != cudaSuccess || nb < 1) nb = 1;
            occ[dev][slot] = nb;
        }
        const int ups = (npix + rows - 1) / rows;
        const long long total_units = (long long)ups * d->B;
        long long g3 = (long long)sms[dev] * occ[dev][slot];
        const long long total_chunks = (long long)((ups + 7) / 8) * d->B;        // kChunk = 8 units per chunk (gn_apply_v3_kernel)
        if (g3 > total_chunks) g3 = total_chunks;
        if (g3 < 1) g3 = 1;
        gn_apply_v3_kernel<<<(unsigned)g3, threads, 0, stream>>>(*d, nc8, rows, ups, total_units);
        return ok();
    }
    static const int use_v2 = [] { const char* e = getenv("DSB_GN_APPLY_V2"); return e ? atoi(e) : 1; }();
    if (use_v2 && d->resample == 0) {
        gn_apply_v2_kernel<<<grid, threads, 0, stream>>>(*d, pix_per_cta, nc8, rows);
        return ok();
    }
    if (d->resample == 1) gn_apply_kernel<1><<<grid, threads, 0, stream>>>(*d, pix_per_cta, nc8, rows);
    else if (d->resample == 3) gn_apply_kernel<3><<<grid, threads, 0, stream>>>(*d, pix_per_cta, nc8, rows);
    else if (d->resample == 2) gn_apply_kernel<2><<<grid, threads, 0, stream>>>(*d, pix_per_cta, nc8, rows);
    else gn_apply_kernel<0><<<grid, threads, 0, stream>>>(*d, pix_per_cta, nc8, rows);
    return ok();
}

extern "C" int ds_embed_launch(const ds_embed_desc* d, cudaStream_t stream) {
    if (d->C % 4 || d->rows <= 0 || d->T <= 0 || d->vocab <= 0) return -2;
    const long long total = d->rows * (d->C / 4);
    embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_layernorm_launch(const ds_layernorm_desc* d, cudaStream_t stream) {
    if (d->C % 4 || d->C > 2048) return -2;
    const int wpb = 8;
    if (d->fmt == 1) {
        if (d->nplanes != 2) return -2;
        layernorm_f8_kernel<<<(unsigned)((d->rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(*d);
        return ok();
    }
    layernorm_kernel<<<(unsigned)((d->rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_geglu_launch(const ds_geglu_desc* d, cudaStream_t stream) {
    if (d->I % 4) return -2;
    const long long total = d->rows * (d->I / 4);
    if (d->mode == 1) {
        if (d->fmt != 0) return -2;
        quick_gelu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(*d);
        return ok();
    }
    if (d->fmt == 1) {
        if (d->nplanes != 2) return -2;
        geglu_f8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(*d);
        return ok();
    }
    geglu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_softmax_launch(const ds_softmax_desc* d, cudaStream_t stream) {
    const bool dense = (d->pitch_in == 0 || d->pitch_in == d->L) && (d->pitch_out == 0 || d->pitch_out == d->L) && (d->L % 4 == 0);
    if (dense && d->L <= 1024) {
        const unsigned blocks = (unsigned)((d->rows + 7) / 8);
        if (d->L <= 256) softmax_reg_kernel<2, false><<<blocks, 256, 0, stream>>>(*d);
        else softmax_reg_kernel<8, false><<<blocks, 256, 0, stream>>>(*d);
        return ok();
    }
    if (dense && d->L <= 8192) {
        if (d->L <= 4096) softmax_reg_kernel<4, true><<<(unsigned)d->rows, 256, 0, stream>>>(*d);
        else softmax_reg_kernel<8, true><<<(unsigned)d->rows, 256, 0, stream>>>(*d);
        return ok();
    }
    const int wpb = 8;
    const long long blocks = (d->rows + wpb - 1) / wpb;
    softmax_kernel<<<(unsigned)blocks, wpb * 32, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_posemb_launch(const ds_posemb_desc* d, cudaStream_t stream) {
    posemb_kernel<<<d->nsig, 128, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_linear_launch(const ds_linear_desc* d, cudaStream_t stream) {
    const int wpb = 4;
    const int bx = (d->out_f + wpb - 1) / wpb;
    int by = d->n_rows < 16 ? d->n_rows : 16;
    if (by < 1) by = 1;
    if (d->in_f <= 256) linear_kernel<256><<<dim3(bx, by), wpb * 32, 0, stream>>>(*d);
    else if (d->in_f <= 512) linear_kernel<512><<<dim3(bx, by), wpb * 32, 0, stream>>>(*d);
    else if (d->in_f <= 1024) linear_kernel<1024><<<dim3(bx, by), wpb * 32, 0, stream>>>(*d);
    else if (d->in_f <= 2048) linear_kernel<2048><<<dim3(bx, by), wpb * 32, 0, stream>>>(*d);
    else return -2;
    return ok();
}

extern "C" int ds_prep_input_launch(const ds_prep_input_desc* d, cudaStream_t stream) {
    if (d->C > 64) return -2;
    const long long total = (long long)d->B * d->HW * 8;
    prep_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(*d);
    return ok();
}

extern "C" int ds_chanmean_launch(const ds_chanmean_desc* d, cudaStream_t stream) {
    const int wpb = 8;
    chanmean_kernel<<<(unsigned)((d->rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(*d);
    return ok();
}
