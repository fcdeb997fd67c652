// Fused softmax attention for sm_100a (head dim padded to 64): O = softmax(scale * Q K^T) V without materialising the L x Lk
// score matrix in HBM.  Reference: diff-solvers-main/models/networks_edm.py:105-118 (AttentionOp) / :174-178 (UNetBlock attention),
// models/ldm/modules/attention.py:152-196 (CrossAttention.forward).
//
// One CTA per (sample, head, 128-query tile); keys/values stream through in blocks of 128:
//
//   warp 0      TMA producer   Q tile once; K_j [128 keys x 64] and V_j^T [64 x 128 keys] (fp16 hi/lo planes) into 2-stage rings
//   warp 1      MMA issuer     S_j = Q K_j^T  (3 split-precision passes, 128x128x16 tcgen05.mma, two S buffers in TMEM)
//                              O_j = P_j V_j  (3 passes, 128x64x16, fresh TMEM accumulator per block)
//   warps 2..9  softmax        one query row and one half (64 keys) of the block per thread: running max / sum, p = exp2(s - m)
//                              split into fp16 hi/lo and written to shared memory in the 128B-swizzled K-major layout the MMA
//                              reads; the running output lives in registers (O = alpha * O + O_j), so the TMEM accumulator
//                              never needs rescaling.
//
// QK^T of block j+1 is issued before the softmax of block j finishes (two S buffers); the softmax row-max pass of block j+1
// overlaps P_j V_j.  Same split-precision contract as the GEMM kernel: Q, K, V and P are fp16 hi + lo planes, products are
// hi*hi + lo*hi + hi*lo in fp32.
#include "ops.h"
#include "ptx.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>

namespace dsb {

int encode_map(CUtensorMap* m, const void* ptr, int rank, const int64_t* dims, const int64_t* strides_bytes, const int32_t* box);

static constexpr int kAttnThreads = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 softmax
static constexpr int kQBytes = 2 * 16384;          // Q hi, lo: 128 rows x 64 fp16 each
static constexpr int kKStage = 2 * 16384;          // K hi, lo: 128 keys x 64 fp16
static constexpr int kVStage = 4 * 8192;           // V^T [plane][key block of 64]: 64 d-rows x 64 keys fp16
static constexpr int kPBytes = 4 * 16384;          // P [plane][key block of 64]: 128 rows x 64 keys fp16
static constexpr int kOffK = kQBytes;
static constexpr int kOffV = kOffK + 2 * kKStage;
static constexpr int kOffP = kOffV + 2 * kVStage;
static constexpr int kOffCtl = kOffP + kPBytes;    // 224 KB

struct alignas(64) AttnKernelParams {
    CUtensorMap tmQ, tmK, tmV;
    int B, nh, L, Lk, q_c0, k_c0, q_tiles;
    float scale_log2e;
    __half* out;
    long long o_plane;
    int o_pitch;
};

struct AttnCtl {
    uint64_t q_full;
    uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full, o_full, o_empty;
    uint32_t tmem_base;
    float xch[2][128];          // row maximum / row sum exchange between the two key halves of a row
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1) attn_kernel(const __grid_constant__ AttnKernelParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    AttnCtl* ctl = reinterpret_cast<AttnCtl*>(smem + kOffCtl);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x % p.q_tiles;
    const int z = blockIdx.x / p.q_tiles;
    const int h = z % p.nh;
    const int b = z / p.nh;
    const int nkv = (p.Lk + 127) >> 7;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmQ);
        tma_prefetch_desc(&p.tmK);
        tma_prefetch_desc(&p.tmV);
        mbar_init(&ctl->q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&ctl->k_full[s], 1);
            mbar_init(&ctl->k_empty[s], 1);
            mbar_init(&ctl->v_full[s], 1);
            mbar_init(&ctl->v_empty[s], 1);
            mbar_init(&ctl->s_full[s], 1);
            mbar_init(&ctl->s_empty[s], 8);
        }
        mbar_init(&ctl->p_full, 8);
        mbar_init(&ctl->o_full, 1);
        mbar_init(&ctl->o_empty, 8);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(&ctl->tmem_base, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    // TMEM columns: S buffers at 0 and 128, per-block P.V result at 256 (64 columns)

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&ctl->q_full, kQBytes);
            tma_load_3d(&p.tmQ, &ctl->q_full, smem, p.q_c0 + h * 64, qt * 128, b);
            tma_load_3d(&p.tmQ, &ctl->q_full, smem + 16384, p.q_c0 + h * 64, qt * 128, p.B + b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&ctl->k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&ctl->k_full[s], kKStage);
                uint8_t* sk = smem + kOffK + s * kKStage;
                tma_load_3d(&p.tmK, &ctl->k_full[s], sk, p.k_c0 + h * 64, j * 128, b);
                tma_load_3d(&p.tmK, &ctl->k_full[s], sk + 16384, p.k_c0 + h * 64, j * 128, p.B + b);
                mbar_wait(&ctl->v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&ctl->v_full[s], kVStage);
                uint8_t* sv = smem + kOffV + s * kVStage;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_3d(&p.tmV, &ctl->v_full[s], sv + pl * 16384 + kb * 8192, j * 128 + kb * 64, h * 64, pl * p.B + b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128);
            const uint32_t idesc_o = umma_idesc_f16(64);
            const uint32_t sq = smem_u32(smem);
            const uint32_t sp = smem_u32(smem + kOffP);
            auto issue_qk = [&](int j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&ctl->k_full[s], ph);
                mbar_wait(&ctl->s_empty[s], ph ^ 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem + kOffK + s * kKStage);
                const uint32_t d_tmem = tmem_base + s * 128;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint64_t da = umma_desc_sw128(sq + (pass == 1 ? 16384 : 0));
                    const uint64_t db = umma_desc_sw128(sk + (pass == 2 ? 16384 : 0));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc_s, (pass > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&ctl->k_empty[s]);
                umma_commit(&ctl->s_full[s]);
            };
            mbar_wait(&ctl->q_full, 0);
            issue_qk(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) issue_qk(j + 1);
                const int s = j & 1;
                mbar_wait(&ctl->v_full[s], (j >> 1) & 1);
                mbar_wait(&ctl->o_empty, (j & 1) ^ 1);        // softmax warps have read the previous block's P.V result
                mbar_wait(&ctl->p_full, j & 1);
                tc_fence_after();
                const uint32_t sv = smem_u32(smem + kOffV + s * kVStage);
                const uint32_t d_tmem = tmem_base + 256;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_pl = sp + (pass == 1 ? 32768 : 0);
                    const uint32_t b_pl = sv + (pass == 2 ? 16384 : 0);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = umma_desc_sw128(a_pl + kb * 16384);
                        const uint64_t db = umma_desc_sw128(b_pl + kb * 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc_o, (pass > 0 || kb > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&ctl->v_empty[s]);
                umma_commit(&ctl->o_full);
            }
        }
    } else {
        // Eight softmax warps: warps w and w + 4 share a TMEM lane quadrant (w % 4); each thread owns one query row and one HALF of
        // the block's 128 keys (half = key block of 64) plus the matching 32 output columns.  Two warps per scheduler hide each
        // other's latencies; the halves meet once per block (row maximum, through shared memory + a named barrier).
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;               // 0: warps 2..5, 1: warps 6..9
        const int row = quad * 32 + lane;               // query row inside the tile
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint8_t* sP = smem + kOffP + half * 16384;
        float m = -INFINITY, l = 0.f;
        float O[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) O[i] = 0.f;
        auto add_block_output = [&]() {
            uint32_t v[32];
            DSB_TMEM_LD_32(t_row + 256 + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) O[i] += __uint_as_float(v[i]);
        };
        for (int j = 0; j < nkv; ++j) {
            const int sb = j & 1;
            mbar_wait(&ctl->s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            const int kvalid = p.Lk - j * 128 - half * 64;   // keys of this thread's half that exist (>= 64: all)
            const uint32_t t_s = t_row + sb * 128 + half * 64;
            // pass 1: maximum over this thread's 64 keys (four independent chains, second chunk requested early)
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            {
                uint32_t v[2][32];
                DSB_TMEM_LD_32(t_s, v[0]);
                tmem_ld_wait();
                DSB_TMEM_LD_32(t_s + 32, v[1]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (kvalid >= 64) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[c][i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < kvalid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[c][i]));
                    }
                    if (c == 0) tmem_ld_wait();
                }
            }
            const float mx_own = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // meet the other half of the row.  Slots alternate with the block parity so that one barrier per block is enough:
            // the slot written in block j is next overwritten (by the partner) in block j + 1, after this thread has read it.
            ctl->xch[half ^ (j & 1)][row] = mx_own;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mx = fmaxf(mx_own, ctl->xch[(half ^ 1) ^ (j & 1)][row]);
            const float m_new = fmaxf(m, mx * p.scale_log2e);
            const float alpha = ex2_approx(m - m_new);  // first block: exp2(-inf) = 0
            if (j > 0) {
                // P_{j-1} V_{j-1} has completed: its result is readable and the P buffer is free again
                mbar_wait(&ctl->o_full, (j - 1) & 1);
                tc_fence_after();
                add_block_output();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->o_empty);
            }
            if (alpha != 1.f) {
#pragma unroll
                for (int i = 0; i < 32; ++i) O[i] *= alpha;
                l *= alpha;
            }
            // pass 2: p = exp2(s * scale * log2e - m), split into fp16 hi / lo (packed half2 conversions), swizzled K-major store
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            {
                uint32_t v[2][32];
                DSB_TMEM_LD_32(t_s, v[0]);
                tmem_ld_wait();
                DSB_TMEM_LD_32(t_s + 32, v[1]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int col = c * 32 + g * 8 + 2 * i;
                            float p0 = ex2_approx(fmaf(__uint_as_float(v[c][g * 8 + 2 * i]), p.scale_log2e, -m_new));
                            float p1 = ex2_approx(fmaf(__uint_as_float(v[c][g * 8 + 2 * i + 1]), p.scale_log2e, -m_new));
                            if (kvalid < 64) {
                                if (col >= kvalid) p0 = 0.f;
                                if (col + 1 >= kvalid) p1 = 0.f;
                            }
                            l4[i] += p0 + p1;
                            const __half2 h2 = __floats2half2_rn(p0, p1);
                            const float2 hf = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
                            hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
                            lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                        const int chunk = c * 4 + g;                         // 16-byte chunk of the 128-byte row
                        const uint32_t off = row * 128 + ((chunk ^ (row & 7)) << 4);
                        *reinterpret_cast<uint4*>(sP + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(sP + 32768 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    if (c == 0) tmem_ld_wait();
                }
            }
            l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            m = m_new;
            tc_fence_before();
            fence_proxy_async();                        // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ctl->s_empty[sb]);
                mbar_arrive(&ctl->p_full);
            }
        }
        mbar_wait(&ctl->o_full, (nkv - 1) & 1);
        tc_fence_after();
        add_block_output();
        // row sum: both halves
        asm volatile("bar.sync 1, 256;" ::: "memory");
        ctl->xch[half][row] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += ctl->xch[half ^ 1][row];
        const int grow = qt * 128 + row;
        if (grow < p.L) {
            const float inv = 1.f / l;
            __half* o = p.out + ((long long)b * p.L + grow) * p.o_pitch + h * 64 + half * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a0 = O[g * 8 + 2 * i] * inv, a1 = O[g * 8 + 2 * i + 1] * inv;
                    const __half2 h2 = __floats2half2_rn(a0, a1);
                    const float2 hf = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                    hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
                    lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                *reinterpret_cast<uint4*>(o + g * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(o + p.o_plane + g * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ host
int attn_build(const ds_attn_desc* d, AttnKernelParams* kp) {
    if (d->nplanes != 2 || d->B <= 0 || d->nh <= 0 || d->L <= 0 || d->Lk <= 0 || !(d->scale > 0.f)) return -30;
    if (d->q_pitch % 8 || d->k_pitch % 8 || d->vt_pitch % 8 || d->o_pitch % 8 || d->q_c0 % 8 || d->k_c0 % 8) return -31;
    if (d->q_c0 + d->nh * 64 > d->q_pitch || d->k_c0 + d->nh * 64 > d->k_pitch || d->nh * 64 > d->o_pitch || d->Lk > d->vt_pitch) return -32;
    {
        const int64_t dims[3] = {d->q_pitch, d->L, (int64_t)2 * d->B};
        const int64_t str[2] = {(int64_t)d->q_pitch * 2, (int64_t)d->L * d->q_pitch * 2};
        const int32_t box[3] = {64, 128, 1};
        if (encode_map(&kp->tmQ, d->q, 3, dims, str, box)) return -33;
    }
    {
        const int64_t dims[3] = {d->k_pitch, d->Lk, (int64_t)2 * d->B};
        const int64_t str[2] = {(int64_t)d->k_pitch * 2, (int64_t)d->Lk * d->k_pitch * 2};
        const int32_t box[3] = {64, 128, 1};
        if (encode_map(&kp->tmK, d->k, 3, dims, str, box)) return -34;
    }
    {
        const int64_t dims[3] = {d->Lk, (int64_t)d->nh * 64, (int64_t)2 * d->B};
        const int64_t str[2] = {(int64_t)d->vt_pitch * 2, (int64_t)d->nh * 64 * d->vt_pitch * 2};
        const int32_t box[3] = {64, 64, 1};
        if (encode_map(&kp->tmV, d->vt, 3, dims, str, box)) return -35;
    }
    kp->B = d->B; kp->nh = d->nh; kp->L = d->L; kp->Lk = d->Lk; kp->q_c0 = d->q_c0; kp->k_c0 = d->k_c0;
    kp->q_tiles = (d->L + 127) / 128;
    kp->scale_log2e = d->scale * 1.4426950408889634f;
    kp->out = reinterpret_cast<__half*>(d->out);
    kp->o_plane = (long long)d->B * d->L * d->o_pitch;
    kp->o_pitch = d->o_pitch;
    return 0;
}

size_t attn_params_size() { return sizeof(AttnKernelParams); }

int attn_run(const AttnKernelParams* kp, cudaStream_t stream) {
    static bool attr_set = false;
    const size_t smem = kOffCtl + sizeof(AttnCtl) + 1024;
    if (!attr_set) {
        if (cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -36;
        attr_set = true;
    }
    const long long grid = (long long)kp->B * kp->nh * kp->q_tiles;
    if (grid <= 0 || grid > 0x7fffffffLL) return -37;
    attn_kernel<<<(unsigned)grid, kAttnThreads, smem, stream>>>(*kp);
    return cudaGetLastError() == cudaSuccess ? 0 : -38;
}

}  // namespace dsb

extern "C" int ds_attn_launch(const ds_attn_desc* d, cudaStream_t stream) {
    dsb::AttnKernelParams kp;
    int rc = dsb::attn_build(d, &kp);
    if (rc) return rc;
    return dsb::attn_run(&kp, stream);
}
