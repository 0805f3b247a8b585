// vmgym_train.cu — the dense layers of PPOAgent.update (src/agents/ppo.py:91-109 Network, :229-295 update) as hand-written
// tcgen05 GEMMs with fused epilogues, bf16 operands / fp32 accumulation (TMEM):
//
//   tc_gemm_kernel      C[M, N] = sum_k A(m, k) B(n, k) for operands that are K-major (row-major [rows, K], e.g. activations and
//                       nn.Linear weights in the forward pass) or MN-major (row-major [K, rows]: the same tensors seen by the
//                       backward GEMMs, which contract over samples or over output features) — no transposed copies anywhere.
//                       Epilogue (one pass out of tensor memory): + bias, tanh, * (1 - y^2) (tanh backward), fp32 store or
//                       accumulate (weight gradients summed over sample chunks), bf16 store (next layer's operand), and the
//                       row sums of A through one extra N = 16 MMA against a tile of ones (bias gradients for free).
//   layers:  forward   a = tanh(x W^T + b)                  A = x (K-major),  B = W (K-major)
//            backward  dz_prev = (dz W) * (1 - a_prev^2)    A = dz (K-major), B = W (MN-major)
//                      dW += dz^T a_prev, db += sum dz      A = dz (MN-major), B = a_prev (MN-major), row sums of A
//   small kernels: fp32 -> padded bf16 cast, the value head (row dot) and its backward, the per-sample PPO loss coefficients
//   (clipped surrogate ppo.py:267-269, clipped value loss :271-280, entropy bonus :282).
// The masked multi-categorical head in front of the output layer lives in vmgym_gemm.cu (policy_fused kernels: forward statistics
// and, for the update, the logit gradients written as bf16 — the only [samples, V*A] tensor of the update that reaches HBM).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/vmgym.h"
#include "vmgym_tc.cuh"

extern "C" void vmgym_internal_set_error(const char* msg);

namespace vmgym_train {
using namespace vmgym_gemm;

constexpr int T_BM = 128, T_BN = 256, T_BK = 64;
constexpr int T_MAX_STAGES = 6;
constexpr int T_A_BYTES = T_BM * T_BK * 2;                 // 16 KiB
constexpr int T_B_BYTES = T_BN * T_BK * 2;                 // 32 KiB (256-column tiles; 16 KiB with 128-column tiles)
constexpr int T_RING_BYTES = 4 * (T_A_BYTES + T_B_BYTES);  // 192 KiB operand ring: 4 stages of 48 KiB or 6 stages of 32 KiB
constexpr int T_ONES_BYTES = 16 * T_BK * 2;                // 2 KiB: 16 rows x 64 bf16 of 1.0 (any layout: all elements equal)
constexpr int T_TMEM_COLS = 512;                           // 256 accumulator columns + 16 for the row sums -> next power of two
constexpr int T_THREADS = 320;                            // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr size_t T_SMEM_BYTES = (size_t)T_RING_BYTES + T_ONES_BYTES + 1024 /* alignment */ + 256 /* barriers */;

struct GemmArgs {
    int M, N, K;
    int a_mn, b_mn;                    // 0: K-major operand (row-major [rows, K]); 1: MN-major (row-major [K, rows])
    const float* bias;                 // [N] or nullptr
    int act;                           // 0 none, 1 tanh, 2 relu
    int f32_pre;                       // the fp32 output takes the value BEFORE the activation (bias included)
    int bf16_split;                    // the bf16 output is the split operand [hi | lo | hi] of vmgym_cast_split_bf16, segments N wide
    const __nv_bfloat16* mul_y;        // [M, ldy]: result *= 1 - y^2, or nullptr
    long long ldy;
    float* c_f32;                      // fp32 output [M, ldc_f32] or nullptr
    long long ldc_f32;
    int accumulate;                    // c_f32 += result instead of =
    __nv_bfloat16* c_bf16;             // bf16 output [M, ldc_bf16] or nullptr
    long long ldc_bf16;
    float* row_sum;                    // [M]: (+)= sum_k A(m, k), written by the CTAs of the first N tile, or nullptr
    int tile_n;                        // columns per CTA tile: 256, or 128 when 256-wide tiles would leave SMs idle
    int k_splits;                      // > 1: split-K (gridDim.z): each CTA contracts a slice of K and adds its partial tile to c_f32 /
                                       // row_sum atomically (weight-gradient GEMMs with few output tiles and a long sample axis)
};

// MN-major, 128B-swizzled operand tile: 64-element (128 B) runs along M/N, K rows 128 B apart, 8-row groups 1024 B apart (SBO),
// 64-element M/N blocks 8192 B apart (LBO) — the layout of 64 x 64 TMA boxes stored back to back
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t umma_idesc(int m, int n, int a_mn, int b_mn)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float tanh_fast(float x)
{
    // one MUFU op; |error| ~ 2^-11 relative, far inside the bf16 rounding (2^-9) the result gets anyway
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(T_THREADS, 1) tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_b, const GemmArgs g)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_ones = smem + (size_t)T_RING_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ones + T_ONES_BYTES);
    uint64_t* empty_bar = full_bar + T_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + T_MAX_STAGES;
    // narrower tiles leave room for a deeper ring: the K loop of a small GEMM is bound by the latency of its operand loads
    const int T_STAGE_BYTES = T_A_BYTES + g.tile_n * T_BK * 2;
    const int T_STAGES = g.tile_n <= 128 ? 6 : 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * T_BM, n0 = blockIdx.x * g.tile_n;
    const int bn = min(g.tile_n, (g.N - n0 + 15) & ~15);                 // columns of this tile, rounded up to the MMA's N granularity
    const int k_blocks_all = (g.K + T_BK - 1) / T_BK;
    const int kb_per = (k_blocks_all + g.k_splits - 1) / g.k_splits;
    const int kb0 = blockIdx.z * kb_per;                               // this CTA's K slice: k-blocks [kb0, kb0 + k_blocks)
    const int k_blocks = max(0, min(kb_per, k_blocks_all - kb0));
    const bool do_rows = g.row_sum != nullptr && blockIdx.x == 0;
    const bool split = g.k_splits > 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < T_MAX_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(T_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && do_rows) {
        // the tile of ones, written through the generic proxy and made visible to the tensor core's async proxy
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(s_ones);
        for (int i = threadIdx.x - 64; i < T_ONES_BYTES / 2; i += T_THREADS - 64) o[i] = __float2bfloat16(1.0f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const uint32_t a_bytes = T_A_BYTES, b_bytes = (uint32_t)(g.b_mn ? ((bn + 63) / 64) * 8192 : g.tile_n * T_BK * 2);
            for (int kb = 0; kb < k_blocks; kb++) {
                const int s = kb % T_STAGES;
                mbar_wait(&empty_bar[s], (((uint32_t)(kb / T_STAGES)) & 1u) ^ 1u);
                unsigned char* sa = smem + (size_t)s * T_STAGE_BYTES;
                unsigned char* sb = sa + T_A_BYTES;
                mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
                const int kc = (kb0 + kb) * T_BK;                       // first K index of this block
                if (g.a_mn) {                                           // two 64 (M) x 64 (K) boxes
                    tma_load_2d(sa, &map_a, m0, kc, &full_bar[s]);
                    tma_load_2d(sa + 8192, &map_a, m0 + 64, kc, &full_bar[s]);
                } else {
                    tma_load_2d(sa, &map_a, kc, m0, &full_bar[s]);      // one 64 (K) x 128 (M) box
                }
                if (g.b_mn) {
                    for (int j = 0; j * 64 < bn; j++) tma_load_2d(sb + j * 8192, &map_b, n0 + j * 64, kc, &full_bar[s]);
                } else {
                    tma_load_2d(sb, &map_b, kc, n0, &full_bar[s]);      // one 64 (K) x tile_n (N) box
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the K loop (uniform control flow), one elected lane issues each K block as one
        // instruction group (vmgym_tc.cuh umma_bf16_k64_ex).  All 512 TMEM columns are this CTA's: the accumulator starts at column 0.
        if (tmem_base != 0u) __trap();
        const uint32_t idesc = umma_idesc(T_BM, bn, g.a_mn, g.b_mn);
        const uint32_t idesc_rows = umma_idesc(T_BM, 16, g.a_mn, 0);
        const uint32_t ones_lo = umma_desc_lo(smem_u32(s_ones));            // all elements equal: any K-major view of it is "ones"
        // descriptor words: K-major = umma_desc_sw128 (LBO 16 B, K step 32 B), MN-major = umma_desc_mn_sw128 (LBO 8192 B, K step 2048 B)
        const uint32_t hi = UMMA_DESC_HI;
        const uint32_t lbo_a = g.a_mn ? (uint32_t)(8192 >> 4) << 16 : 1u << 16, lbo_b = g.b_mn ? (uint32_t)(8192 >> 4) << 16 : 1u << 16;
        const uint32_t step_a = g.a_mn ? 2048u >> 4 : (UMMA_K * 2) >> 4, step_b = g.b_mn ? 2048u >> 4 : (UMMA_K * 2) >> 4;
        const uint32_t ring_lo = (smem_u32(smem) >> 4) & 0x3FFFu;
        uint32_t s = 0, phase = 0;
#pragma unroll 1
        for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(&full_bar[s], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t a_lo = ring_lo + s * (uint32_t)(T_STAGE_BYTES >> 4);
                const uint32_t b_lo = a_lo + (T_A_BYTES >> 4);
                umma_bf16_k64_ex(0u, a_lo | lbo_a, b_lo | lbo_b, step_a, step_b, hi, hi, idesc, kb ? 1u : 0u);
                // row sums: the same A blocks against the tile of ones (K-major, every K step reads the same 16 x 16 ones)
                if (do_rows) umma_bf16_k64_ex(T_BN, a_lo | lbo_a, ones_lo, step_a, 0u, hi, hi, idesc_rows, kb ? 1u : 0u);
                umma_commit(&empty_bar[s]);
                if (kb == k_blocks - 1) umma_commit(tmem_full_bar);
            }
            __syncwarp();
            if (++s == (uint32_t)T_STAGES) { s = 0; phase ^= 1u; }
        }
        if (k_blocks == 0) {                                                // an empty K slice of a split launch: release the epilogue
            if (elect_one()) umma_commit(tmem_full_bar);
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..9; a warp may touch TMEM lanes 32 (warp % 4) .. +31 = 32 rows of the tile (thread = row); the two
        // warps of a lane quarter split the tile's columns (the tile's epilogue is not overlapped with a main loop, so its
        // latency counts in full: measured 2x on the forward layers) =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                                  // 0: first half of the tile's 32-column chunks, 1: second half
        const int row = m0 + q * 32 + lane;
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = row < g.M && k_blocks > 0;                // (an empty K slice of a split launch adds nothing)
        const int n_chunks = (bn + 31) >> 5, c_mid = (n_chunks + 1) >> 1;
        const int c_begin = half ? c_mid * 32 : 0, c_end = half ? bn : min(bn, c_mid * 32);
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            uint32_t r[32];
            tmem_ld_row32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (!row_ok) continue;
            const int col0 = n0 + c0;
            const int nc = min(32, g.N - col0);
            if (nc <= 0) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float x = __uint_as_float(r[j]);
                if (g.bias && j < nc) x += g.bias[col0 + j];
                v[j] = x;
            }
            if (g.c_f32 && g.f32_pre) {
                float* dst = g.c_f32 + (long long)row * g.ldc_f32 + col0;
                if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < nc; j++) dst[j] = v[j];
                }
            }
            if (g.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = tanh_fast(v[j]);
            } else if (g.act == 2) {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.0f);
            }
            if (g.mul_y) {
                const __nv_bfloat16* y = g.mul_y + (long long)row * g.ldy + col0;
                if (nc == 32 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 u = *reinterpret_cast<const uint4*>(y + j);
                        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const float2 f = __bfloat1622float2(h[t]);
                            v[j + 2 * t] *= 1.0f - f.x * f.x;
                            v[j + 2 * t + 1] *= 1.0f - f.y * f.y;
                        }
                    }
                } else {
                    for (int j = 0; j < nc; j++) { const float f = __bfloat162float(y[j]); v[j] *= 1.0f - f * f; }
                }
            }
            if (g.c_f32 && g.f32_pre) {
                // already stored above
            } else if (g.c_f32 && split) {
                float* dst = g.c_f32 + (long long)row * g.ldc_f32 + col0;
#pragma unroll
                for (int j = 0; j < 32; j++) if (j < nc) atomicAdd(dst + j, v[j]);
            } else if (g.c_f32) {
                float* dst = g.c_f32 + (long long)row * g.ldc_f32 + col0;
                if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (g.accumulate) { const float4 p = *reinterpret_cast<const float4*>(dst + j); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                        *reinterpret_cast<float4*>(dst + j) = o;
                    }
                } else {
                    for (int j = 0; j < nc; j++) dst[j] = g.accumulate ? dst[j] + v[j] : v[j];
                }
            }
            if (g.c_bf16 && g.bf16_split) {
                __nv_bfloat16* dst = g.c_bf16 + (long long)row * g.ldc_bf16 + col0;
                if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (g.N & 7) == 0) {
                    // 16-byte stores: 8 columns of the hi segment, of the lo segment and of the second hi segment at a time
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 uh, ul;
                        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&uh);
                        __nv_bfloat162* hl = reinterpret_cast<__nv_bfloat162*>(&ul);
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const __nv_bfloat162 hi = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                            const float2 hf = __bfloat1622float2(hi);
                            hh[t] = hi;
                            hl[t] = __floats2bfloat162_rn(v[j + 2 * t] - hf.x, v[j + 2 * t + 1] - hf.y);
                        }
                        *reinterpret_cast<uint4*>(dst + j) = uh;
                        *reinterpret_cast<uint4*>(dst + g.N + j) = ul;
                        *reinterpret_cast<uint4*>(dst + 2 * g.N + j) = uh;
                    }
                } else {
                    for (int j = 0; j < nc; j++) {
                        const __nv_bfloat16 hi = __float2bfloat16(v[j]);
                        const __nv_bfloat16 lo = __float2bfloat16(v[j] - __bfloat162float(hi));
                        dst[j] = hi; dst[g.N + j] = lo; dst[2 * g.N + j] = hi;
                    }
                }
            } else if (g.c_bf16) {
                __nv_bfloat16* dst = g.c_bf16 + (long long)row * g.ldc_bf16 + col0;
                if (nc == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int t = 0; t < 4; t++) h[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                        *reinterpret_cast<uint4*>(dst + j) = u;
                    }
                } else {
                    for (int j = 0; j < nc; j++) dst[j] = __float2bfloat16(v[j]);
                }
            }
        }
        if (do_rows && half == 0) {
            uint32_t r[32];
            tmem_ld_row32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)T_BN, r);    // 16 identical columns (+ 16 unused)
            if (row_ok) {
                if (split) atomicAdd(&g.row_sum[row], __uint_as_float(r[0]));
                else g.row_sum[row] = (g.accumulate ? g.row_sum[row] : 0.0f) + __uint_as_float(r[0]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T_TMEM_COLS) : "memory");
}

// row-major bf16 matrix [rows, cols] (row stride ld elements) -> 2-D tensor map with an (inner x outer) box, 128B swizzle
static int make_map2(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_inner, int box_outer)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// ---- small kernels ------------------------------------------------------------------------------------------------------

// fp32 [rows, cols] -> bf16 [rows, cols_pad] (zero padding), 8 output elements per thread
__global__ void cast_pad_kernel(const float* __restrict__ src, long long rows, int cols, long long lds, __nv_bfloat16* __restrict__ dst, int cols_pad)
{
    const int groups = cols_pad >> 3;
    const long long total = rows * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i - r * groups) * 8;
        const float* s = src + r * lds + c;
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const float a = c + 2 * t < cols ? s[2 * t] : 0.0f, b = c + 2 * t + 1 < cols ? s[2 * t + 1] : 0.0f;
            h[t] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4*>(dst + r * cols_pad + c) = u;
    }
}

// fp32 [rows, cols] -> bf16 [rows, 3 * cols_pad]: the value split into hi = bf16(x) and lo = bf16(x - hi) and laid out as
// [hi | lo | hi] (order 0, activations) or [hi | hi | lo] (order 1, weights), so that ONE bf16 GEMM over the concatenated K
// computes x_hi w_hi + x_lo w_hi + x_hi w_lo = x w to ~2^-16 relative — used for the first layer, whose inputs are raw
// observations (PM indices up to P + 1 next to sizes in [0, 1]): plain bf16 operands lose the low bits that matter there.
__global__ void cast_split_kernel(const float* __restrict__ src, long long rows, int cols, long long lds, __nv_bfloat16* __restrict__ dst, int cols_pad,
                                  int order)
{
    const long long total = rows * cols_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols_pad;
        const int c = (int)(i - r * cols_pad);
        const float x = c < cols ? src[r * lds + c] : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16(x);
        const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi));
        __nv_bfloat16* d = dst + r * 3ll * cols_pad + c;
        d[0] = hi;
        d[cols_pad] = order == 0 ? lo : hi;
        d[2 * cols_pad] = order == 0 ? hi : lo;
    }
}

// value head (ppo.py:95-101 last layer, Linear(H, 1)): out[m] = sum_k h[m, k] w[k] + b; warp per row
__global__ void value_head_kernel(const __nv_bfloat16* __restrict__ h, long long rows, int H, const float* __restrict__ w, const float* __restrict__ b,
                                  float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const __nv_bfloat16* hr = h + row * H;
    float acc = 0.0f;
    for (int k = lane * 2; k < H; k += 64) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hr + k));
        acc += f.x * w[k] + f.y * w[k + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc + b[0];
}

// backward of the value head: dz[m, k] = dv[m] w[k] (1 - h[m, k]^2) (bf16), dw[k] += sum_m dv[m] h[m, k], db += sum_m dv[m].
// Block = 64 rows; thread t owns columns t, t + 256, ...; per-block partial sums leave through atomics.
__global__ void __launch_bounds__(256) value_head_backward_kernel(const __nv_bfloat16* __restrict__ h, long long rows, int H, const float* __restrict__ w,
                                                                  const float* __restrict__ dv, __nv_bfloat16* __restrict__ dz, float* __restrict__ dw,
                                                                  float* __restrict__ db)
{
    const long long r0 = (long long)blockIdx.x * 64;
    const long long r1 = min(rows, r0 + 64);
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        const float wk = w[k];
        float acc = 0.0f;
        for (long long r = r0; r < r1; r++) {
            const float hv = __bfloat162float(h[r * H + k]), d = dv[r];
            acc += d * hv;
            dz[r * H + k] = __float2bfloat16(d * wk * (1.0f - hv * hv));
        }
        atomicAdd(&dw[k], acc);
    }
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (long long r = r0; r < r1; r++) s += dv[r];
        atomicAdd(db, s);
    }
}

// Per-sample loss pieces of ppo.py:259-282 and their derivatives.  With r = exp(new_logprob - old_logprob):
//   policy   L_i = max(-r A, -clamp(r, 1 - eps, 1 + eps) A)                 dL/dlogp = -A r unless the clamped branch is the max
//   value    L_i = 0.5 max((v - R)^2, (v_old + clamp(v - v_old, +-eps) - R)^2)  (or the unclipped square)
//   entropy  -ent_coef H_i
// everything divided by n_total (the means over the global minibatch).  c_lp / c_v are the derivatives w.r.t. the sample's summed
// log-prob and its value; sums[0] += sum of log-ratios (KL estimate, ppo.py:263), sums[1] += loss, both in fp64.
__global__ void ppo_loss_kernel(const float* __restrict__ new_lp, const float* __restrict__ old_lp, const float* __restrict__ adv,
                                const float* __restrict__ ent, const float* __restrict__ v, const float* __restrict__ v_old,
                                const float* __restrict__ ret, long long n, float eps, float ent_coef, float vf_coef, int vf_clip, float inv_n,
                                float* __restrict__ c_lp, float* __restrict__ c_v, double* __restrict__ sums)
{
    __shared__ double sh[2][8];
    double s_lr = 0.0, s_loss = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float lr = new_lp[i] - old_lp[i];
        const float r = __expf(lr), A = adv[i];
        const float rc = fminf(fmaxf(r, 1.0f - eps), 1.0f + eps);
        const float s1 = -r * A, s2 = -rc * A;
        const bool inside = r >= 1.0f - eps && r <= 1.0f + eps;
        c_lp[i] = (inside || s1 > s2) ? -A * r * inv_n : 0.0f;
        const float d = v[i] - ret[i];
        float lv = d * d, gv = 2.0f * d;
        if (vf_clip) {
            const float dvv = v[i] - v_old[i];
            const float vc = v_old[i] + fminf(fmaxf(dvv, -eps), eps);
            const float dc = vc - ret[i], lc = dc * dc;
            const float gc = (dvv > -eps && dvv < eps) ? 2.0f * dc : 0.0f;
            if (lc > lv) { lv = lc; gv = gc; }
            else if (lc == lv) gv = 0.5f * (gv + gc);                      // torch.max splits the gradient between equal arguments
        }
        c_v[i] = vf_coef * 0.5f * gv * inv_n;
        s_lr += (double)lr;
        s_loss += ((double)fmaxf(s1, s2) - (double)ent_coef * (double)ent[i] + (double)vf_coef * 0.5 * (double)lv) * (double)inv_n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s_lr += __shfl_xor_sync(0xffffffffu, s_lr, o); s_loss += __shfl_xor_sync(0xffffffffu, s_loss, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = s_lr; sh[1][warp] = s_loss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) { a += sh[0][wq]; b += sh[1][wq]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

static bool attr_done_train(bool mark)
{
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (mark) done[dev] = true;
    return done[dev];
}

}  // namespace vmgym_train

extern "C" int vmgym_tc_gemm(const void* d_a, int32_t a_mn, int64_t lda, const void* d_b, int32_t b_mn, int64_t ldb, int64_t M, int64_t N,
                             int64_t K, const float* d_bias, int32_t act, const void* d_mul_y, int64_t ldy, float* d_c_f32, int64_t ldc_f32,
                             int32_t accumulate, void* d_c_bf16, int64_t ldc_bf16, float* d_row_sum, void* stream)
{
    using namespace vmgym_train;
    if (!d_a || !d_b || (!d_c_f32 && !d_c_bf16 && !d_row_sum) || M < 0 || N < 0 || K <= 0) {
        vmgym_internal_set_error("vmgym_tc_gemm: null operand");
        return VMGYM_EINVAL;
    }
    if (M == 0 || N == 0) return VMGYM_OK;
    const int act_fn = act & 3, f32_pre = (act >> 2) & 1, bf16_split = (act >> 3) & 1;
    if ((lda & 7) || (ldb & 7) || ((uintptr_t)d_a & 15) || ((uintptr_t)d_b & 15) || act_fn == 3 || (act >> 4)) {
        vmgym_internal_set_error("vmgym_tc_gemm: operand row strides must be multiples of 8 elements and bases 16-byte aligned (TMA)");
        return VMGYM_EINVAL;
    }
    int n_sm = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const long long m_tiles = (M + T_BM - 1) / T_BM;
    int tile_n = (((N + T_BN - 1) / T_BN) * m_tiles < n_sm && N > 128) ? 128 : T_BN;   // narrower tiles when 256-wide ones leave SMs idle
    static const int force_tile = getenv("VMGYM_TC_TILE_N") ? atoi(getenv("VMGYM_TC_TILE_N")) : 0;      // A/B experiments
    if (force_tile == 128 || force_tile == 256) tile_n = force_tile;
    CUtensorMap map_a, map_b;
    // K-major operand: tensor [rows, K], box = 64 (K) x tile rows.  MN-major operand: tensor [K, rows], box = 64 (rows) x 64 (K).
    const int ra = a_mn ? make_map2(&map_a, d_a, K, M, lda, 64, 64) : make_map2(&map_a, d_a, M, K, lda, T_BK, T_BM);
    const int rb = b_mn ? make_map2(&map_b, d_b, K, N, ldb, 64, 64) : make_map2(&map_b, d_b, N, K, ldb, T_BK, tile_n);
    if (ra || rb) {
        vmgym_internal_set_error("vmgym_tc_gemm: cuTensorMapEncodeTiled failed");
        return VMGYM_ECUDA;
    }
    if (!attr_done_train(false)) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES);
        if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
        attr_done_train(true);
    }
    GemmArgs g;
    g.M = (int)M; g.N = (int)N; g.K = (int)K; g.a_mn = a_mn ? 1 : 0; g.b_mn = b_mn ? 1 : 0;
    g.bias = d_bias; g.act = act_fn; g.f32_pre = f32_pre; g.bf16_split = bf16_split; g.mul_y = (const __nv_bfloat16*)d_mul_y; g.ldy = ldy;
    g.c_f32 = d_c_f32; g.ldc_f32 = ldc_f32; g.accumulate = accumulate;
    g.c_bf16 = (__nv_bfloat16*)d_c_bf16; g.ldc_bf16 = ldc_bf16; g.row_sum = d_row_sum;
    // split-K for accumulating fp32 outputs with too few tiles to fill the machine (dW = dz^T a over tens of thousands of samples
    // into a 512 x 512 matrix is 8 tiles): slices of the sample axis on gridDim.z, partial tiles added atomically
    g.tile_n = tile_n;
    const long long tiles = ((N + tile_n - 1) / tile_n) * m_tiles;
    const long long k_blocks = (K + T_BK - 1) / T_BK;
    g.k_splits = 1;
    if (accumulate && d_c_f32 && !d_c_bf16 && !d_mul_y && !d_bias && act == 0 && tiles < n_sm && k_blocks >= 16) {   // (act == 0: no flags either)
        long long sp = (2ll * n_sm + tiles - 1) / tiles;
        if (sp > k_blocks / 8) sp = k_blocks / 8;                       // at least 8 k-blocks per slice
        if (sp > 1) {
            const long long per = (k_blocks + sp - 1) / sp;
            sp = (k_blocks + per - 1) / per;                            // no empty slice
        }
        g.k_splits = (int)(sp < 1 ? 1 : sp);
    }
    dim3 grid((unsigned)((N + tile_n - 1) / tile_n), (unsigned)m_tiles, (unsigned)g.k_splits);
    tc_gemm_kernel<<<grid, T_THREADS, T_SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_b, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
    return VMGYM_OK;
}

extern "C" int vmgym_cast_pad_bf16(const float* d_src, int64_t rows, int64_t cols, int64_t lds, void* d_dst_bf16, int64_t cols_pad, void* stream)
{
    if (!d_src || !d_dst_bf16 || rows < 0 || cols < 1 || cols_pad < cols || (cols_pad & 7) || ((uintptr_t)d_dst_bf16 & 15)) {
        vmgym_internal_set_error("vmgym_cast_pad_bf16: bad arguments (cols_pad must be a multiple of 8 >= cols)");
        return VMGYM_EINVAL;
    }
    if (rows == 0) return VMGYM_OK;
    const long long total = rows * (cols_pad >> 3);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    vmgym_train::cast_pad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_src, rows, (int)cols, lds, (__nv_bfloat16*)d_dst_bf16, (int)cols_pad);
    return cudaGetLastError() == cudaSuccess ? VMGYM_OK : VMGYM_ECUDA;
}

extern "C" int vmgym_cast_split_bf16(const float* d_src, int64_t rows, int64_t cols, int64_t lds, void* d_dst_bf16, int64_t cols_pad, int32_t order,
                                     void* stream)
{
    if (!d_src || !d_dst_bf16 || rows < 0 || cols < 1 || cols_pad < cols || (cols_pad & 7) || ((uintptr_t)d_dst_bf16 & 15) || (order != 0 && order != 1)) {
        vmgym_internal_set_error("vmgym_cast_split_bf16: bad arguments (cols_pad must be a multiple of 8 >= cols, order 0 or 1)");
        return VMGYM_EINVAL;
    }
    if (rows == 0) return VMGYM_OK;
    long long blocks = (rows * cols_pad + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    vmgym_train::cast_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_src, rows, (int)cols, lds, (__nv_bfloat16*)d_dst_bf16,
                                                                                       (int)cols_pad, order);
    return cudaGetLastError() == cudaSuccess ? VMGYM_OK : VMGYM_ECUDA;
}

extern "C" int vmgym_value_head(const void* d_h_bf16, int64_t rows, int32_t hidden, const float* d_w, const float* d_b, float* d_out, void* stream)
{
    if (!d_h_bf16 || !d_w || !d_b || !d_out || rows < 0 || hidden < 2 || (hidden & 1)) {
        vmgym_internal_set_error("vmgym_value_head: bad arguments");
        return VMGYM_EINVAL;
    }
    if (rows == 0) return VMGYM_OK;
    const long long blocks = (rows * 32 + 255) / 256;
    vmgym_train::value_head_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)d_h_bf16, rows, hidden, d_w, d_b, d_out);
    return cudaGetLastError() == cudaSuccess ? VMGYM_OK : VMGYM_ECUDA;
}

extern "C" int vmgym_value_head_backward(const void* d_h_bf16, int64_t rows, int32_t hidden, const float* d_w, const float* d_dv, void* d_dz_bf16,
                                         float* d_dw, float* d_db, void* stream)
{
    if (!d_h_bf16 || !d_w || !d_dv || !d_dz_bf16 || !d_dw || !d_db || rows < 0 || hidden < 1) {
        vmgym_internal_set_error("vmgym_value_head_backward: bad arguments");
        return VMGYM_EINVAL;
    }
    if (rows == 0) return VMGYM_OK;
    const long long blocks = (rows + 63) / 64;
    vmgym_train::value_head_backward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)d_h_bf16, rows, hidden, d_w, d_dv,
                                                                                             (__nv_bfloat16*)d_dz_bf16, d_dw, d_db);
    return cudaGetLastError() == cudaSuccess ? VMGYM_OK : VMGYM_ECUDA;
}

extern "C" int vmgym_ppo_loss(const float* d_new_logprob, const float* d_old_logprob, const float* d_adv, const float* d_entropy, const float* d_value,
                              const float* d_old_value, const float* d_return, int64_t n, float eps_clip, float ent_coef, float vf_coef,
                              int32_t vf_loss_clip, float inv_n_total, float* d_c_logprob, float* d_c_value, double* d_sums, void* stream)
{
    if (!d_new_logprob || !d_old_logprob || !d_adv || !d_entropy || !d_value || !d_old_value || !d_return || !d_c_logprob || !d_c_value || !d_sums || n < 0) {
        vmgym_internal_set_error("vmgym_ppo_loss: null operand");
        return VMGYM_EINVAL;
    }
    if (n == 0) return VMGYM_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    vmgym_train::ppo_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_new_logprob, d_old_logprob, d_adv, d_entropy, d_value, d_old_value,
                                                                                     d_return, n, eps_clip, ent_coef, vf_coef, vf_loss_clip, inv_n_total,
                                                                                     d_c_logprob, d_c_value, d_sums);
    return cudaGetLastError() == cudaSuccess ? VMGYM_OK : VMGYM_ECUDA;
}
