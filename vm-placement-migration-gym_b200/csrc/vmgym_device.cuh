// vmgym_device.cuh — device-side building blocks shared by the env kernels (sm_100a).
//
// Layout notes: one env = one contiguous record in HBM (vmgym_layout); a warp owns one
// env at a time, stages the record into shared memory with one bulk-async copy (cp.async.bulk, the 1-D TMA
// path), works on it there, and writes it back with one bulk-async store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vmgym.h"

namespace vmgym {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SIZE_TABLE = 128;   // size codes are hundredths 0..100 (7 bits)
constexpr int SVC_BRACKETS = 64;  // the service inverse-CDF search starts from the bracket of the top 6 bits of u
constexpr int ARR_CDF_SMEM = 64;  // arrival inverse-CDF thresholds kept in shared memory when the table is this small

struct DevLayout {
    int P, V, A, Pp, Vp, D;
    int off_mem, off_rem, off_place, off_cpuc, off_memc, off_cap, off_scal, rec_bytes;
    // per-warp shared memory (byte offsets from the warp's base)
    int sm_cpu32, sm_mem32, sm_act, sm_tmp, sm_fit, sm_prop, sm_stats, sm_bar, sm_team, sm_stride;
    int sm_kl;       // scratch of the `kl` reward (compacted size codes of the existing VMs): sm_tmp, or the agent's float32 view (dead by then)
    // CTA-wide shared memory
    int sm_tables;   // bytes reserved in front of the per-warp regions
    int svc_cdf_smem;  // service inverse-CDF entries kept in shared memory (0 = read the global table)
};

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copies (TMA without a tensor map; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// numpy float64 add-reduce order (np.sum / np.mean / np.var on a contiguous array): 0 + pairwise(a), with
// 8 interleaved accumulators per block of <= 128 and recursive halving above (third-party: numpy
// loops_utils.h.src DOUBLE_pairwise_sum).  Every calling lane executes the same
// sequence; the adds are strict IEEE fp64 (this translation unit is compiled with -fmad=false).
// ---------------------------------------------------------------------------------------------------
// Element source of a reduction: a[i] or sz[codes[i]], optionally as squared deviation from `sub`.
struct SumSrc {
    const double* a;
    const uint8_t* codes;
    const double* sz;
    double sub;
    int sq;
};
__device__ __forceinline__ double sum_elem(const SumSrc& s, int i)
{
    const double x = s.codes ? s.sz[s.codes[i]] : s.a[i];
    if (!s.sq) return x;
    const double d = x - s.sub;
    return d * d;
}

__device__ __forceinline__ double np_leaf_sum(const SumSrc& f, int s, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += sum_elem(f, s + i);
        return res;
    }
    double r0 = sum_elem(f, s), r1 = sum_elem(f, s + 1), r2 = sum_elem(f, s + 2), r3 = sum_elem(f, s + 3),
           r4 = sum_elem(f, s + 4), r5 = sum_elem(f, s + 5), r6 = sum_elem(f, s + 6), r7 = sum_elem(f, s + 7);
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
        r0 += sum_elem(f, s + i); r1 += sum_elem(f, s + i + 1); r2 += sum_elem(f, s + i + 2); r3 += sum_elem(f, s + i + 3);
        r4 += sum_elem(f, s + i + 4); r5 += sum_elem(f, s + i + 5); r6 += sum_elem(f, s + i + 6); r7 += sum_elem(f, s + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; i++) res += sum_elem(f, s + i);
    return res;
}

static __device__ __noinline__ double np_sum(const SumSrc f, int n)
{
    if (n <= 128) return np_leaf_sum(f, 0, n);
    // explicit post-order walk of the halving recursion (depth <= 10 for n <= 65536)
    int st_s[12], st_n[12], st_phase[12];
    double st_left[12];
    int sp = 0;
    st_s[0] = 0; st_n[0] = n; st_phase[0] = 0; st_left[0] = 0.;
    double ret = 0.;
    while (sp >= 0) {
        const int s = st_s[sp], m = st_n[sp];
        if (m <= 128) { ret = np_leaf_sum(f, s, m); sp--; continue; }
        int n2 = m / 2;
        n2 -= n2 % 8;
        if (st_phase[sp] == 0) {
            st_phase[sp] = 1;
            sp++; st_s[sp] = s; st_n[sp] = n2; st_phase[sp] = 0;
        } else if (st_phase[sp] == 1) {
            st_left[sp] = ret; st_phase[sp] = 2;
            sp++; st_s[sp] = s + n2; st_n[sp] = m - n2; st_phase[sp] = 0;
        } else {
            ret = st_left[sp] + ret;
            sp--;
        }
    }
    return ret;
}

// k / 100.0 (correctly rounded, == np.around(u, 2) for the size code k) without the fp64 division routine: one
// Newton-style correction of k * 0.01 with exact fma residuals.  Checked against k / 100.0 for every k < 256
// (tests/test_abi.py::test_code_to_f64_table restates it with exact rationals; the parity tests cover it on the GPU).
__device__ __forceinline__ double code_to_f64(int k)
{
    const double kd = (double)k;
    const double q0 = __dmul_rn(kd, 0.01);
    const double r = __fma_rn(-q0, 100.0, kd);
    return __fma_rn(r, 0.01, q0);
}


// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG; Salmon et al. SC'11).  Used in VMGYM_TRACE_PHILOX mode.
// ---------------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0, hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

}  // namespace vmgym
