// vmgym_tc.cuh — Blackwell (sm_100a) tensor-core primitives shared by the GEMM / fused-head kernels (vmgym_gemm.cu) and the PPO
// update kernels (vmgym_train.cu): mbarrier, TMA (cp.async.bulk.tensor), tcgen05.mma / commit / ld wrappers, UMMA shared-memory
// and instruction descriptors, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmgym_gemm {

constexpr int BM = 128, BN = 128, BK = 64;      // BK * sizeof(bf16) = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = (BM + BN) * BK * 2; // 32 KiB
constexpr int TMEM_COLS = 128;                  // fp32 accumulator: 128 lanes x 128 columns
constexpr int THREADS = 192;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 1024 /* barriers, TMEM ptr, bias tile */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart):
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO = 1024>>4 | version 1 (sm_100) | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9 / 10-12 = 1), K-major both, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One 64-wide K block (4 MMAs of K = 16) of K-major SW128 operands as a single instruction group: the descriptors differ only in
// their low words (start address >> 4, + 2 per 32-byte K step), so the four MMAs need three 32-bit adds and no per-MMA descriptor
// rebuild.  `a_lo` / `b_lo` = low words of umma_desc_sw128(address of the K block), UMMA_DESC_HI the constant high word.
// Measured (profiles/r2_fused_head.md): rebuilding 64-bit descriptors per MMA in a lane-0 branch cost ~100 SASS instructions and
// ~675 cycles per K block — 3x the MMA time — which made the single issuing thread the bottleneck of the fused head.
constexpr uint32_t UMMA_DESC_HI = 64u | (1u << 14) | (2u << 29);        // SBO = 1024 B, descriptor version 1, 128 B swizzle
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_bf16_k64(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate_first)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, db0, da1, db1, da2, db2, da3, db3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "add.u32 a1, %1, 2;\n\t"
        "add.u32 b1, %2, 2;\n\t"
        "add.u32 a2, %1, 4;\n\t"
        "add.u32 b2, %2, 4;\n\t"
        "add.u32 a3, %1, 6;\n\t"
        "add.u32 b3, %2, 6;\n\t"
        "mov.b64 da0, {%1, %5};\n\t"
        "mov.b64 db0, {%2, %5};\n\t"
        "mov.b64 da1, {a1, %5};\n\t"
        "mov.b64 db1, {b1, %5};\n\t"
        "mov.b64 da2, {a2, %5};\n\t"
        "mov.b64 db2, {b2, %5};\n\t"
        "mov.b64 da3, {a3, %5};\n\t"
        "mov.b64 db3, {b3, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da2, db2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da3, db3, %3, q;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(UMMA_DESC_HI)
        : "memory");
}
// The same for any operand layout: `a_step` / `b_step` = descriptor low-word increment per K = 16 step (2 for a K-major SW128 tile,
// 128 for an MN-major one: two 8-row groups of 1024 B), `hi_a` / `hi_b` the operands' constant descriptor high words.
__device__ __forceinline__ void umma_bf16_k64_ex(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_step, uint32_t b_step,
                                                 uint32_t hi_a, uint32_t hi_b, uint32_t idesc, uint32_t accumulate_first)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, db0, da1, db1, da2, db2, da3, db3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "add.u32 a1, %1, %7;\n\t"
        "add.u32 b1, %2, %8;\n\t"
        "add.u32 a2, a1, %7;\n\t"
        "add.u32 b2, b1, %8;\n\t"
        "add.u32 a3, a2, %7;\n\t"
        "add.u32 b3, b2, %8;\n\t"
        "mov.b64 da0, {%1, %5};\n\t"
        "mov.b64 db0, {%2, %6};\n\t"
        "mov.b64 da1, {a1, %5};\n\t"
        "mov.b64 db1, {b1, %6};\n\t"
        "mov.b64 da2, {a2, %5};\n\t"
        "mov.b64 db2, {b2, %6};\n\t"
        "mov.b64 da3, {a3, %5};\n\t"
        "mov.b64 db3, {b3, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da2, db2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da3, db3, %3, q;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(hi_a), "r"(hi_b), "r"(a_step), "r"(b_step)
        : "memory");
}
// true in exactly one lane of a converged warp (the lane that issues TMA / MMA / commit)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive accumulator columns of this thread's row (TMEM lane) into registers
__device__ __forceinline__ void tmem_ld_row32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// row-major [rows, K] bf16 matrix -> 2-D tensor map with a (BK x box_rows) box, 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace vmgym_gemm
