// ccx_umma.cuh — thin inline-PTX layer over the sm_100a tensor-core path used by the net kernels:
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), TMEM alloc / ld, mbarrier completion, and the
// shared-memory "core matrix" operand layout (SWIZZLE_NONE, K-major) both operands use.
//
// Operand layout (K-major, no swizzle; CUTLASS "INTERLEAVE": ((8,n),2):((1,SBO),LBO) in 16-byte units):
//   a [rows x K] bf16 operand is stored as 8-row x 16-byte core matrices (128 contiguous bytes each);
//   byte offset of element (r, k) = (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2
//   with LBO = 128 (the K/8 core matrices of an 8-row group are contiguous) and SBO = (K/8)*128.
// The A operand is [M=128 x K] activations; the B operand is the weight matrix stored TRANSPOSED as
// [N x K] (row n = output channel) in the same scheme.  One MMA consumes K = 16 (two core matrices per row
// group); k-step s starts at byte offset s*256.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint32_t op_offset(int r, int k, int K)        // bytes
{
    return (uint32_t)((r >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__host__ __device__ constexpr uint32_t op_bytes(int rows, int K) { return (uint32_t)(rows * K * 2); }

// shared-memory matrix descriptor (SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    return d;                                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// instruction descriptor: (bf16 | fp16) x same -> f32, both operands K-major, M = 128.
// a_format / b_format: 1 = BF16, 0 = F16 (cute::UMMA::F16F32Format)
__host__ __device__ constexpr uint32_t make_idesc(int N, bool fp16 = false)
{
    const uint32_t fmt = fp16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate)
{
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// D[128 x N] (+)= A[128 x K] * B^T[N x K]; issued by ONE thread.  a_base/b_base: smem byte addresses of
// operands laid out with op_offset(., ., K_a) / op_offset(., ., K_b); k0 = first K column of each to use.
__device__ __forceinline__ void gemm_issue(uint32_t tmem_d, uint32_t a_base, int Ka, int ka0, uint32_t b_base, int Kb, int kb0,
                                           int K, int N, bool accumulate_first, bool fp16 = false)
{
    const uint32_t idesc = make_idesc(N, fp16);
    for (int s = 0; s < K / 16; s++) {
        uint64_t ad = make_desc(a_base + (uint32_t)((ka0 >> 3) + 2 * s) * 128u, 128u, (uint32_t)(Ka >> 3) * 128u);
        uint64_t bd = make_desc(b_base + (uint32_t)((kb0 >> 3) + 2 * s) * 128u, 128u, (uint32_t)(Kb >> 3) * 128u);
        mma_bf16(tmem_d, ad, bd, idesc, accumulate_first || s > 0);
    }
}

// Descriptor arithmetic for issue loops: only the 14-bit start-address field changes between the MMAs of a
// layer, so a layer keeps (lo, hi) words of its operands' base descriptors and adds byte offsets to `lo`
// (one integer add per MMA instead of rebuilding the descriptor; shared memory < 256 KB so no carry).
struct DescBase { uint32_t lo, hi; };
__device__ __forceinline__ DescBase desc_base(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    DescBase d;
    d.lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    d.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return d;
}
__device__ __forceinline__ uint64_t desc_at(DescBase b, uint32_t byte_off)
{
    return ((uint64_t)b.hi << 32) | (uint64_t)(b.lo + (byte_off >> 4));
}
// one lane of a fully converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// K-major standard-layout operands (op_offset): D (+)= A[., ka0 .. ka0+K) * B[., kb0 .. kb0+K)^T
template <int K>
__device__ __forceinline__ void gemm_issue_d(uint32_t tmem_d, DescBase a, int ka0, DescBase b, int kb0, uint32_t idesc, bool accumulate_first)
{
#pragma unroll
    for (int s = 0; s < K / 16; s++)
        mma_bf16(tmem_d, desc_at(a, (uint32_t)((ka0 >> 3) + 2 * s) * 128u), desc_at(b, (uint32_t)((kb0 >> 3) + 2 * s) * 128u), idesc,
                 accumulate_first || s > 0);
}

__device__ __forceinline__ void commit(uint64_t *bar)     // arrives on `bar` when all prior MMAs of this thread are done
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- bulk asynchronous copy global -> shared (TMA engine, no tensor map): one thread issues, completion is
// signalled on an mbarrier as a transaction-byte count.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM: called by one full warp; writes the base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    __syncwarp();                            // .sync.aligned: the warp must be converged (it may come out of a spin wait)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    __syncwarp();
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// (row-contiguous A operands — element (r, k) at (k/8)*LBO + r*16 + (k%8)*2, SBO = 128 — are addressed by the kernels through
// desc_base / desc_at with a start address that already includes the row shift: see k_net_trunk_tc4's conv B)

// ---- A operand from TMEM (tcgen05.mma "ts" form): row r of A lives in TMEM lane r, two consecutive K elements
// per 32-bit column (element 2j in the low half of column j); one MMA consumes K = 16 = 8 columns.
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate)
{
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
template <int K>
__device__ __forceinline__ void gemm_issue_ts(uint32_t tmem_d, uint32_t tmem_a, DescBase b, int kb0, uint32_t idesc, bool accumulate_first)
{
#pragma unroll
    for (int s = 0; s < K / 16; s++)
        mma_bf16_ts(tmem_d, tmem_a + 8u * s, desc_at(b, (uint32_t)((kb0 >> 3) + 2 * s) * 128u), idesc, accumulate_first || s > 0);
}

// 32 lanes x N consecutive 32-bit columns, thread i of the warp writes lane (base + i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
                 " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace umma
