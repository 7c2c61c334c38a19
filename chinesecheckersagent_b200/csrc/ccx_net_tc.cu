// ccx_net_tc.cu — bf16 tensor-core (tcgen05 + TMEM) path of the policy/value net (model.py:58-145).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ccx_device.cuh"
#include "ccx_internal.h"
#include "ccx_umma.cuh"
#include <new>
#include <cstdlib>

// ---- self-test of the UMMA plumbing: D[128 x N] = A[128 x K] * Bt[N x K]^T -------------------------------
__global__ void __launch_bounds__(128)
k_umma_selftest(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ Bt, int K, int N, float *__restrict__ D)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sa = smem, *sb = smem + umma::op_bytes(128, K);
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 128 * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16 *>(sa + umma::op_offset(r, k, K)) = A[i];
    }
    for (int i = t; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16 *>(sb + umma::op_offset(r, k, K)) = Bt[i];
    }
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        umma::gemm_issue(tmem, umma::smem_u32(sa), K, 0, umma::smem_u32(sb), K, 0, K, N, false);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; j++) if (c + j < N) D[t * N + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 64);
}

// Second self-test: the A operand in the ROW-CONTIGUOUS layout the 3x3 conv uses — element (r, k) at
// (k/8)*LBO + r*16 + (k%8)*2 with LBO = rows*16, SBO = 128 — and a descriptor whose start address is moved by
// `shift` rows (16 B each, NOT a multiple of the 128-byte core matrix): D = A[shift .. shift+127] * Bt^T.
__global__ void __launch_bounds__(128)
k_umma_selftest_rows(const __nv_bfloat16 *__restrict__ A, int rows, int shift, const __nv_bfloat16 *__restrict__ Bt, int K, int N,
                     float *__restrict__ D)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sa = smem, *sb = smem + (size_t)rows * K * 2;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < rows * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16 *>(sa + (k >> 3) * rows * 16 + r * 16 + (k & 7) * 2) = A[i];
    }
    for (int i = t; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16 *>(sb + umma::op_offset(r, k, K)) = Bt[i];
    }
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        const uint32_t idesc = umma::make_idesc(N, false);
        for (int s = 0; s < K / 16; s++) {
            uint64_t ad = umma::make_desc(umma::smem_u32(sa) + (uint32_t)(2 * s) * rows * 16u + (uint32_t)shift * 16u, (uint32_t)rows * 16u, 128u);
            uint64_t bd = umma::make_desc(umma::smem_u32(sb) + (uint32_t)(2 * s) * 128u, 128u, (uint32_t)(K >> 3) * 128u);
            umma::mma_bf16(tmem, ad, bd, idesc, s > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; j++) if (c + j < N) D[t * N + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 64);
}

// Third self-test: the A operand read from TMEM (packed 16-bit pairs written with tcgen05.st): D = A * Bt^T, K = 64.
__global__ void __launch_bounds__(128)
k_umma_selftest_ts(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ Bt, int N, float *__restrict__ D)
{
    constexpr int K = 64;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16 *>(smem + umma::op_offset(r, k, K)) = Bt[i];
    }
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 32; j++) w[j] = reinterpret_cast<const uint32_t *>(A + (size_t)t * K)[j];   // elements 2j (low), 2j+1 (high)
        umma::tmem_st32(trow + 64, w);
        umma::tmem_wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
            umma::gemm_issue_ts<K>(tmem, tmem + 64, umma::desc_base(umma::smem_u32(smem), 128u, K / 8 * 128u), 0, umma::make_idesc(N, false), false);
            umma::commit(&bar);
        }
        __syncwarp();
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        umma::tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 16; j++) D[t * N + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 128);
}

extern "C" {

int ccx_debug_umma_gemm_ts(ccx_handle *h, const void *A, const void *Bt, int32_t N, float *D)
{
    if (!h || !A || !Bt || !D || N % 16 || N < 16 || N > 64) return CCX_ERR_ARG;
    size_t smem = umma::op_bytes(N, 64);
    k_umma_selftest_ts<<<1, 128, smem, h->stream>>>((const __nv_bfloat16 *)A, (const __nv_bfloat16 *)Bt, N, D);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_debug_umma_gemm_rows(ccx_handle *h, const void *A, int32_t rows, int32_t shift, const void *Bt, int32_t K, int32_t N, float *D)
{
    if (!h || !A || !Bt || !D || K % 16 || K < 16 || K > 128 || N % 16 || N < 16 || N > 64 || rows % 8 || shift < 0 || shift + 128 > rows)
        return CCX_ERR_ARG;
    size_t smem = (size_t)rows * K * 2 + umma::op_bytes(N, K);
    if (smem > 200 * 1024) return CCX_ERR_ARG;
    CCX_CUDA(h, cudaFuncSetAttribute(k_umma_selftest_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest_rows<<<1, 128, smem, h->stream>>>((const __nv_bfloat16 *)A, rows, shift, (const __nv_bfloat16 *)Bt, K, N, D);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_debug_umma_gemm(ccx_handle *h, const void *A, const void *Bt, int32_t K, int32_t N, float *D)
{
    if (!h || !A || !Bt || !D || K % 16 || K < 16 || K > 512 || N % 16 || N < 16 || N > 64) return CCX_ERR_ARG;
    size_t smem = umma::op_bytes(128, K) + umma::op_bytes(N, K);
    CCX_CUDA(h, cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<<<1, 128, smem, h->stream>>>((const __nv_bfloat16 *)A, (const __nv_bfloat16 *)Bt, K, N, D);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

}  // extern "C"

// =====================================================================================================
// bf16 tensor-core forward pass.
//
// Kernel 1 (k_net_trunk_tc): persistent CTAs, 128 threads, one 128-row tile = 5 positions x 25 cells
// (+3 idle rows) at a time.  Thread t owns row t for the whole network: it keeps the residual stream of
// its cell in fp32 registers, reads each layer's accumulator row from TMEM (tcgen05.ld), applies bias /
// ReLU / skip, and writes the bf16 operand row of the next layer straight into shared memory in the UMMA
// core-matrix layout — for the 3x3 conv as an im2col scatter into the rows of its 8 neighbours.  One elected
// thread issues the MMAs; completion is tracked with one mbarrier.  Activations never leave the SM.
// Residual-block weights (26 KB bf16 per block, pre-arranged on the host in the operand layout) stream from
// L2 through a double buffer with cp.async, one block ahead of the math.
// Kernel 2 (k_policy_dense_tc): logits = flat(policy conv)[B x 400] * W[400 x 294] as 128-position tiles.
//
// Weight blobs (built by model.py pack_weights_tc):
//   bf16 blob, byte offsets: CONV1 (N64,K64) | HEADS (N32,K64: 16 policy-conv cols, value-conv col, zeros) |
//     9 x [A (N32,K64) | B (N32,K288) | C (N64,K32)] | policy dense: 2 N-halves x [K 0..207 (N160,K208) | K 208..399 (N160,K192)]
//   fp32 blob: conv1_b[64] heads_b[32] 9 x (a_b[32] b_b[32] c_b[64]) pold_b[320] d1_w[25][32] d1_b[32] vh_w[32] vh_b[1]
namespace tcl {
constexpr int W_CONV1 = 0, W_HEADS = 8192, W_BLOCK0 = 12288, W_BLOCK = 26624, W_BA = 0, W_BB = 4096, W_BC = 4096 + 18432;
constexpr int W_POLD = W_BLOCK0 + 9 * W_BLOCK;                 // 251,904
constexpr int POLD_C0 = 160 * 208 * 2, POLD_C1 = 160 * 192 * 2, POLD_HALF = POLD_C0 + POLD_C1;
constexpr int W_TOTAL = W_POLD + 2 * POLD_HALF;                // 507,904 bytes
constexpr int F_CONV1 = 0, F_HEADS = 64, F_BLOCK0 = 96, F_BLOCK = 128, F_POLD = F_BLOCK0 + 9 * F_BLOCK;
constexpr int F_D1W = F_POLD + 320, F_D1B = F_D1W + 800, F_VHW = F_D1B + 32, F_VHB = F_VHW + 32, F_TOTAL = F_VHB + 1;
// shared memory map of the trunk kernel (bytes): 110,784 B so that TWO CTAs are resident per SM
constexpr int S_XA = 0;                                  // [128 x 64] operand: block input / conv1 im2col
constexpr int S_IM0 = S_XA + 128 * 64 * 2;               // [128 x 96] im2col operand of one kernel row (dy) of the 3x3 conv
constexpr int S_IM1 = S_IM0 + 128 * 96 * 2;              //   ... double buffered across the three kernel rows
constexpr int S_M2 = S_IM1 + 128 * 96 * 2;               // [128 x 32] operand: 3x3 conv output
constexpr int S_WC1 = S_M2 + 128 * 32 * 2;               // conv1 weights, resident
constexpr int S_WA = S_WC1 + 8192;                       // streamed: conv A of the current block (then the heads)
constexpr int S_WB = S_WA + 4096;                        // streamed: conv B
constexpr int S_WC = S_WB + 18432;                       // streamed: conv C
constexpr int S_PLANES = S_WC + 4096, S_VALC = S_PLANES + 5 * 343 + 13;
constexpr int S_TOTAL = S_VALC + 128 * 4;
static_assert(S_TOTAL == 110784, "trunk kernel shared memory budget (2 CTAs / SM)");
}  // namespace tcl

struct ccx_net_tc {
    uint8_t *wb = nullptr;      // bf16 operand blob
    float *fb = nullptr;        // fp32 biases + value-head dense
    __nv_bfloat16 *polc = nullptr;   // [cap][400] policy-conv activations between the two kernels (16-bit)
    int64_t cap = 0;
    int fp16 = 0;               // 0 = bf16 operands, 1 = IEEE half operands (same kernels, other instruction descriptor)
};

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// two fp32 -> one 32-bit word of 16-bit operands: bf16 (FP16 = false) or IEEE half (FP16 = true)
template <bool FP16> __device__ __forceinline__ uint32_t pack2(float a, float b)
{
    if (FP16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t *>(&h); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// cooperative 16-byte-granular global -> shared copy as one cp.async group (an empty group if bytes == 0)
__device__ __forceinline__ void refill(uint32_t dst, const uint8_t *src, int bytes, int t)
{
    for (int i = t; i < bytes / 16; i += 128) cp_async16(dst + i * 16, src + i * 16);
    cp_async_commit();
}

template <bool FP16>
__global__ void __launch_bounds__(128, 2)
k_net_trunk_tc(const uint8_t *__restrict__ wb, const float *__restrict__ fb, const uint8_t *__restrict__ planes, int64_t n,
               __nv_bfloat16 *__restrict__ polc, float *__restrict__ value)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar, bar_g;                  // bar: "this layer's MMAs are done"; bar_g: "kernel row 0 of the 3x3 is done"
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t sbase = umma::smem_u32(smem);
    const int64_t n_tiles = (n + 4) / 5;

    // Weight streaming: wA / wB / wC each hold ONE layer and are refilled with the next layer that will use the
    // slot as soon as the layer that was using it has finished (the refill then has two layer-times to land).
    // Every phase issues exactly one cp.async group (possibly empty), so "the group that loaded my weights" is
    // always at least three groups old and cp.async.wait_group 2 in front of each phase is sufficient.
    for (int i = t; i < 8192 / 16; i += 128) cp_async16(sbase + tcl::S_WC1 + i * 16, wb + tcl::W_CONV1 + i * 16);
    refill(sbase + tcl::S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096, t);            // group: conv1 + A0
    refill(sbase + tcl::S_WB, wb + tcl::W_BLOCK0 + tcl::W_BB, 18432, t);           // group: B0
    refill(sbase + tcl::S_WC, wb + tcl::W_BLOCK0 + tcl::W_BC, 4096, t);            // group: C0
    if (t == 0) { umma::mbar_init(&bar, 1); umma::mbar_init(&bar_g, 1); }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes
    uint32_t phase = 0, phase_g = 0;

    const int p_local = t / 25, cell = t % 25, cy = cell / 5, cx = cell % 5;
    const bool row_live = t < 125;

    // one layer = operands ready -> MMA -> completion.  `pre` runs after cp.async.wait and before the barrier.
    auto layer_sync = [&]() {
        cp_async_wait<2>();
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pos0 = tile * 5;
        const int n_pos = (int)min((int64_t)5, n - pos0);
        // ---- stage the input planes (uint8, values 0..6) ------------------------------------------------
        for (int i = t; i < n_pos * 343; i += 128) smem[tcl::S_PLANES + i] = planes[pos0 * 343 + i];
        cp_async_commit();                                   // (empty group: keeps the group count per phase uniform)
        __syncthreads();
        // ---- conv1 operand: im2col of the 3x3 'valid' window, K = 63 (+1 zero) (model.py:62) ---------------
        {
            const bool ok = row_live && p_local < n_pos;
            const uint8_t *pl = smem + tcl::S_PLANES + p_local * 343;
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int kk = c8 * 8 + q;
                    const int tap = kk / 7, ch = kk % 7, dy = tap / 3, dx = tap % 3;
                    v[q] = (ok && kk < 63) ? (float)pl[((cy + dy) * 7 + (cx + dx)) * 7 + ch] : 0.f;
                }
                uint4 o = make_uint4(pack2<FP16>(v[0], v[1]), pack2<FP16>(v[2], v[3]), pack2<FP16>(v[4], v[5]), pack2<FP16>(v[6], v[7]));
                *reinterpret_cast<uint4 *>(smem + tcl::S_XA + umma::op_offset(t, c8 * 8, 64)) = o;
            }
        }
        layer_sync();
        float x[64];                        // residual stream of this row, fp32
        if (t == 0) {
            umma::fence_after_sync();
            umma::gemm_issue(tmem, sbase + tcl::S_XA, 64, 0, sbase + tcl::S_WC1, 64, 0, 64, 64, false, FP16);
            umma::commit(&bar);
        }
        cp_async_commit();                                   // (empty group)
        umma::mbar_wait(&bar, phase); phase ^= 1;
        umma::fence_after_sync();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float v[32];
            umma::tmem_ld32(trow + 32 * h, v);
#pragma unroll
            for (int j = 0; j < 32; j++) x[32 * h + j] = fmaxf(v[j] + __ldg(fb + tcl::F_CONV1 + 32 * h + j), 0.f);
        }
#pragma unroll
        for (int c8 = 0; c8 < 8; c8++)
            *reinterpret_cast<uint4 *>(smem + tcl::S_XA + umma::op_offset(t, c8 * 8, 64)) =
                make_uint4(pack2<FP16>(x[c8 * 8], x[c8 * 8 + 1]), pack2<FP16>(x[c8 * 8 + 2], x[c8 * 8 + 3]),
                           pack2<FP16>(x[c8 * 8 + 4], x[c8 * 8 + 5]), pack2<FP16>(x[c8 * 8 + 6], x[c8 * 8 + 7]));

        // ---- 9 bottleneck residual blocks (model.py:120-145) --------------------------------------------------
        for (int b = 0; b < 9; b++) {
            const uint8_t *wblk = wb + tcl::W_BLOCK0 + b * tcl::W_BLOCK;
            const uint8_t *wnext = wb + tcl::W_BLOCK0 + ((b + 1) % 9) * tcl::W_BLOCK;
            const float *bias = fb + tcl::F_BLOCK0 + b * tcl::F_BLOCK;
            (void)wblk;
            // A: 1x1 conv 64 -> 32, ReLU
            layer_sync();
            if (t == 0) {
                umma::fence_after_sync();
                umma::gemm_issue(tmem, sbase + tcl::S_XA, 64, 0, sbase + tcl::S_WA, 64, 0, 64, 32, false, FP16);
                umma::commit(&bar);
            }
            float bv[32];
#pragma unroll
            for (int j = 0; j < 32; j++) bv[j] = __ldg(bias + j);           // bias loads overlap the MMA
            umma::mbar_wait(&bar, phase); phase ^= 1;
            umma::fence_after_sync();
            // the A slot is free: stream in conv A of the next block, or the heads after the last block
            refill(sbase + tcl::S_WA, b < 8 ? wnext + tcl::W_BA : wb + tcl::W_HEADS, 4096, t);
            uint4 o[4];
            {
                float v[32];
                umma::tmem_ld32(trow, v);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) r[q] = fmaxf(v[c * 8 + q] + bv[c * 8 + q], 0.f);
                    o[c] = make_uint4(pack2<FP16>(r[0], r[1]), pack2<FP16>(r[2], r[3]), pack2<FP16>(r[4], r[5]), pack2<FP16>(r[6], r[7]));
                }
            }
#pragma unroll
            for (int j = 0; j < 32; j++) bv[j] = __ldg(bias + 32 + j);
            // B: 3x3 'same' conv 32 -> 32 as three accumulating K = 96 GEMMs, one per kernel row dy; the im2col
            // operand of a kernel row is scattered by the threads that own the source cells, the zero padding is
            // written by the thread that owns the output cell
#pragma unroll
            for (int g = 0; g < 3; g++) {
                const int dy = g - 1;
                const int im = (g & 1) ? tcl::S_IM1 : tcl::S_IM0;
                if (g == 2) { umma::mbar_wait(&bar_g, phase_g); phase_g ^= 1; umma::fence_after_sync(); }   // kernel row 0 has released IM0
                if (row_live) {
#pragma unroll
                    for (int dxi = 0; dxi < 3; dxi++) {
                        const int dx = dxi - 1;
                        const int oy = cy - dy, ox = cx - dx;               // output cell that reads this cell through tap (dy, dx)
                        if (oy >= 0 && oy <= 4 && ox >= 0 && ox <= 4) {
                            const int orow = p_local * 25 + oy * 5 + ox;
#pragma unroll
                            for (int c = 0; c < 4; c++)
                                *reinterpret_cast<uint4 *>(smem + im + umma::op_offset(orow, dxi * 32 + c * 8, 96)) = o[c];
                        }
                        const int iy = cy + dy, ix = cx + dx;               // source cell of MY output through tap (dy, dx)
                        if (iy < 0 || iy > 4 || ix < 0 || ix > 4) {
#pragma unroll
                            for (int c = 0; c < 4; c++)
                                *reinterpret_cast<uint4 *>(smem + im + umma::op_offset(t, dxi * 32 + c * 8, 96)) = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
                if (g == 0) layer_sync();
                else { umma::fence_async_smem(); umma::fence_before_sync(); __syncthreads(); }
                if (t == 0) {
                    umma::fence_after_sync();
                    umma::gemm_issue(tmem + 32, sbase + im, 96, 0, sbase + tcl::S_WB, 288, 96 * g, 96, 32, g > 0, FP16);
                    if (g == 0) umma::commit(&bar_g);
                    if (g == 2) umma::commit(&bar);
                }
            }
            umma::mbar_wait(&bar, phase); phase ^= 1;
            umma::fence_after_sync();
            refill(sbase + tcl::S_WB, wnext + tcl::W_BB, 18432, t);
            {
                float v[32];
                umma::tmem_ld32(trow + 32, v);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) r[q] = fmaxf(v[c * 8 + q] + bv[c * 8 + q], 0.f);
                    *reinterpret_cast<uint4 *>(smem + tcl::S_M2 + umma::op_offset(t, c * 8, 32)) =
                        make_uint4(pack2<FP16>(r[0], r[1]), pack2<FP16>(r[2], r[3]), pack2<FP16>(r[4], r[5]), pack2<FP16>(r[6], r[7]));
                }
            }
            // C: 1x1 conv 32 -> 64, + skip, ReLU (model.py:137-144)
            layer_sync();
            if (t == 0) {
                umma::fence_after_sync();
                umma::gemm_issue(tmem + 64, sbase + tcl::S_M2, 32, 0, sbase + tcl::S_WC, 32, 0, 32, 64, false, FP16);
                umma::commit(&bar);
            }
            umma::mbar_wait(&bar, phase); phase ^= 1;
            umma::fence_after_sync();
            refill(sbase + tcl::S_WC, wnext + tcl::W_BC, 4096, t);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float v[32];
                umma::tmem_ld32(trow + 64 + 32 * h, v);
#pragma unroll
                for (int j = 0; j < 32; j++) x[32 * h + j] = fmaxf(v[j] + __ldg(bias + 64 + 32 * h + j) + x[32 * h + j], 0.f);
            }
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++)
                *reinterpret_cast<uint4 *>(smem + tcl::S_XA + umma::op_offset(t, c8 * 8, 64)) =
                    make_uint4(pack2<FP16>(x[c8 * 8], x[c8 * 8 + 1]), pack2<FP16>(x[c8 * 8 + 2], x[c8 * 8 + 3]),
                               pack2<FP16>(x[c8 * 8 + 4], x[c8 * 8 + 5]), pack2<FP16>(x[c8 * 8 + 6], x[c8 * 8 + 7]));
        }
        // ---- heads: policy conv 64 -> 16 and value conv 64 -> 1 in one N = 32 GEMM (model.py:91, 108); weights in the A slot
        layer_sync();
        if (t == 0) {
            umma::fence_after_sync();
            umma::gemm_issue(tmem, sbase + tcl::S_XA, 64, 0, sbase + tcl::S_WA, 64, 0, 64, 32, false, FP16);
            umma::commit(&bar);
        }
        umma::mbar_wait(&bar, phase); phase ^= 1;
        umma::fence_after_sync();
        refill(sbase + tcl::S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096, t);          // conv A of block 0 for the next tile
        {
            float v[32];
            umma::tmem_ld32(trow, v);
            float *valc = reinterpret_cast<float *>(smem + tcl::S_VALC);
            valc[t] = fmaxf(v[16] + __ldg(fb + tcl::F_HEADS + 16), 0.f);
            if (row_live && p_local < n_pos) {
                uint4 o[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) r[q] = fmaxf(v[c * 8 + q] + __ldg(fb + tcl::F_HEADS + c * 8 + q), 0.f);
                    o[c] = make_uint4(pack2<FP16>(r[0], r[1]), pack2<FP16>(r[2], r[3]), pack2<FP16>(r[4], r[5]), pack2<FP16>(r[6], r[7]));
                }
                uint4 *dst = reinterpret_cast<uint4 *>(polc + (pos0 + p_local) * 400 + cell * 16);      // Flatten in (y, x, c) order
                dst[0] = o[0]; dst[1] = o[1];
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        // ---- value head: dense_1 25 -> 32 ReLU, value_head 32 -> 1 tanh (model.py:95-103), fp32 ---------------------
        for (int p = warp; p < n_pos; p += 4) {
            const float *valc = reinterpret_cast<const float *>(smem + tcl::S_VALC) + p * 25;
            float acc = __ldg(fb + tcl::F_D1B + lane);
            for (int k = 0; k < 25; k++) acc = fmaf(valc[k], __ldg(fb + tcl::F_D1W + k * 32 + lane), acc);
            float s = fmaxf(acc, 0.f) * __ldg(fb + tcl::F_VHW + lane);
#pragma unroll
            for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
            if (lane == 0) value[pos0 + p] = tanhf(s + __ldg(fb + tcl::F_VHB));
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 128);
}

// =====================================================================================================
// Trunk kernel v3.  Same arithmetic and weight blobs as k_net_trunk_tc; what changes is the machinery
// around the 3x3 conv and the epilogues:
//  * a tile is FOUR positions whose 5x5 cells sit at row  p*30 + y*6 + x  of the 128-row operand (a guard row
//    after every board row, 8 idle rows at the end).  The 3x3 conv input is kept as three copies (one per
//    kernel row dy, pre-shifted vertically by the epilogue that produces it) in a ROW-CONTIGUOUS operand
//    layout, and the horizontal taps dx = -1, 0, +1 are the same copy addressed one row earlier / later
//    through the matrix descriptor (16-byte granular start address) — the guard rows supply the zero
//    padding.  Nine accumulating K = 32 GEMMs, ONE barrier phase, 6 shared-memory stores per thread instead
//    of an explicit im2col (36 stores per thread and three barrier phases in v2);
//  * 256 threads: warps w and w+4 share TMEM lane group w and split the output columns, so every epilogue
//    is half as long and 16 warps per SM (2 CTAs) hide each other's latencies;
//  * biases and the value-head dense live in shared memory (v2 re-read them from global inside every
//    epilogue, on the critical path), the next tile's input planes are prefetched into registers.
#ifdef CCX_TRUNK_TIMING
__device__ long long g_trunk_ts[2][1024];
#define TS(k) do { if (blockIdx.x == 0 && tile == blockIdx.x && (t == 0 || t == 255) && ts_n < 1024) g_trunk_ts[t ? 1 : 0][ts_n++] = clock64(); } while (0)
#else
#define TS(k) do { } while (0)
#endif
namespace tc3 {
constexpr int THREADS = 256, POS = 4, POS_ROWS = 30, LIVE_ROWS = 120;
constexpr int YROWS = 136, Y_LBO = YROWS * 16, Y_COPY = 4 * Y_LBO;        // 1 guard row + 128 + 1 guard row, padded to 8
constexpr int S_X = 0;                                   // [128 x 64] operand (block input / conv1 im2col); first half doubles as
                                                         // the [128 x 32] operand of conv C
constexpr int S_Y = S_X + 128 * 64 * 2;                  // three row-contiguous [136 x 32] copies of conv A's output
constexpr int S_WC1 = S_Y + 3 * Y_COPY;                  // conv1 weights, resident
constexpr int S_WA = S_WC1 + 8192, S_WB = S_WA + 4096, S_WC = S_WB + 18432;     // streamed per layer
constexpr int S_F = S_WC + 4096;                         // fp32 blob (biases, value-head dense)
constexpr int F_BYTES = ((tcl::F_TOTAL * 4 + 15) / 16) * 16;
constexpr int PLANES_BYTES = 1376;                       // 4 x 343 = 1372, padded
constexpr int S_PLANES = S_F + F_BYTES;                  // double buffer
constexpr int S_VALC = S_PLANES + 2 * PLANES_BYTES;
constexpr int S_TOTAL = S_VALC + 512;
static_assert(S_TOTAL <= 113 * 1024, "two CTAs per SM");
static_assert(S_Y % 128 == 0 && S_WC1 % 128 == 0 && S_F % 16 == 0 && S_PLANES % 16 == 0, "alignment");
}  // namespace tc3

template <bool FP16>
__global__ void __launch_bounds__(tc3::THREADS, 2)
k_net_trunk_tc3(const uint8_t *__restrict__ wb, const float *__restrict__ fb, const uint8_t *__restrict__ planes, int64_t n,
                __nv_bfloat16 *__restrict__ polc, float *__restrict__ value)
{
    using namespace tc3;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int rg = warp & 3, h = warp >> 2;                  // TMEM lane group, column half
    const int r = rg * 32 + lane;                            // operand row of this thread
    const int p_local = r / POS_ROWS, rem = r % POS_ROWS, cy = rem / 6, cx = rem % 6;
    const bool live = r < LIVE_ROWS && cx < 5;
    const int cell = cy * 5 + cx;
    const uint32_t sbase = umma::smem_u32(smem);
    const float *sF = reinterpret_cast<const float *>(smem + S_F);
    const int64_t n_tiles = (n + POS - 1) / POS;
    const bool planes_aligned = (reinterpret_cast<uintptr_t>(planes) & 3) == 0;

    auto refill3 = [&](int dst, const uint8_t *src, int bytes) {
        for (int i = t; i < bytes / 16; i += THREADS) cp_async16(sbase + dst + i * 16, src + i * 16);
        cp_async_commit();
    };
    // words q*256 + t (< 343) of a tile's input planes; bytes past the last position read as zero
    auto load_planes = [&](int64_t tile, uint32_t (&w)[2]) {
        const int64_t pos0 = tile * POS;
        const int bytes = (int)min((int64_t)POS, n - pos0) * 343;
        const uint8_t *src = planes + pos0 * 343;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = q * THREADS + t;
            uint32_t v = 0;
            if (idx * 4 + 4 <= bytes && planes_aligned) v = __ldg(reinterpret_cast<const uint32_t *>(src) + idx);
            else
                for (int k = 0; k < 4; k++) if (idx * 4 + k < bytes) v |= (uint32_t)__ldg(src + idx * 4 + k) << (8 * k);
            w[q] = v;
        }
    };
    auto store_planes = [&](int buf, const uint32_t (&w)[2]) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = q * THREADS + t;
            if (idx < 344) reinterpret_cast<uint32_t *>(smem + S_PLANES + buf * PLANES_BYTES)[idx] = w[q];
        }
    };

    for (int i = t; i < 8192 / 16; i += THREADS) cp_async16(sbase + S_WC1 + i * 16, wb + tcl::W_CONV1 + i * 16);
    refill3(S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096);                // group: conv1 + A0
    refill3(S_WB, wb + tcl::W_BLOCK0 + tcl::W_BB, 18432);               // group: B0
    refill3(S_WC, wb + tcl::W_BLOCK0 + tcl::W_BC, 4096);                // group: C0
    for (int i = t; i < S_WC1 / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);   // guard rows stay zero
    for (int i = t; i < tcl::F_TOTAL; i += THREADS) reinterpret_cast<float *>(smem + S_F)[i] = __ldg(fb + i);
    {
        uint32_t w0[2];
        load_planes(blockIdx.x, w0);
        store_planes(0, w0);
    }
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(rg * 32) << 16);
    uint32_t phase = 0;
    int buf = 0;
    // base descriptors of every operand (constant for the whole kernel) and the instruction descriptors
    const umma::DescBase dX64 = umma::desc_base(sbase + S_X, 128u, 64 / 8 * 128u), dX32 = umma::desc_base(sbase + S_X, 128u, 32 / 8 * 128u);
    const umma::DescBase dY = umma::desc_base(sbase + S_Y, Y_LBO, 128u);
    const umma::DescBase dWC1 = umma::desc_base(sbase + S_WC1, 128u, 64 / 8 * 128u), dWA = umma::desc_base(sbase + S_WA, 128u, 64 / 8 * 128u);
    const umma::DescBase dWB = umma::desc_base(sbase + S_WB, 128u, 288 / 8 * 128u), dWC = umma::desc_base(sbase + S_WC, 128u, 32 / 8 * 128u);
    constexpr uint32_t ID32 = umma::make_idesc(32, FP16), ID64 = umma::make_idesc(64, FP16);
#ifdef CCX_TRUNK_TIMING
    int ts_n = 0;
    int64_t tile = blockIdx.x;
#endif

    auto layer_sync = [&]() {
        TS(0);
        cp_async_wait<2>();
        TS(1);
        umma::fence_async_smem();
        TS(2);
        umma::fence_before_sync();
        __syncthreads();
        TS(3);
    };
    auto wait_mma = [&]() {
        TS(4);
        umma::mbar_wait(&bar, phase); phase ^= 1;
        umma::fence_after_sync();
        TS(5);
    };
    auto pack8 = [&](const float *rr) {
        return make_uint4(pack2<FP16>(rr[0], rr[1]), pack2<FP16>(rr[2], rr[3]), pack2<FP16>(rr[4], rr[5]), pack2<FP16>(rr[6], rr[7]));
    };

#ifdef CCX_TRUNK_TIMING
    for (tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#else
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#endif
        const int64_t pos0 = tile * POS;
        const int n_pos = (int)min((int64_t)POS, n - pos0);
        uint32_t pw[2] = {0u, 0u};
        if (tile + gridDim.x < n_tiles) load_planes(tile + gridDim.x, pw);       // lands while this tile computes
        cp_async_commit();                                   // (empty group: keeps the group count per phase uniform)
        // ---- conv1 operand: im2col of the 3x3 'valid' window, K = 63 (+1 zero) (model.py:62); this thread: 32 of the 64 columns
        {
            const uint8_t *pl = smem + S_PLANES + buf * PLANES_BYTES + (live ? p_local * 343 : 0);
#pragma unroll
            for (int c8 = 0; c8 < 4; c8++) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int kk = h * 32 + c8 * 8 + q;
                    const int tap = kk / 7, ch = kk % 7, dy = tap / 3, dx = tap % 3;
                    v[q] = (live && kk < 63) ? (float)pl[((cy + dy) * 7 + (cx + dx)) * 7 + ch] : 0.f;
                }
                *reinterpret_cast<uint4 *>(smem + S_X + umma::op_offset(r, h * 32 + c8 * 8, 64)) = pack8(v);
            }
        }
        layer_sync();
        if (warp == 0) {
            umma::fence_after_sync();
            if (umma::elect_one()) {
                umma::gemm_issue_d<64>(tmem + 64, dX64, 0, dWC1, 0, ID64, false);
                umma::commit(&bar);
            }
            __syncwarp();
        }
        cp_async_commit();                                   // (empty group)
        wait_mma();
        float x[32];                                         // residual stream: 32 of this row's 64 channels, fp32
        {
            float v[32];
            umma::tmem_ld32(trow + 64 + h * 32, v);
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = fmaxf(v[j] + sF[tcl::F_CONV1 + h * 32 + j], 0.f);
#pragma unroll
            for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(smem + S_X + umma::op_offset(r, h * 32 + c * 8, 64)) = pack8(x + c * 8);
        }
        // ---- 9 bottleneck residual blocks (model.py:120-145) --------------------------------------------------
        for (int b = 0; b < 9; b++) {
            const uint8_t *wnext = wb + tcl::W_BLOCK0 + ((b + 1) % 9) * tcl::W_BLOCK;
            const float *bias = sF + tcl::F_BLOCK0 + b * tcl::F_BLOCK;
            // A: 1x1 conv 64 -> 32, ReLU
            layer_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    umma::gemm_issue_d<64>(tmem, dX64, 0, dWA, 0, ID32, false);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            refill3(S_WA, b < 8 ? wnext + tcl::W_BA : wb + tcl::W_HEADS, 4096);
            {
                float v[16];
                umma::tmem_ld16(trow + h * 16, v);
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = fmaxf(v[q] + bias[h * 16 + q], 0.f);
                const uint4 o0 = pack8(v), o1 = pack8(v + 8);
                if (live) {
                    // copy d serves kernel row dy = d - 1: output cell (y - dy, x) reads this cell, so the value goes to row r - 6*dy
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const int oy = cy - (d - 1);
                        if (oy >= 0 && oy <= 4) {
                            uint8_t *dst = smem + S_Y + d * Y_COPY + (2 * h) * Y_LBO + (1 + r - 6 * (d - 1)) * 16;
                            *reinterpret_cast<uint4 *>(dst) = o0;
                            *reinterpret_cast<uint4 *>(dst + Y_LBO) = o1;
                        }
                    }
                }
            }
            // B: 3x3 'same' conv 32 -> 32 = nine accumulating K = 32 GEMMs on row-shifted views of the three copies
            layer_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
#pragma unroll
                    for (int d = 0; d < 3; d++)
#pragma unroll
                        for (int dxi = 0; dxi < 3; dxi++)
#pragma unroll
                            for (int ks = 0; ks < 2; ks++)
                                umma::mma_bf16(tmem + 32, umma::desc_at(dY, (uint32_t)(d * Y_COPY + dxi * 16 + 2 * ks * Y_LBO)),
                                               umma::desc_at(dWB, (uint32_t)(((d * 3 + dxi) * 32 / 8 + 2 * ks) * 128)), ID32, (d | dxi | ks) != 0);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            refill3(S_WB, wnext + tcl::W_BB, 18432);
            {
                float v[16];
                umma::tmem_ld16(trow + 32 + h * 16, v);
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = fmaxf(v[q] + bias[32 + h * 16 + q], 0.f);
                *reinterpret_cast<uint4 *>(smem + S_X + umma::op_offset(r, h * 16, 32)) = pack8(v);
                *reinterpret_cast<uint4 *>(smem + S_X + umma::op_offset(r, h * 16 + 8, 32)) = pack8(v + 8);
            }
            // C: 1x1 conv 32 -> 64, + skip, ReLU (model.py:137-144)
            layer_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    umma::gemm_issue_d<32>(tmem + 64, dX32, 0, dWC, 0, ID64, false);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            refill3(S_WC, wnext + tcl::W_BC, 4096);
            {
                float v[32];
                umma::tmem_ld32(trow + 64 + h * 32, v);
#pragma unroll
                for (int j = 0; j < 32; j++) x[j] = fmaxf(v[j] + bias[64 + h * 32 + j] + x[j], 0.f);
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(smem + S_X + umma::op_offset(r, h * 32 + c * 8, 64)) = pack8(x + c * 8);
            }
        }
        // ---- heads: policy conv 64 -> 16 and value conv 64 -> 1 in one N = 32 GEMM (model.py:91, 108); weights in the A slot
        layer_sync();
        if (warp == 0) {
            umma::fence_after_sync();
            if (umma::elect_one()) {
                umma::gemm_issue_d<64>(tmem, dX64, 0, dWA, 0, ID32, false);
                umma::commit(&bar);
            }
            __syncwarp();
        }
        wait_mma();
        refill3(S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096);                 // conv A of block 0 for the next tile
        {
            float v[16];
            umma::tmem_ld16(trow + h * 16, v);
            if (live && p_local < n_pos) {
                if (h == 0) {
#pragma unroll
                    for (int q = 0; q < 16; q++) v[q] = fmaxf(v[q] + sF[tcl::F_HEADS + q], 0.f);
                    uint4 *dst = reinterpret_cast<uint4 *>(polc + (pos0 + p_local) * 400 + cell * 16);      // Flatten in (y, x, c) order
                    dst[0] = pack8(v); dst[1] = pack8(v + 8);
                } else {
                    reinterpret_cast<float *>(smem + S_VALC)[p_local * 25 + cell] = fmaxf(v[0] + sF[tcl::F_HEADS + 16], 0.f);
                }
            }
        }
        store_planes(buf ^ 1, pw);
        umma::fence_before_sync();
        __syncthreads();
        // ---- value head: dense_1 25 -> 32 ReLU, value_head 32 -> 1 tanh (model.py:95-103), fp32 ---------------------
        if (warp < n_pos) {
            const float *valc = reinterpret_cast<const float *>(smem + S_VALC) + warp * 25;
            float acc = sF[tcl::F_D1B + lane];
            for (int k = 0; k < 25; k++) acc = fmaf(valc[k], sF[tcl::F_D1W + k * 32 + lane], acc);
            float sv = fmaxf(acc, 0.f) * sF[tcl::F_VHW + lane];
#pragma unroll
            for (int off = 16; off; off >>= 1) sv += __shfl_xor_sync(0xFFFFFFFFu, sv, off);
            if (lane == 0) value[pos0 + warp] = tanhf(sv + sF[tcl::F_VHB]);
        }
        buf ^= 1;
    }
    cp_async_wait<0>();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 128);
}

// logits[B x 294] = polc[B x 400] (bf16) * W (bf16) + b  —  128 positions x 160 outputs per CTA, K in two chunks
template <bool FP16>
__global__ void __launch_bounds__(128, 1)
k_policy_dense_tc(const uint8_t *__restrict__ wb, const float *__restrict__ fb, const __nv_bfloat16 *__restrict__ polc, int64_t n,
                  float *__restrict__ logits)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const int half = blockIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * 128;
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr int SA = 0, SB = 128 * 208 * 2;
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;
    for (int chunk = 0; chunk < 2; chunk++) {
        const int k0 = chunk ? 208 : 0, Kc = chunk ? 192 : 208;
        // A chunk: rows row0..row0+127, columns k0..k0+Kc of polc (16-byte pieces into the operand layout)
        const int pieces = Kc / 8;
        for (int i = t; i < 128 * pieces; i += 128) {
            const int r = i / pieces, k8 = i % pieces;
            const int64_t row = row0 + r;
            const uint32_t dst = sbase + SA + umma::op_offset(r, k8 * 8, Kc);
            if (row < n) cp_async16(dst, polc + row * 400 + k0 + k8 * 8);
            else *reinterpret_cast<uint4 *>(smem + SA + umma::op_offset(r, k8 * 8, Kc)) = make_uint4(0, 0, 0, 0);
        }
        const uint8_t *wsrc = wb + tcl::W_POLD + half * tcl::POLD_HALF + (chunk ? tcl::POLD_C0 : 0);
        const int wbytes = chunk ? tcl::POLD_C1 : tcl::POLD_C0;
        for (int i = t; i < wbytes / 16; i += 128) cp_async16(sbase + SB + i * 16, wsrc + i * 16);
        cp_async_commit();
        cp_async_wait<0>();
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            umma::gemm_issue(tmem, sbase + SA, Kc, 0, sbase + SB, Kc, 0, Kc, 160, chunk > 0, FP16);
            umma::commit(&bar);
        }
        umma::mbar_wait(&bar, phase); phase ^= 1;       // operands are free to be overwritten once the MMAs are done
        umma::fence_after_sync();
    }
    const int64_t row = row0 + t;
    for (int c = 0; c < 160; c += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        if (row < n) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int col = half * 160 + c + j;
                if (col < CCX_NUM_ACTIONS) logits[row * CCX_NUM_ACTIONS + col] = v[j] + __ldg(fb + tcl::F_POLD + col);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 256);
}

// =====================================================================================================
// Trunk kernel v4 = v3 with the 1x1 convs' A operands and the residual stream moved into TENSOR MEMORY:
//  * the fp32 residual x lives in TMEM columns [0,64): conv C accumulates straight onto it (D += m2 * Wc with
//    D = x), its epilogue applies bias + ReLU in place and writes the 16-bit copy of x that conv A of the next
//    block reads — also in TMEM (columns [64,96), two channels per column, tcgen05.mma "ts" form).  conv B's
//    output goes the same way to conv C (columns [64,80)).  Only the 3x3 conv still reads its A operand from
//    shared memory (the three row-shifted copies), so a block's tensor-core operand traffic out of shared
//    memory drops from 122 KB to 98 KB per tile and two of its three epilogues need no shared-memory store and
//    no generic->async proxy fence;
//  * no residual registers and no [128 x 64] shared operand: 70 KB of shared memory, <= 85 registers,
//    128 TMEM columns per CTA -> THREE CTAs (24 warps) per SM instead of two.
// TMEM columns: X fp32 [0,64) | XB: 16-bit x [64,96), reused as M2B: 16-bit conv-B output [64,80) |
//               AO: conv A / conv B / heads accumulator [96,128).
namespace tc4 {
constexpr int THREADS = 256, POS = 4, POS_ROWS = 30, LIVE_ROWS = 120;
constexpr int YROWS = 130, Y_LBO = YROWS * 16, Y_COPY = 4 * Y_LBO;        // 1 guard row + 128 + 1 guard row
constexpr int T_X = 0, T_XB = 64, T_AO = 96;
constexpr int S_Y = 0;
constexpr int S_WC1 = S_Y + 3 * Y_COPY;
constexpr int S_WA = S_WC1 + 8192, S_WB = S_WA + 4096, S_WC = S_WB + 18432;
constexpr int S_F = S_WC + 4096;                         // floats [0, F_POLD) then [F_D1W, F_TOTAL)
constexpr int NF_A = tcl::F_POLD, NF_B = tcl::F_TOTAL - tcl::F_D1W;
constexpr int F_BYTES = (((NF_A + NF_B) * 4 + 15) / 16) * 16;
constexpr int FO_D1W = NF_A, FO_D1B = FO_D1W + 800, FO_VHW = FO_D1B + 32, FO_VHB = FO_VHW + 32;
constexpr int S_PLANES = S_F + F_BYTES;
constexpr int S_VALC = S_PLANES + 1376;
constexpr int S_TOTAL = S_VALC + 512;
static_assert(S_TOTAL <= 74 * 1024, "three CTAs per SM");
static_assert(S_WC1 % 128 == 0 && S_F % 16 == 0 && S_PLANES % 16 == 0, "alignment");
}  // namespace tc4

template <bool FP16>
__global__ void __launch_bounds__(tc4::THREADS, 3)
k_net_trunk_tc4(const uint8_t *__restrict__ wb, const float *__restrict__ fb, const uint8_t *__restrict__ planes, int64_t n,
                __nv_bfloat16 *__restrict__ polc, float *__restrict__ value)
{
    using namespace tc4;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar, barW[4];                // MMAs of a phase done; weight slots A, B, C and conv1 landed
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int rg = warp & 3, h = warp >> 2;                  // TMEM lane group, column half
    const int r = rg * 32 + lane;                            // operand row of this thread
    const int p_local = r / POS_ROWS, rem = r % POS_ROWS, cy = rem / 6, cx = rem % 6;
    const bool live = r < LIVE_ROWS && cx < 5;
    const int cell = cy * 5 + cx;
    const uint32_t sbase = umma::smem_u32(smem);
    const float *sF = reinterpret_cast<const float *>(smem + S_F);
    const int64_t n_tiles = (n + POS - 1) / POS;
    const bool planes_aligned = (reinterpret_cast<uintptr_t>(planes) & 3) == 0;

    // weight slots are refilled by ONE lane with a bulk copy (TMA engine); only the MMA-issuing lane ever waits for them
    auto refill_slot = [&](int slot, int dst, const uint8_t *src, uint32_t bytes) {
        umma::mbar_expect_tx(&barW[slot], bytes);
        umma::bulk_g2s(sbase + dst, src, bytes, &barW[slot]);
    };
    // words t and 256 + t (< 343) of a tile's input planes; bytes past the last position read as zero
    auto load_planes = [&](int64_t tile, uint32_t (&w)[2]) {
        const int bytes = (int)min((int64_t)POS, n - tile * POS) * 343;
        const uint8_t *src = planes + tile * (POS * 343);
        w[0] = 0u; w[1] = 0u;
        if (bytes == POS * 343 && planes_aligned) {
            w[0] = __ldg(reinterpret_cast<const uint32_t *>(src) + t);
            if (t < 343 - THREADS) w[1] = __ldg(reinterpret_cast<const uint32_t *>(src) + THREADS + t);
        } else {
#pragma unroll
            for (int q = 0; q < 2; q++)
                for (int k = 0; k < 4; k++) {
                    const int byte = (q * THREADS + t) * 4 + k;
                    if (byte < bytes) w[q] |= (uint32_t)__ldg(src + byte) << (8 * k);
                }
        }
    };
    auto store_planes = [&](const uint32_t (&w)[2]) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = q * THREADS + t;
            if (idx < 344) reinterpret_cast<uint32_t *>(smem + S_PLANES)[idx] = w[q];
        }
    };

    for (int i = t; i < S_WC1 / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);   // guard rows stay zero
    for (int i = t; i < NF_A + NF_B; i += THREADS) reinterpret_cast<float *>(smem + S_F)[i] = __ldg(fb + (i < NF_A ? i : tcl::F_D1W + i - NF_A));
    {
        uint32_t w0[2];
        load_planes(blockIdx.x, w0);
        store_planes(w0);
    }
    if (t == 0) {
        umma::mbar_init(&bar, 1);
#pragma unroll
        for (int q = 0; q < 4; q++) umma::mbar_init(&barW[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        refill_slot(3, S_WC1, wb + tcl::W_CONV1, 8192);
        refill_slot(0, S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096);
        refill_slot(1, S_WB, wb + tcl::W_BLOCK0 + tcl::W_BB, 18432);
        refill_slot(2, S_WC, wb + tcl::W_BLOCK0 + tcl::W_BC, 4096);
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    uint32_t phW0 = 0, phW1 = 0, phW2 = 0;           // parities of the weight-slot barriers (tracked by the issuing lane)
    bool conv1_ready = false;
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(rg * 32) << 16);
    uint32_t phase = 0;
    const umma::DescBase dY = umma::desc_base(sbase + S_Y, Y_LBO, 128u);
    const umma::DescBase dWC1 = umma::desc_base(sbase + S_WC1, 128u, 64 / 8 * 128u), dWA = umma::desc_base(sbase + S_WA, 128u, 64 / 8 * 128u);
    const umma::DescBase dWB = umma::desc_base(sbase + S_WB, 128u, 288 / 8 * 128u), dWC = umma::desc_base(sbase + S_WC, 128u, 32 / 8 * 128u);
    constexpr uint32_t ID32 = umma::make_idesc(32, FP16), ID64 = umma::make_idesc(64, FP16);

    // operands written to TMEM: make the stores visible to the tensor core, then the CTA barrier
    auto tmem_sync = [&]() {
        umma::tmem_wait_st();
        umma::fence_before_sync();
        __syncthreads();
    };
    // operands written to shared memory (conv A's output copies)
    auto smem_sync = [&]() {
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
    };
    auto wait_mma = [&]() {
        umma::mbar_wait(&bar, phase); phase ^= 1;
        umma::fence_after_sync();
    };
    auto pack8 = [&](const float *rr) {
        return make_uint4(pack2<FP16>(rr[0], rr[1]), pack2<FP16>(rr[2], rr[3]), pack2<FP16>(rr[4], rr[5]), pack2<FP16>(rr[6], rr[7]));
    };
    // x (32 fp32 columns of this thread, already bias+ReLU'ed) -> TMEM X in place and its 16-bit copy XB, 16 columns at a time
    auto finish_x = [&](const float *bias32, bool add_bias_from_tmem_value_only) {
        (void)add_bias_from_tmem_value_only;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float v[16];
            umma::tmem_ld16(trow + T_X + h * 32 + half * 16, v);
            uint32_t f[16], pk[8];
#pragma unroll
            for (int j = 0; j < 16; j++) { v[j] = fmaxf(v[j] + bias32[half * 16 + j], 0.f); f[j] = __float_as_uint(v[j]); }
#pragma unroll
            for (int j = 0; j < 8; j++) pk[j] = pack2<FP16>(v[2 * j], v[2 * j + 1]);
            umma::tmem_st16(trow + T_X + h * 32 + half * 16, f);
            umma::tmem_st8(trow + T_XB + h * 16 + half * 8, pk);
        }
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pos0 = tile * POS;
        const int n_pos = (int)min((int64_t)POS, n - pos0);
        uint32_t pw[2] = {0u, 0u};
        if (tile + gridDim.x < n_tiles) load_planes(tile + gridDim.x, pw);       // lands while this tile computes
        // ---- conv1 operand: im2col of the 3x3 'valid' window, K = 63 (+1 zero) (model.py:62), packed into TMEM XB
        {
            const uint8_t *pl = smem + S_PLANES + (live ? p_local * 343 : 0);
#pragma unroll
            for (int c8 = 0; c8 < 2; c8++) {
                uint32_t pk[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float v2[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int kk = h * 32 + c8 * 16 + q * 2 + e;
                        const int tap = kk / 7, ch = kk % 7, dy = tap / 3, dx = tap % 3;
                        v2[e] = (live && kk < 63) ? (float)pl[((cy + dy) * 7 + (cx + dx)) * 7 + ch] : 0.f;
                    }
                    pk[q] = pack2<FP16>(v2[0], v2[1]);
                }
                umma::tmem_st8(trow + T_XB + h * 16 + c8 * 8, pk);
            }
        }
        tmem_sync();
        if (warp == 0) {
            umma::fence_after_sync();
            if (umma::elect_one()) {
                if (!conv1_ready) { umma::mbar_wait(&barW[3], 0); conv1_ready = true; }
                umma::gemm_issue_ts<64>(tmem + T_X, tmem + T_XB, dWC1, 0, ID64, false);
                umma::commit(&bar);
            }
            __syncwarp();
        }
        wait_mma();
        finish_x(sF + tcl::F_CONV1 + h * 32, false);
        // ---- 9 bottleneck residual blocks (model.py:120-145) --------------------------------------------------
        for (int b = 0; b < 9; b++) {
            const uint8_t *wnext = wb + tcl::W_BLOCK0 + ((b + 1) % 9) * tcl::W_BLOCK;
            const float *bias = sF + tcl::F_BLOCK0 + b * tcl::F_BLOCK;
            // A: 1x1 conv 64 -> 32, ReLU; A operand = XB in TMEM
            tmem_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    umma::mbar_wait(&barW[0], phW0); phW0 ^= 1;
                    umma::gemm_issue_ts<64>(tmem + T_AO, tmem + T_XB, dWA, 0, ID32, false);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            if (warp == 0 && umma::elect_one()) refill_slot(0, S_WA, b < 8 ? wnext + tcl::W_BA : wb + tcl::W_HEADS, 4096);
            {
                float v[16];
                umma::tmem_ld16(trow + T_AO + h * 16, v);
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = fmaxf(v[q] + bias[h * 16 + q], 0.f);
                const uint4 o0 = pack8(v), o1 = pack8(v + 8);
                if (live) {
                    // copy d serves kernel row dy = d - 1: output cell (y - dy, x) reads this cell, so the value goes to row r - 6*dy
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const int oy = cy - (d - 1);
                        if (oy >= 0 && oy <= 4) {
                            uint8_t *dst = smem + S_Y + d * Y_COPY + (2 * h) * Y_LBO + (1 + r - 6 * (d - 1)) * 16;
                            *reinterpret_cast<uint4 *>(dst) = o0;
                            *reinterpret_cast<uint4 *>(dst + Y_LBO) = o1;
                        }
                    }
                }
            }
            // B: 3x3 'same' conv 32 -> 32 = nine accumulating K = 32 GEMMs on row-shifted views of the three copies
            smem_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    umma::mbar_wait(&barW[1], phW1); phW1 ^= 1;
#pragma unroll
                    for (int d = 0; d < 3; d++)
#pragma unroll
                        for (int dxi = 0; dxi < 3; dxi++)
#pragma unroll
                            for (int ks = 0; ks < 2; ks++)
                                umma::mma_bf16(tmem + T_AO, umma::desc_at(dY, (uint32_t)(d * Y_COPY + dxi * 16 + 2 * ks * Y_LBO)),
                                               umma::desc_at(dWB, (uint32_t)(((d * 3 + dxi) * 32 / 8 + 2 * ks) * 128)), ID32, (d | dxi | ks) != 0);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            if (warp == 0 && umma::elect_one()) refill_slot(1, S_WB, wnext + tcl::W_BB, 18432);
            {
                float v[16];
                umma::tmem_ld16(trow + T_AO + h * 16, v);
                uint32_t pk[8];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    pk[q] = pack2<FP16>(fmaxf(v[2 * q] + bias[32 + h * 16 + 2 * q], 0.f), fmaxf(v[2 * q + 1] + bias[32 + h * 16 + 2 * q + 1], 0.f));
                umma::tmem_st8(trow + T_XB + h * 8, pk);         // M2B: 32 channels = 16 columns
            }
            // C: 1x1 conv 32 -> 64 accumulated onto the residual (model.py:137-144): X += M2B * Wc, then bias + ReLU in place
            tmem_sync();
            if (warp == 0) {
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    umma::mbar_wait(&barW[2], phW2); phW2 ^= 1;
                    umma::gemm_issue_ts<32>(tmem + T_X, tmem + T_XB, dWC, 0, ID64, true);
                    umma::commit(&bar);
                }
                __syncwarp();
            }
            wait_mma();
            if (warp == 0 && umma::elect_one()) refill_slot(2, S_WC, wnext + tcl::W_BC, 4096);
            finish_x(bias + 64 + h * 32, true);
        }
        // ---- heads: policy conv 64 -> 16 and value conv 64 -> 1 in one N = 32 GEMM (model.py:91, 108); weights in the A slot
        tmem_sync();
        if (warp == 0) {
            umma::fence_after_sync();
            if (umma::elect_one()) {
                umma::mbar_wait(&barW[0], phW0); phW0 ^= 1;
                umma::gemm_issue_ts<64>(tmem + T_AO, tmem + T_XB, dWA, 0, ID32, false);
                umma::commit(&bar);
            }
            __syncwarp();
        }
        wait_mma();
        if (warp == 0 && umma::elect_one()) refill_slot(0, S_WA, wb + tcl::W_BLOCK0 + tcl::W_BA, 4096);   // conv A of block 0 for the next tile
        {
            float v[16];
            umma::tmem_ld16(trow + T_AO + h * 16, v);
            if (live && p_local < n_pos) {
                if (h == 0) {
#pragma unroll
                    for (int q = 0; q < 16; q++) v[q] = fmaxf(v[q] + sF[tcl::F_HEADS + q], 0.f);
                    uint4 *dst = reinterpret_cast<uint4 *>(polc + (pos0 + p_local) * 400 + cell * 16);      // Flatten in (y, x, c) order
                    dst[0] = pack8(v); dst[1] = pack8(v + 8);
                } else {
                    reinterpret_cast<float *>(smem + S_VALC)[p_local * 25 + cell] = fmaxf(v[0] + sF[tcl::F_HEADS + 16], 0.f);
                }
            }
        }
        store_planes(pw);
        umma::fence_before_sync();
        __syncthreads();
        // ---- value head: dense_1 25 -> 32 ReLU, value_head 32 -> 1 tanh (model.py:95-103), fp32 ---------------------
        if (warp < n_pos) {
            const float *valc = reinterpret_cast<const float *>(smem + S_VALC) + warp * 25;
            float acc = sF[FO_D1B + lane];
            for (int k = 0; k < 25; k++) acc = fmaf(valc[k], sF[FO_D1W + k * 32 + lane], acc);
            float sv = fmaxf(acc, 0.f) * sF[FO_VHW + lane];
#pragma unroll
            for (int off = 16; off; off >>= 1) sv += __shfl_xor_sync(0xFFFFFFFFu, sv, off);
            if (lane == 0) value[pos0 + warp] = tanhf(sv + sF[FO_VHB]);
        }
    }
    // the last refills (block 0's weights for a tile that never comes) must land before the CTA's shared memory is released
    if (warp == 0 && umma::elect_one()) {
        umma::mbar_wait(&barW[0], phW0); umma::mbar_wait(&barW[1], phW1); umma::mbar_wait(&barW[2], phW2);
        if (!conv1_ready) umma::mbar_wait(&barW[3], 0);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 128);
}

// Policy dense v2: 128 positions x 80 outputs per CTA (grid = row tiles x 4 column quarters: 128 CTAs at the
// self-play batch of 4,096 instead of 64), both K chunks in flight at once (two cp.async groups, the first
// chunk's MMAs run while the second lands), logits staged through shared memory for coalesced stores.
namespace pd2 {
constexpr int NQ = 80;
constexpr int S_A0 = 0, S_A1 = S_A0 + 128 * 208 * 2, S_W0 = S_A1 + 128 * 192 * 2, S_W1 = S_W0 + NQ * 208 * 2;
constexpr int S_TOTAL = S_W1 + NQ * 192 * 2;           // 166,400 B
constexpr int OUT_LD = NQ + 1;                          // fp32 staging [128][81] over the A region once the MMAs are done
static_assert(128 * OUT_LD * 4 <= S_W0, "staging fits in the A region");
}  // namespace pd2

template <bool FP16>
__global__ void __launch_bounds__(128, 1)
k_policy_dense_tc2(const uint8_t *__restrict__ wb, const float *__restrict__ fb, const __nv_bfloat16 *__restrict__ polc, int64_t n,
                   float *__restrict__ logits)
{
    using namespace pd2;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const int quarter = blockIdx.y, half = quarter >> 1, c0 = (quarter & 1) * NQ;
    const int64_t row0 = (int64_t)blockIdx.x * 128;
    const uint32_t sbase = umma::smem_u32(smem);
    if (t == 0) umma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
#pragma unroll
    for (int chunk = 0; chunk < 2; chunk++) {
        const int k0 = chunk ? 208 : 0, Kc = chunk ? 192 : 208, pieces = Kc / 8;
        const int sa = chunk ? S_A1 : S_A0, sw = chunk ? S_W1 : S_W0;
        for (int i = t; i < 128 * pieces; i += 128) {
            const int r = i / pieces, k8 = i % pieces;
            const int64_t row = row0 + r;
            if (row < n) cp_async16(sbase + sa + umma::op_offset(r, k8 * 8, Kc), polc + row * 400 + k0 + k8 * 8);
            else *reinterpret_cast<uint4 *>(smem + sa + umma::op_offset(r, k8 * 8, Kc)) = make_uint4(0, 0, 0, 0);
        }
        // this quarter's 80 weight rows are a contiguous slice of the half's [160 x Kc] operand
        const uint8_t *wsrc = wb + tcl::W_POLD + half * tcl::POLD_HALF + (chunk ? tcl::POLD_C0 : 0) + (c0 / 8) * pieces * 128;
        for (int i = t; i < NQ * Kc * 2 / 16; i += 128) cp_async16(sbase + sw + i * 16, wsrc + i * 16);
        cp_async_commit();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t ID = umma::make_idesc(NQ, FP16);
    cp_async_wait<1>();
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        if (umma::elect_one())
            umma::gemm_issue_d<208>(tmem, umma::desc_base(sbase + S_A0, 128u, 208 / 8 * 128u), 0, umma::desc_base(sbase + S_W0, 128u, 208 / 8 * 128u), 0, ID, false);
        __syncwarp();
    }
    cp_async_wait<0>();
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
            umma::gemm_issue_d<192>(tmem, umma::desc_base(sbase + S_A1, 128u, 192 / 8 * 128u), 0, umma::desc_base(sbase + S_W1, 128u, 192 / 8 * 128u), 0, ID, true);
            umma::commit(&bar);
        }
        __syncwarp();
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    float *stage = reinterpret_cast<float *>(smem);
#pragma unroll
    for (int c = 0; c < NQ; c += 16) {
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 16; j++) stage[t * OUT_LD + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    const int col_base = half * 160 + c0;
    for (int i = t; i < 128 * NQ; i += 128) {
        const int r = i / NQ, c = i % NQ, col = col_base + c;
        if (row0 + r < n && col < CCX_NUM_ACTIONS) logits[(row0 + r) * CCX_NUM_ACTIONS + col] = stage[r * OUT_LD + c] + __ldg(fb + tcl::F_POLD + col);
    }
    if (warp == 0) umma::tmem_free(tmem, 128);
}

static ccx_net_tc *tc_of(ccx_handle *h, bool create)
{
    if (!h->net_tc && create) h->net_tc = new (std::nothrow) ccx_net_tc();
    return h->net_tc;
}

void ccx_net_tc_free(ccx_handle *h)
{
    ccx_net_tc *tc = h->net_tc;
    if (!tc) return;
    cudaFree(tc->wb); cudaFree(tc->fb); cudaFree(tc->polc);
    delete tc;
    h->net_tc = nullptr;
}

extern "C" {

#ifdef CCX_TRUNK_TIMING
int ccx_debug_trunk_timing(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_trunk_ts, sizeof(long long) * 2048); }
#endif
int ccx_net_tc_blob_bytes(void) { return tcl::W_TOTAL; }
int ccx_net_tc_num_floats(void) { return tcl::F_TOTAL; }

int ccx_net_load_tc(ccx_handle *h, const void *bf16_blob_host, int64_t blob_bytes, const float *f32_host, int64_t n_floats,
                    int32_t fp16)
{
    if (!h || !bf16_blob_host || !f32_host || blob_bytes != tcl::W_TOTAL || n_floats != tcl::F_TOTAL || (fp16 != 0 && fp16 != 1))
        return CCX_ERR_ARG;
    ccx_net_tc *tc = tc_of(h, true);
    if (!tc) return CCX_ERR_NOMEM;
    if (!tc->wb) CCX_CUDA(h, cudaMalloc(&tc->wb, tcl::W_TOTAL));
    if (!tc->fb) CCX_CUDA(h, cudaMalloc(&tc->fb, sizeof(float) * tcl::F_TOTAL));
    CCX_CUDA(h, cudaMemcpyAsync(tc->wb, bf16_blob_host, tcl::W_TOTAL, cudaMemcpyHostToDevice, h->stream));
    CCX_CUDA(h, cudaMemcpyAsync(tc->fb, f32_host, sizeof(float) * tcl::F_TOTAL, cudaMemcpyHostToDevice, h->stream));
    CCX_CUDA(h, cudaStreamSynchronize(h->stream));
    tc->fp16 = fp16;
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcl::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcl::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_policy_dense_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd2::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_policy_dense_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd2::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc4::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc4::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_net_trunk_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::S_TOTAL));
    CCX_CUDA(h, cudaFuncSetAttribute(k_policy_dense_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 208 * 2 + 160 * 208 * 2));
    CCX_CUDA(h, cudaFuncSetAttribute(k_policy_dense_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 208 * 2 + 160 * 208 * 2));
    return CCX_OK;
}

// planes: uint8 (n,7,7,7) on the device -> logits float32[n][294], value float32[n]
int ccx_net_forward_tc(ccx_handle *h, int64_t n, const uint8_t *planes, float *logits, float *value)
{
    if (!h || n < 0 || (n && (!planes || !logits || !value))) return CCX_ERR_ARG;
    ccx_net_tc *tc = tc_of(h, false);
    if (!tc || !tc->wb) return CCX_ERR_STATE;
    if (n == 0) return CCX_OK;
    if (tc->cap < n) {
        if (tc->polc) CCX_CUDA(h, cudaFree(tc->polc));
        tc->polc = nullptr; tc->cap = 0;
        CCX_CUDA(h, cudaMalloc(&tc->polc, sizeof(__nv_bfloat16) * 400 * (size_t)n));
        tc->cap = n;
    }
    static const bool use_v2 = getenv("CCX_TRUNK_V2") != nullptr;        // A/B switches for profiling the older kernels
    static const bool use_v3 = getenv("CCX_TRUNK_V3") != nullptr;
    if (use_v2) {
        int64_t tiles = (n + 4) / 5;
        unsigned grid = (unsigned)(tiles < 2 * h->num_sms ? tiles : 2 * h->num_sms);     // two resident CTAs per SM
        if (tc->fp16) k_net_trunk_tc<true><<<grid, 128, tcl::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
        else k_net_trunk_tc<false><<<grid, 128, tcl::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
    } else if (use_v3) {
        int64_t tiles = (n + tc3::POS - 1) / tc3::POS;
        unsigned grid = (unsigned)(tiles < 2 * h->num_sms ? tiles : 2 * h->num_sms);     // two resident CTAs per SM
        if (tc->fp16) k_net_trunk_tc3<true><<<grid, tc3::THREADS, tc3::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
        else k_net_trunk_tc3<false><<<grid, tc3::THREADS, tc3::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
    } else {
        int64_t tiles = (n + tc4::POS - 1) / tc4::POS;
        unsigned grid = (unsigned)(tiles < 3 * h->num_sms ? tiles : 3 * h->num_sms);     // three resident CTAs per SM
        if (tc->fp16) k_net_trunk_tc4<true><<<grid, tc4::THREADS, tc4::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
        else k_net_trunk_tc4<false><<<grid, tc4::THREADS, tc4::S_TOTAL, h->stream>>>(tc->wb, tc->fb, planes, n, tc->polc, value);
    }
    CCX_LAUNCHED(h);
    if (use_v2) {
        dim3 g2((unsigned)((n + 127) / 128), 2);
        constexpr int SM2 = 128 * 208 * 2 + 160 * 208 * 2;
        if (tc->fp16) k_policy_dense_tc<true><<<g2, 128, SM2, h->stream>>>(tc->wb, tc->fb, tc->polc, n, logits);
        else k_policy_dense_tc<false><<<g2, 128, SM2, h->stream>>>(tc->wb, tc->fb, tc->polc, n, logits);
    } else {
        dim3 g2((unsigned)((n + 127) / 128), 4);
        if (tc->fp16) k_policy_dense_tc2<true><<<g2, 128, pd2::S_TOTAL, h->stream>>>(tc->wb, tc->fb, tc->polc, n, logits);
        else k_policy_dense_tc2<false><<<g2, 128, pd2::S_TOTAL, h->stream>>>(tc->wb, tc->fb, tc->polc, n, logits);
    }
    CCX_LAUNCHED(h);
    return CCX_OK;
}

}  // extern "C"
