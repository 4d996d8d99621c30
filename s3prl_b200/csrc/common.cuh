// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// the bf16 hi+lo split used for error-compensated tensor-core products, exact-erf GELU.
//
// Numerics contract: every fp32 contraction on the hot path is computed as
//   x*w ~= hi(x)*hi(w) + hi(x)*lo(w) + lo(x)*hi(w)     (bf16 operands, fp32 TMEM accumulation)
// with hi(v) = bf16_rn(v), lo(v) = bf16_rn(v - hi(v)); relative error per product ~2^-16.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace s3b {

// ----------------------------------------------------------------------------------------------
// small utilities
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float gelu_erf(float x) {
    // exact-erf GELU, nn.GELU() default (reference: wav2vec2_model.py:2893,2902,2904; 1899-1900):
    //   gelu(x) = 0.5 x (1 + erf(x/sqrt2)) = 0.5 x * (x >= 0 ? 2 - erfc(|x|/sqrt2) : erfc(|x|/sqrt2))
    // erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1/(1 + p z)   (Abramowitz-Stegun 7.1.26,
    // |error| <= 1.5e-7). ~17 instructions with two MUFU ops instead of erff's ~27, and no cancellation for
    // x < 0: measured max |error| vs float64 4.2e-7 over [-12, 12] (torch's own fp32 GELU: 1.2e-6), see
    // tests/test_oracle_cpu.py::test_gelu_formula. conv0 and the fc1 epilogue are issue-bound on this function.
#ifdef S3B_GELU_LIBM
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
#else
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * (z * -1.4426950408889634f)));
    const float erfc_z = poly * t * e;
    return 0.5f * x * (x >= 0.0f ? 2.0f - erfc_z : erfc_z);
#endif
}

// hi/lo split of an fp32 value into two bf16 values (round-to-nearest-even both times)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Blackwell packed fp32 arithmetic (FADD2) and three-input max (FMNMX3): the softmax and split epilogues are
// instruction-issue bound, these halve their add / max counts.
__device__ __forceinline__ void fsub2(float& x, float& y, float a0, float a1, float b0, float b1) {
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rd, ra, rb; "
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(x), "=f"(y)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2(float& x, float& y, float a0, float a1, float b0, float b1) {
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; "
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(x), "=f"(y)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// pack two floats -> two bf16x2 words (hi word, lo word); element 0 in the low half.
// 5 instructions per pair: F2FP (cvt.rn.bf16x2), SHF, LOP3, FADD2 (both residuals), F2FP
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    float ra, rb;
    fsub2(ra, rb, a, b, __uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// ----------------------------------------------------------------------------------------------
// "f16q8" operand format: x ~= h16 + l8 * 2^-kQ8ShiftA with h16 = fp16_rn(x) and two e4m3 planes
//   h8 = e4m3(h16)                        (the operand of the  A_hi * W_lo  correction MMA)
//   l8 = e4m3((x - h16) * 2^kQ8ShiftA)    (the operand of the  A_lo * W_hi  correction MMA)
// Weights: w16 = fp16_rn(w), wh8 = e4m3(w * 2^kQ8ShiftB), wl8 = e4m3((w - w16) * 2^kQ8ShiftD). Both correction
// products therefore carry the factor 2^kQ8Scale (= ShiftA + ShiftB = ShiftD); they are accumulated FIRST and the first
// main fp16 MMA of a tile scales the accumulator back with tcgen05.mma's scale-input-d immediate (D = A*B + D * 2^-15).
// 4 bytes per activation element, like the bf16 hi/lo pair it replaces; 8 instructions per element pair.
// ----------------------------------------------------------------------------------------------
static constexpr int kQ8ShiftA = 11;  // activations: residual of an 11-bit significand, scaled into e4m3's range
static constexpr int kQ8ShiftB = 4;   // weights |w| <= 28 stay finite in e4m3 (satfinite clamps beyond)
static constexpr int kQ8ShiftD = 15;  // weight residuals (|w - w16| <= 2^-11 |w|)
static constexpr int kQ8Scale = 15;   // = kQ8ShiftA + kQ8ShiftB = kQ8ShiftD (activation h8 plane is unscaled)
static_assert(kQ8ShiftA + kQ8ShiftB == kQ8Scale && kQ8ShiftD == kQ8Scale, "correction products must share one scale");

// two floats -> fp16x2 word (element 0 in the low half), e4m3x2 of the fp16 values, e4m3x2 of the scaled residuals
__device__ __forceinline__ void split_q8_pack2(float a, float b, uint32_t& h16, uint16_t& h8, uint16_t& l8) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h16) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(h8) : "r"(h16));
    float ha, hb;
    asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi; }"
        : "=f"(ha), "=f"(hb)
        : "r"(h16));
    float ra, rb;
    fsub2(ra, rb, a, b, ha, hb);
    const float sc = (float)(1 << kQ8ShiftA);
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(l8) : "f"(rb * sc), "f"(ra * sc));
}
// four consecutive elements -> one 8-byte fp16 store + two 4-byte e4m3 stores
__device__ __forceinline__ void store_q8x4(const float4& y, __nv_bfloat16* p16, uint8_t* ph8, uint8_t* pl8, size_t elem) {
    uint32_t h0, h1;
    uint16_t a0, a1, b0, b1;
    split_q8_pack2(y.x, y.y, h0, a0, b0);
    split_q8_pack2(y.z, y.w, h1, a1, b1);
    *reinterpret_cast<uint2*>(p16 + elem) = make_uint2(h0, h1);
    *reinterpret_cast<uint32_t*>(ph8 + elem) = (uint32_t)a0 | ((uint32_t)a1 << 16);
    *reinterpret_cast<uint32_t*>(pl8 + elem) = (uint32_t)b0 | ((uint32_t)b1 << 16);
}

// GEMM-operand output of a producer kernel (LayerNorm, conv-0, GEMM epilogues, attention, pos_conv combine):
// fmt 0 = bf16 hi / lo planes (bf16x3 scheme), fmt 1 = fp16 plane + two e4m3 planes (f16q8 scheme).
struct OutPlanes {
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    uint8_t* h8;
    uint8_t* l8;
    int fmt;
};
__host__ __device__ inline OutPlanes no_planes() { return OutPlanes{nullptr, nullptr, nullptr, nullptr, 0}; }

// four consecutive elements starting at element index `elem` (elem % 4 == 0)
__device__ __forceinline__ void store_planes4(const OutPlanes& o, const float4& y, size_t elem) {
    if (o.fmt != 0) {
        store_q8x4(y, o.hi, o.h8, o.l8, elem);
    } else {
        uint32_t h0, l0, h1, l1;
        split_pack2(y.x, y.y, h0, l0);
        split_pack2(y.z, y.w, h1, l1);
        *reinterpret_cast<uint2*>(o.hi + elem) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(o.lo + elem) = make_uint2(l0, l1);
    }
}
// eight consecutive elements (elem % 8 == 0): 16-byte stores on the 2-byte planes
__device__ __forceinline__ void store_planes8(const OutPlanes& o, const float (&y)[8], size_t elem) {
    if (o.fmt != 0) {
        uint32_t h[4];
        uint16_t a[4], b[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_q8_pack2(y[2 * e], y[2 * e + 1], h[e], a[e], b[e]);
        *reinterpret_cast<uint4*>(o.hi + elem) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint2*>(o.h8 + elem) =
            make_uint2((uint32_t)a[0] | ((uint32_t)a[1] << 16), (uint32_t)a[2] | ((uint32_t)a[3] << 16));
        *reinterpret_cast<uint2*>(o.l8 + elem) =
            make_uint2((uint32_t)b[0] | ((uint32_t)b[1] << 16), (uint32_t)b[2] | ((uint32_t)b[3] << 16));
    } else {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pack2(y[2 * e], y[2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4*>(o.hi + elem) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(o.lo + elem) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
// two consecutive elements (elem % 2 == 0)
__device__ __forceinline__ void store_planes2(const OutPlanes& o, float y0, float y1, size_t elem) {
    if (o.fmt != 0) {
        uint32_t h;
        uint16_t a, b;
        split_q8_pack2(y0, y1, h, a, b);
        *reinterpret_cast<uint32_t*>(o.hi + elem) = h;
        *reinterpret_cast<uint16_t*>(o.h8 + elem) = a;
        *reinterpret_cast<uint16_t*>(o.l8 + elem) = b;
    } else {
        uint32_t h, l;
        split_pack2(y0, y1, h, l);
        *reinterpret_cast<uint32_t*>(o.hi + elem) = h;
        *reinterpret_cast<uint32_t*>(o.lo + elem) = l;
    }
}
// read back element pair `elem` (even) of an operand in either format
__device__ __forceinline__ float2 load_planes2(const OutPlanes& o, size_t elem) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(o.hi + elem);
    if (o.fmt != 0) {
        const uint16_t l = *reinterpret_cast<const uint16_t*>(o.l8 + elem);
        float h0, h1;
        asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi; }"
            : "=f"(h0), "=f"(h1)
            : "r"(h));
        uint32_t lf;  // e4m3x2 -> f16x2
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(lf) : "h"(l));
        float l0, l1;
        asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi; }"
            : "=f"(l0), "=f"(l1)
            : "r"(lf));
        const float inv = 1.0f / (float)(1 << kQ8ShiftA);
        return make_float2(fmaf(l0, inv, h0), fmaf(l1, inv, h1));
    }
    const uint32_t l = *reinterpret_cast<const uint32_t*>(o.lo + elem);
    return make_float2(__uint_as_float(h << 16) + __uint_as_float(l << 16),
                       __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u));
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One lane of a CONVERGED warp. Unlike `lane == 0`, ptxas knows that the region guarded by elect.sync is
// single-threaded and feeds tcgen05 / TMA operands through uniform registers directly (no per-instruction
// ELECT/R2UR waterfall loop: ~50 issue cycles saved per tcgen05.mma).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
            printf("s3b: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 128B-swizzled tiles
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 (5th-gen tensor cores, TMEM accumulators)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 inputs, fp32 accumulate; single-CTA group.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane (lane_base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// A operand from TENSOR MEMORY (".ts" form): A[m][k] lives at lane m, 32-bit column k/2 (two consecutive K elements
// per column, even k in the low half) starting at tmem_a; B is a shared-memory descriptor as usual.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: thread i of the warp writes 16 consecutive 32-bit columns of lane (lane_base + i)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two SMs of one TPC issue ONE tcgen05.mma with M = 256; each CTA keeps its own 128
// rows of A and HALF of the B rows in its shared memory, which halves the per-SM operand traffic.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load executed by either CTA of the pair into its OWN smem; the transaction bytes are credited to the
// LEADER CTA's mbarrier (bit 24 of a shared::cluster address is the CTA's rank inside the pair).
__device__ __forceinline__ void tma_load_3d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f8f6f4 (e4m3 x e4m3, K = 32 per instruction, fp32 accumulate), CTA pair
__device__ __forceinline__ void umma_q8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with the scale-input-d immediate: D = A*B + D * 2^-kQ8Scale
__device__ __forceinline__ void umma_f16_2cta_scaled(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                     uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, 15;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc)
        : "memory");
}
static_assert(kQ8Scale == 15, "the scale-input-d immediate above is spelled out as 15");
// commit of the pair's MMAs, arriving on the same barrier offset in BOTH CTAs (mask 0b11)
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle (the layout TMA SWIZZLE_128B produces for a {64 x rows} bf16 box):
//   8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major (set to 1), version 1 (sm_100).
// Generic form: ROW_BYTES = 128 (SWIZZLE_128B, 8-row atoms 1024 B apart) or 64 (SWIZZLE_64B, atoms 512 B apart).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "unsupported swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
}
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (ignored), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
    return d;
}
// Instruction descriptor with A = B = fp16 (kind::f16) — the same bits select e4m3 x e4m3 under kind::f8f6f4.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Instruction descriptor: kind::f16, A=B=bf16 (K-major), D=fp32, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format fp32
           | (1u << 7)          // A format bf16
           | (1u << 10)         // B format bf16
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}

}  // namespace s3b

// Programmatic dependent launch (PDL): a kernel launched with the stream-serialization attribute may start its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel of the stream drains;
// `pdl_wait()` blocks until that previous grid has completed and its writes are visible, so every global read or
// write of a kernel must come after it. `pdl_launch_dependents()` lets the next kernel's CTAs be scheduled as
// soon as every CTA of this grid has issued it. Both are no-ops for a normal launch. S3B_PDL=0 disables.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S3B_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncSetAttribute applies to the current device only: one "done" flag per device so that a process that
// drives several GPUs configures each of them.
struct PerDeviceOnce {
    bool done[64] = {};
    bool& current() {
        int d = 0;
        cudaGetDevice(&d);
        return done[d & 63];
    }
};
