// TMA-fed tcgen05 (kind::tf32) GEMM with fused epilogues for the NAR dense layers.
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Replaces every tf.layers.Dense of the reference graph (nar_model.py:375-473) and the UGRNN
// input projection (:1317); forward, dgrad and wgrad all go through this one kernel by
// choosing operand majors (no transposed copies of activations are ever made):
//   fwd   Y  = X  * W        A = X  (K-major)   B = W   (MN-major: W is [in,out]) or W^T (K-major)
//   dgrad dX = dY * W^T      A = dY (K-major)   B = W   (K-major)
//   wgrad dW = X^T * dY      A = X  (MN-major)  B = dY  (MN-major), split-K + red.add
//
// CTA = 256 threads; T x T accumulator tiles of 128x128 fp32 in TMEM (T = 1: 128 columns, T = 2: all 512):
//   warp 0    TMA producer  (cp.async.bulk.tensor.2d, 128-byte swizzle, 32 fp32 = 128 B per row)
//   warp 1    MMA issuer    (one thread, tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128*T K=8)
//   warp 2    TMEM allocator
//   warps 4-7 3xTF32 operand split during the main loop, then the epilogue (tcgen05.ld 32x32b ->
//             bias / activation / activation-derivative -> st.global.v4 or red.global.add).
//
// Measured on B200 (ncu, profiles/): with 128x128 fp32 tiles the single-pass kernel asks L2 for 64 B per kMAC
// and sits at the chip's L2->SM throughput (tensor pipe 20 % busy).  T = 2 (256x256 per CTA, 64 KB per k-tile
// for 4x the MACs) halves that; it is used for the big single-pass (backward) GEMMs.
//
// 3xTF32: the tensor core reads fp32 bits as tf32 by dropping the low 13 mantissa bits, so the
// "hi" operand is the TMA tile itself; D += Alo*Bhi + Ahi*Blo + Ahi*Bhi restores ~fp32 accuracy
// (the reference is fp32 end to end and logits are divided by temperature 0.1 before exp).
//   MODE 0: single pass.   MODE 1: 3x, both lo tiles produced in-kernel.
//   MODE 2: 3x, B_lo (weights) read from HBM by TMA (nar_adam_tf maintains it), only A is split in-kernel.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

namespace nar {
namespace gemm {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 32;            // 32 fp32 = 128 B = one 128-byte swizzle span
constexpr int UMMA_K = 8;         // tf32: 32 B of K per instruction
constexpr int TILE_BYTES_1 = 128 * BK * 4;     // one 128 x 32 fp32 operand tile
constexpr int NUM_THREADS = 256;

// Two shared-memory rings.  The operand ring (TMA destination) is deep to hide TMA latency; the 3xTF32 "lo"
// tiles only live from the split to the MMAs that consume them, so their ring is shallow.
//   MODE 3: 3x with the A operand in TENSOR MEMORY (tcgen05.mma TS form): the split warps read the TMA tile once
//           and write A_hi | A_lo into TMEM (tcgen05.st); the MMAs then read only B / B_lo from shared memory.
//           (MODE 2 re-reads the A slices from shared memory for each of the 3 MMAs and is bound by the shared-
//           memory port: ~176 KB of smem traffic per k-tile vs ~112 KB here.)  Needs A K-major and a B_lo plane.
// OCC = 2: half-depth rings so that TWO CTAs are resident per SM - the prologue (barrier init, TMEM alloc, first TMA
// round trip) and the epilogue of one tile overlap the main loop of the other (the kernel is not persistent).
//   MODE 4: bf16x3 - the same error compensation on the kind::f16 tensor path, which runs at TWICE the tf32 rate and
//           halves the weight bytes: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits kept), D +=
//           A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  A: fp32 K-major tile by TMA, split by the same 128 threads into PACKED
//           bf16 pairs in tensor memory (16 + 16 columns per 32-k tile instead of 32 + 32).  B: a pre-split, TRANSPOSED
//           bf16 plane maintained next to the weights (nar_pack_bf16x3): row n holds, per block of 32 k, the 32 hi
//           values followed by the 32 lo values = one 128-byte swizzle row, so ONE K-major TMA box brings both halves
//           and a stage is 32 KB instead of 48 KB.  Measured motivation: the 3xTF32 forward GEMMs sit on the SM's
//           L2->shared-memory ingest (~32 B/clk per SM: 289 us = 2.67 GB / 148 SMs), not on the tensor pipe.
template <int MODE, int TM, int TN, int OCC = 1> struct Cfg {
  static constexpr bool BF16 = MODE == 4;
  static constexpr bool SPLIT3 = MODE != 0;
  static constexpr bool BLO = MODE == 2 || MODE == 3;
  static constexpr bool ATMEM = MODE == 3 || MODE == 4;
  static constexpr int A_SLOT_COLS = BF16 ? 32 : 64;                        // tensor-memory columns of one A_hi | A_lo k-tile (per 128 rows)
  static constexpr int A_BYTES = TM * TILE_BYTES_1;
  static constexpr int B_BYTES = TN * TILE_BYTES_1;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES * (BLO ? 2 : 1);     // one operand stage (A | B [| B_lo])
  static constexpr int LO_STAGE_BYTES = (SPLIT3 && !ATMEM) ? (BLO ? A_BYTES : A_BYTES + B_BYTES) : 0;
  static constexpr int STAGES = OCC == 2 ? ((MODE == 3 || TM * TN == 2) ? 2 : 3)
                                         : (SPLIT3 ? (BLO ? (TM == 2 ? 3 : 4) : 5) : (TM * TN == 4 ? 3 : 6));  // operand ring depth
  static constexpr int LO_STAGES = SPLIT3 ? (ATMEM ? (BF16 ? 4 : ((TM == 2 || OCC == 2) ? 2 : 4)) : 2) : 0;   // lo ring depth (shared memory, or TMEM for MODE 3 / 4)
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES + (ATMEM ? 0 : LO_STAGES * LO_STAGE_BYTES);
  static constexpr int SMEM_BYTES = TILE_BYTES + 256 + 1024;                // tiles + barriers + align slack
  static constexpr int TMEM_COLS = ATMEM ? (OCC == 2 ? 256 : 512) : TM * TN * 128;             // accumulators (+ A ring: TM x (A_hi | A_lo) of 32 columns per stage)
  static constexpr int TMEM_A_BASE = TM * TN * 128;                         // MODE 3: A ring starts after the accumulators
  static_assert((TM == 1 && TN == 1) || MODE == 0 || (MODE == 3 && TM == 2 && TN == 1), "tile shapes: 128x128; 256x256 single pass; 256x128 with A in TMEM");
  static_assert(!ATMEM || TMEM_A_BASE + LO_STAGES * TM * A_SLOT_COLS <= TMEM_COLS, "tensor memory budget");
  static_assert(OCC == 1 || (TM * TN == 1 && (MODE == 0 || MODE == 3 || MODE == 4)) || (TM * TN == 2 && MODE == 0),
                "two CTAs per SM: 128x128 tiles (single pass or A-in-TMEM), 256x128 or 128x256 single pass");
};

struct Params {
  int64_t M, N, K;
  float* D; int64_t ldd;
  const float* bias;
  const float* aux; int64_t ld_aux;
  int act, dact, accumulate;
  int k_tiles_per_split;
  int n_tiles;                 // blockIdx.x = m_blk * n_tiles + n_blk (N fastest: CTAs sharing an A tile run together)
  int cluster;                 // 1: launched as clusters of 2 CTAs that own M-tiles 2j / 2j+1 of the SAME N-tile; each CTA
                               // fetches half of the B tile(s) and TMA-multicasts it into both CTAs' shared memory
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  // bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast form: the box lands at the same shared-memory offset in every CTA of `mask`, and signals the mbarrier at
// the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (TS form): D[tmem] (+)= A[tmem, 128 lanes x 8 columns] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// kind::f16 (bf16 operands, fp32 accumulate), A from tensor memory: packed pairs, 8 columns per K = 16 instruction
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// x -> (bf16(x), bf16(x - bf16(x))) for two consecutive k values, each pair packed low = even k
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {     // arrive on `bar` in every CTA of mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout (2 = SWIZZLE_128B)
// K-major tile  (rows x 128 B, 8-row atoms 1024 B apart): layout 2 (SWIZZLE_128B), LBO unused (1), SBO = 1024.
// MN-major tile: 32-bit operands have ONE legal MN-major layout, SWIZZLE_128B_BASE32B (layout 1; CUTLASS
// Layout_MN_SW128_32B_Atom = Swizzle<2,5,2>, 4 k-rows x 128 B of MN per atom; TMA SWIZZLE_128B_ATOM_32B):
// k rows of 128 B, 4-row atoms 512 B apart (SBO), 32-wide MN chunks one TMA box = 4096 B apart (LBO).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  const uint64_t lbo = MN_MAJOR ? (4096u >> 4) : 1u;
  const uint64_t sbo = MN_MAJOR ? (512u >> 4) : (1024u >> 4);
  const uint64_t layout = MN_MAJOR ? 1ull : 2ull;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ float tf32_lo(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}


// instruction descriptor (UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10),
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
template <bool A_MN, bool B_MN, int N>
__device__ __forceinline__ constexpr uint32_t make_idesc_n() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// kind::f16 with bf16 operands: a/b format BF16 (1<<7, 1<<10), both K-major
template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Epilogue element work for 4 consecutive columns of one row: bias / activation / activation-derivative / store.
// Called after the shared-memory transpose, so the 8 lanes that share a row touch 128 contiguous bytes of D / aux.
__device__ __forceinline__ void epilogue_store4(const Params& p, float4 v, int64_t row, int64_t col, const float4& a) {
  float* d = p.D + row * p.ldd + col;
  if (col + 4 <= p.N) {
    if (p.bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (p.act) { v.x = apply_act(v.x, p.act); v.y = apply_act(v.y, p.act); v.z = apply_act(v.z, p.act); v.w = apply_act(v.w, p.act); }
    if (p.dact) {        // `a` = aux[row, col..col+3], prefetched by the caller for the whole chunk
      v.x *= act_grad_from_output(a.x, p.dact); v.y *= act_grad_from_output(a.y, p.dact);
      v.z *= act_grad_from_output(a.z, p.dact); v.w *= act_grad_from_output(a.w, p.dact);
    }
    if (p.accumulate) atomicAdd(reinterpret_cast<float4*>(d), v);        // red.global.add.v4.f32
    else *reinterpret_cast<float4*>(d) = v;
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col + j < p.N) {
        float x = e[j];
        if (p.bias) x += p.bias[col + j];
        x = apply_act(x, p.act);
        if (p.dact) x *= act_grad_from_output(p.aux[row * p.ld_aux + col + j], p.dact);
        if (p.accumulate) atomicAdd(d + j, x); else d[j] = x;
      }
    }
  }
}

constexpr int STAGE_LD = 36;      // floats per staged accumulator row (32 + 4: 16-byte aligned, conflict-free)

// TMA for one operand tile of `rows_mn` (= 128*T) MN rows at MN coordinate mn0, K element k_elem
template <bool MN_MAJOR, int T>
__device__ __forceinline__ void load_operand(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int mn0, int k_elem) {
  if (!MN_MAJOR) {
    tma_load_2d(dst, map, bar, k_elem, mn0);                       // one box: 128*T rows x 128 B
  } else {
#pragma unroll
    for (int i = 0; i < 4 * T; ++i) tma_load_2d(dst + i * 4096, map, bar, mn0 + i * 32, k_elem);   // boxes of 32 MN x 32 k
  }
}

// cluster form: this CTA (rank r of 2) fetches HALF of the tile and multicasts it to both CTAs.  K-major: `map` has a
// box of 64*T rows; MN-major: 2*T of the 4*T boxes.
template <bool MN_MAJOR, int T>
__device__ __forceinline__ void load_operand_mc(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int mn0, int k_elem, int r) {
  if (!MN_MAJOR) {
    tma_load_2d_mc(dst + r * (64 * T * 128), map, bar, k_elem, mn0 + r * 64 * T, (uint16_t)3);
  } else {
#pragma unroll
    for (int i = 0; i < 2 * T; ++i)
      tma_load_2d_mc(dst + (r * 2 * T + i) * 4096, map, bar, mn0 + (r * 2 * T + i) * 32, k_elem, (uint16_t)3);
  }
}

// ---------------------------------------------------------------- kernel
template <bool A_MN, bool B_MN, int MODE, int TM, int TN, int OCC = 1>
__global__ void __launch_bounds__(NUM_THREADS, OCC)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_blo, const Params p) {
  using C = Cfg<MODE, TM, TN, OCC>;
  constexpr bool SPLIT3 = C::SPLIT3, BLO = C::BLO, ATMEM = C::ATMEM, BF16 = C::BF16;
  static_assert(!BF16 || (!A_MN && !B_MN), "bf16x3: A K-major fp32, B the transposed (K-major) bf16 plane");
  constexpr int LS = C::LO_STAGES > 0 ? C::LO_STAGES : 1;        // lo ring depth (2 in shared memory, 4 in TMEM)
  static_assert(!ATMEM || !A_MN, "A in tensor memory must be K-major");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* tiles = smem;
  uint8_t* lo_tiles = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::TILE_BYTES);
  uint64_t* full = bars;                          // [STAGES]    TMA landed
  uint64_t* empty = bars + C::STAGES;             // [STAGES]    MMAs that read the operand stage retired
  uint64_t* xf = bars + 2 * C::STAGES;            // [4] lo tiles written (128 arrivals)
  uint64_t* lo_empty = xf + 4;                    // [4] MMAs that read the lo stage retired
  uint64_t* tmem_full = xf + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf + 9);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cluster mode: CTAs 2j, 2j+1 (= cluster ranks 0, 1) own M-tiles 2*(j / n_tiles) + rank of N-tile j % n_tiles
  const int crank = p.cluster ? (int)(blockIdx.x & 1u) : 0;
  const int n_blk = p.cluster ? (int)((blockIdx.x >> 1) % p.n_tiles) : (int)(blockIdx.x % p.n_tiles);
  const int m_blk = p.cluster ? (int)((blockIdx.x >> 1) / p.n_tiles) * 2 + crank : (int)(blockIdx.x / p.n_tiles);
  const int k_tiles_total = (int)((p.K + BK - 1) / BK);
  const int kt0 = blockIdx.y * p.k_tiles_per_split;
  const int kt1 = min(kt0 + p.k_tiles_per_split, k_tiles_total);
  const int num_kt = kt1 - kt0;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (BLO) tma_prefetch_desc(&tmap_blo);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      // MODE 3: the A tile is released by the 128 split threads, B by the MMA commit; cluster: the peer CTA's MMAs read
      // the B half this CTA multicasts, so its commit releases the stage too
      mbar_init(&empty[s], (ATMEM ? 129 : 1) + (p.cluster ? 1 : 0));
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&xf[s], 128);
      mbar_init(&lo_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.cluster) cluster_sync_all();              // the peer's barriers exist before anything is multicast at them
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % C::STAGES;
        const uint32_t ph = (uint32_t)(kt / C::STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], C::STAGE_BYTES);
        const int k_elem = (kt0 + kt) * BK;
        const uint32_t a_dst = smem_u32(tiles + s * C::STAGE_BYTES);
        const uint32_t b_dst = a_dst + C::A_BYTES;
        load_operand<A_MN, TM>(a_dst, &tmap_a, &full[s], m_blk * BM * TM, k_elem);
        if (BF16) {
          // one box: 128 n-rows x 128 B = the 32 hi then the 32 lo bf16 of k-tile (kt0 + kt)
          tma_load_2d(b_dst, &tmap_b, &full[s], (kt0 + kt) * 64, n_blk * BN * TN);
        } else if (p.cluster) {
          load_operand_mc<B_MN, TN>(b_dst, &tmap_b, &full[s], n_blk * BN * TN, k_elem, crank);
          if (BLO) load_operand_mc<B_MN, TN>(b_dst + C::B_BYTES, &tmap_blo, &full[s], n_blk * BN * TN, k_elem, crank);
        } else {
          load_operand<B_MN, TN>(b_dst, &tmap_b, &full[s], n_blk * BN * TN, k_elem);
          if (BLO) load_operand<B_MN, TN>(b_dst + C::B_BYTES, &tmap_blo, &full[s], n_blk * BN * TN, k_elem);
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_n<A_MN, B_MN, BN * TN>();
      constexpr uint32_t a_kstep = A_MN ? 1024u : (uint32_t)(UMMA_K * 4);
      constexpr uint32_t b_kstep = B_MN ? 1024u : (uint32_t)(UMMA_K * 4);
      for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % C::STAGES;
        const uint32_t ph = (uint32_t)(kt / C::STAGES) & 1u;
        const int ls = SPLIT3 ? (kt % LS) : 0;
        mbar_wait(&full[s], ph);
        if (SPLIT3) mbar_wait(&xf[ls], (uint32_t)(kt / LS) & 1u);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(tiles + s * C::STAGE_BYTES);
        const uint32_t b_hi = a_hi + C::A_BYTES;
        const uint32_t a_lo = smem_u32(lo_tiles + ls * C::LO_STAGE_BYTES);
        const uint32_t b_lo = BLO ? (b_hi + C::B_BYTES) : (a_lo + C::A_BYTES);
        if (BF16) {
          constexpr uint32_t idesc16 = make_idesc_bf16<BN * TN>();
          const uint32_t acc = tmem_base;
          const uint32_t ta = tmem_base + (uint32_t)(C::TMEM_A_BASE + ls * C::A_SLOT_COLS);      // 16 columns hi | 16 columns lo
#pragma unroll
          for (int k = 0; k < 2; ++k) {                        // two K = 16 steps per 32-k tile; 32 B of k per step in the swizzled row
            const uint64_t db = make_smem_desc<false>(b_hi + k * 32);          // hi half of the row: bytes [0, 64)
            const uint64_t dbl = make_smem_desc<false>(b_hi + 64 + k * 32);    // lo half: bytes [64, 128)
            umma_bf16_ts(acc, ta + 16 + k * 8, db, idesc16, (kt > 0 || k > 0) ? 1u : 0u);       // A_lo * B_hi
            umma_bf16_ts(acc, ta + k * 8, dbl, idesc16, 1u);                                     // A_hi * B_lo
            umma_bf16_ts(acc, ta + k * 8, db, idesc16, 1u);                                      // A_hi * B_hi
          }
          umma_commit(&empty[s]);
          umma_commit(&lo_empty[ls]);
          continue;
        }
        if (ATMEM) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t db = make_smem_desc<B_MN>(b_hi + k * b_kstep);
            const uint64_t dbl = make_smem_desc<B_MN>(b_lo + k * b_kstep);
#pragma unroll
            for (int tm = 0; tm < TM; ++tm) {
              const uint32_t acc = tmem_base + (uint32_t)(tm * BN * TN);
              const uint32_t ta_hi = tmem_base + (uint32_t)(C::TMEM_A_BASE + (ls * TM + tm) * 64);   // 32 columns A_hi | 32 columns A_lo
              umma_tf32_ts(acc, ta_hi + 32 + k * UMMA_K, db, idesc, (kt > 0 || k > 0) ? 1u : 0u);    // A_lo * B_hi
              umma_tf32_ts(acc, ta_hi + k * UMMA_K, dbl, idesc, 1u);                                   // A_hi * B_lo
              umma_tf32_ts(acc, ta_hi + k * UMMA_K, db, idesc, 1u);                                    // A_hi * B_hi
            }
          }
          if (p.cluster) umma_commit_mc(&empty[s], (uint16_t)3); else umma_commit(&empty[s]);
          umma_commit(&lo_empty[ls]);
          continue;
        }
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t db = make_smem_desc<B_MN>(b_hi + k * b_kstep);
#pragma unroll
          for (int tm = 0; tm < TM; ++tm) {
            const uint32_t acc = tmem_base + (uint32_t)(tm * BN * TN);         // accumulator tm: columns [tm*128*TN, ...)
            const uint32_t a_tm = a_hi + tm * TILE_BYTES_1 + k * a_kstep;       // MN-major: 4 boxes of 4096 B = TILE_BYTES_1
            const uint64_t da = make_smem_desc<A_MN>(a_tm);
            const uint32_t first = (kt > 0 || k > 0) ? 1u : 0u;
            if (SPLIT3) {
              const uint64_t dal = make_smem_desc<A_MN>(a_lo + tm * TILE_BYTES_1 + k * a_kstep);
              const uint64_t dbl = make_smem_desc<B_MN>(b_lo + k * b_kstep);
              umma_tf32(acc, dal, db, idesc, first);     // small terms first
              umma_tf32(acc, da, dbl, idesc, 1u);
              umma_tf32(acc, da, db, idesc, 1u);
            } else {
              umma_tf32(acc, da, db, idesc, first);
            }
          }
        }
        if (p.cluster) umma_commit_mc(&empty[s], (uint16_t)3); else umma_commit(&empty[s]);     // frees the operand stage when these MMAs retire
        if (SPLIT3) umma_commit(&lo_empty[ls]);
      }
      umma_commit(tmem_full);       // accumulators complete
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;            // TMEM lane quadrant == warp_idx % 4
    if (SPLIT3) {
      // ===== operand split: lo = x - trunc_tf32(x) =====
      const int te = threadIdx.x - 128;
      for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % C::STAGES;
        const uint32_t ph = (uint32_t)(kt / C::STAGES) & 1u;
        const int ls = kt % LS;
        mbar_wait(&lo_empty[ls], ((uint32_t)(kt / LS) & 1u) ^ 1u);    // MMAs of k-tile kt-LS no longer read this lo stage
        mbar_wait(&full[s], ph);
        if (BF16) {
          // thread = accumulator lane = A row: 32 fp32 k-values -> 16 packed hi pairs + 16 packed lo pairs in tensor memory
          const int row = ew * 32 + lane;
          const uint8_t* arow = tiles + s * C::STAGE_BYTES + row * 128;
          uint32_t hi_p[16], lo_p[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
            split_bf16x2(v.x, v.y, hi_p[2 * c], lo_p[2 * c]);
            split_bf16x2(v.z, v.w, hi_p[2 * c + 1], lo_p[2 * c + 1]);
          }
          const uint32_t ta = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(C::TMEM_A_BASE + ls * C::A_SLOT_COLS);
          tmem_st_32x32b_x16(ta, hi_p);
          tmem_st_32x32b_x16(ta + 16, lo_p);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&xf[ls]);
          mbar_arrive(&empty[s]);        // this thread is done with the A tile of the operand stage
          continue;
        }
        if (ATMEM) {
          // thread = accumulator lane = A row: read the row's 32 k-values (8 swizzled 16-byte chunks) from the TMA tile
          // and write A_hi | A_lo into tensor memory; the MMAs never touch the A tile in shared memory
#pragma unroll
          for (int tm = 0; tm < TM; ++tm) {
            const int row = tm * 128 + ew * 32 + lane;
            const uint8_t* arow = tiles + s * C::STAGE_BYTES + row * 128;
            uint32_t hi_r[32], lo_r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
              hi_r[4 * c + 0] = __float_as_uint(v.x); hi_r[4 * c + 1] = __float_as_uint(v.y);
              hi_r[4 * c + 2] = __float_as_uint(v.z); hi_r[4 * c + 3] = __float_as_uint(v.w);
              lo_r[4 * c + 0] = __float_as_uint(tf32_lo(v.x)); lo_r[4 * c + 1] = __float_as_uint(tf32_lo(v.y));
              lo_r[4 * c + 2] = __float_as_uint(tf32_lo(v.z)); lo_r[4 * c + 3] = __float_as_uint(tf32_lo(v.w));
            }
            const uint32_t ta = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(C::TMEM_A_BASE + (ls * TM + tm) * 64);
            tmem_st_32x32b_x32(ta, hi_r);
            tmem_st_32x32b_x32(ta + 32, lo_r);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&xf[ls]);
          mbar_arrive(&empty[s]);        // this thread is done with the A tile of the operand stage
          continue;
        }
        const float4* hi = reinterpret_cast<const float4*>(tiles + s * C::STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(lo_tiles + ls * C::LO_STAGE_BYTES);
        // A (and, without a B_lo plane in HBM, B: the two tiles are contiguous) -> lo.  All loads are issued
        // before the first use so the split costs one shared-memory round trip per k-tile.
        constexpr int NV = (SPLIT3 && !ATMEM) ? C::LO_STAGE_BYTES / 16 / 128 : 1;   // float4 per thread: 8 (A only) or 16 (A and B)
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = hi[te + i * 128];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i].x = tf32_lo(v[i].x); v[i].y = tf32_lo(v[i].y); v[i].z = tf32_lo(v[i].z); v[i].w = tf32_lo(v[i].w);
          lo[te + i * 128] = v[i];
        }
        fence_proxy_async_smem();
        mbar_arrive(&xf[ls]);
      }
    }
    // ===== epilogue =====
    // tcgen05.ld 32x32b hands thread i row i of the warp's 32-row slab; a direct store would touch 32 different
    // 128-byte lines per instruction.  Each 32x32 chunk is transposed through shared memory (the operand ring is
    // idle by now) so that 8 lanes cover one row's 128 contiguous bytes: coalesced D stores and aux loads.
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(tiles) + ew * (32 * STAGE_LD);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
    for (int tm = 0; tm < TM; ++tm) {
      const int64_t row_base = ((int64_t)m_blk * TM + tm) * BM + ew * 32;
      for (int c = 0; c < BN * TN; c += 32) {
        const int64_t col0 = (int64_t)n_blk * BN * TN + c;
        if (col0 >= p.N) break;               // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(tm * BN * TN + c), r);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * STAGE_LD + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        // aux (forward output of the layer being differentiated) for the whole chunk first: 8 independent 128-bit
        // loads in flight instead of a load -> use -> store chain per row (D may alias aux, so the compiler cannot
        // hoist them itself; every element is read before it is overwritten, rows of different iterations differ)
        float4 av[8];
        if (p.dact) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int64_t row = row_base + it * 4 + rsub;
            av[it] = (row < p.M && col0 + c4 + 4 <= p.N)
                         ? *reinterpret_cast<const float4*>(p.aux + row * p.ld_aux + col0 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = it * 4 + rsub;
          const float4 v = *reinterpret_cast<const float4*>(stage + rl * STAGE_LD + c4);
          if (row_base + rl < p.M) epilogue_store4(p, v, row_base + rl, col0 + c4, av[it]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();              // stay resident until the peer's last multicast commit has landed here
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// operand with logical shape [mn, k]; kmajor: ptr[mn*ld + k] else ptr[k*ld + mn]; T = tile multiplier
static int make_operand_map(const nar_ctx* ctx, CUtensorMap* map, const float* ptr, int64_t mn, int64_t k, int64_t ld, bool kmajor, int T,
                            int box_rows = 128) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld & 3) != 0 || ld <= 0) return NAR_ERR_INVALID;
  cuuint64_t dims[2]; cuuint64_t strides[1]; cuuint32_t box[2]; cuuint32_t estr[2] = {1, 1};
  if (kmajor) { dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)mn; box[0] = BK; box[1] = (cuuint32_t)(box_rows * T); }
  else        { dims[0] = (cuuint64_t)mn; dims[1] = (cuuint64_t)k; box[0] = 32; box[1] = BK; }
  strides[0] = (cuuint64_t)ld * 4;
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NAR_OK : NAR_ERR_INVALID;
}

// the pre-split bf16 weight plane of MODE 4: [n_rows, ld] bf16, K-major, 64 elements (128 B) of it per 32-k tile
static int make_bf16_plane_map(const nar_ctx* ctx, CUtensorMap* map, const void* ptr, int64_t n_rows, int64_t k_tiles, int64_t ld) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld & 7) != 0 || ld < k_tiles * 64) return NAR_ERR_INVALID;
  cuuint64_t dims[2] = {(cuuint64_t)(k_tiles * 64), (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NAR_OK : NAR_ERR_INVALID;
}

// nar_pack_bf16x3: W [K, N] fp32 (row stride ldw) -> plane [N, ld_out] bf16 (see MODE 4).  32 x 32 tiles through shared
// memory: reads coalesced along n, writes 64 contiguous bytes (32 hi or 32 lo values of one n) per half warp.
struct PackDesc { const float* W; uint16_t* out; int K, N, ldw, ld_out; };
constexpr int MAX_PACK = 32;

__global__ void __launch_bounds__(256)
pack_bf16x3_kernel(const PackDesc* __restrict__ descs) {
  __shared__ float tile[32][33];
  const PackDesc d = descs[blockIdx.y];
  const int kb_n = (d.K + 31) / 32, nb_n = (d.N + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int t = blockIdx.x; t < kb_n * nb_n; t += gridDim.x) {
    const int kb = t / nb_n, nb = t - kb * nb_n;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int k = kb * 32 + i, n = nb * 32 + tx;
      tile[i][tx] = (k < d.K && n < d.N) ? d.W[(int64_t)k * d.ldw + n] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {                             // i = n within the tile, tx = k within the block
      const int n = nb * 32 + i;
      if (n < d.N) {
        const float x = tile[tx][i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
        uint16_t* o = d.out + (int64_t)n * d.ld_out + kb * 64 + tx;
        o[0] = *reinterpret_cast<const uint16_t*>(&h);
        o[32] = *reinterpret_cast<const uint16_t*>(&l);
      }
    }
  }
}

template <bool A_MN, bool B_MN, int MODE, int TM, int TN, int OCC = 1>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbl, const Params& p, dim3 grid, cudaStream_t st) {
  auto kern = gemm_tf32_kernel<A_MN, B_MN, MODE, TM, TN, OCC>;
  static bool attr_set = false;     // per instantiation
  if (!attr_set) {
    NAR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE, TM, TN, OCC>::SMEM_BYTES));
    attr_set = true;
  }
  if (p.cluster) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<MODE, TM, TN, OCC>::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    NAR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tbl, p));
    return NAR_OK;
  }
  kern<<<grid, NUM_THREADS, Cfg<MODE, TM, TN, OCC>::SMEM_BYTES, st>>>(ta, tb, tbl, p);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

}  // namespace gemm
}  // namespace nar

static int occ_env_early() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NAR_GEMM_OCC2"); v = e ? atoi(e) : 1; }
  return v;
}

extern "C" int nar_pack_bf16x3(const float* const* W, void* const* out, const int32_t* K, const int32_t* N, const int32_t* ldw,
                               const int32_t* ld_out, int n, void* descs_dev, void* stream) {
  using namespace nar::gemm;
  if (!W || !out || !K || !N || !ldw || !ld_out || !descs_dev || n <= 0 || n > MAX_PACK) return NAR_ERR_INVALID;
  PackDesc h[MAX_PACK];
  int max_tiles = 1;
  for (int i = 0; i < n; ++i) {
    if (!W[i] || !out[i] || K[i] <= 0 || N[i] <= 0 || ld_out[i] < (K[i] + 31) / 32 * 64) return NAR_ERR_INVALID;
    h[i].W = W[i]; h[i].out = static_cast<uint16_t*>(out[i]); h[i].K = K[i]; h[i].N = N[i]; h[i].ldw = ldw[i]; h[i].ld_out = ld_out[i];
    const int t = ((K[i] + 31) / 32) * ((N[i] + 31) / 32);
    max_tiles = t > max_tiles ? t : max_tiles;
  }
  // the descriptor table is written once per distinct set (callers keep it; stream-ordered copy from a pageable buffer
  // would be a sync, so it goes through a kernel-argument-sized async memcpy only when it changed)
  static PackDesc last[MAX_PACK]; static int last_n = 0; static void* last_dev = nullptr;
  if (last_dev != descs_dev || last_n != n || memcmp(last, h, sizeof(PackDesc) * n) != 0) {
    NAR_CHECK_CUDA(cudaMemcpy(descs_dev, h, sizeof(PackDesc) * n, cudaMemcpyHostToDevice));
    memcpy(last, h, sizeof(PackDesc) * n); last_n = n; last_dev = descs_dev;
  }
  dim3 grid((unsigned)(max_tiles > 296 ? 296 : max_tiles), (unsigned)n);
  pack_bf16x3_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const PackDesc*>(descs_dev));
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_gemm_tf32(nar_ctx* ctx, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int a_kmajor,
                             const float* B, int64_t ldb, int b_kmajor, float* D, int64_t ldd,
                             const nar_gemm_epilogue* epi, void* stream) {
  using namespace nar::gemm;
  if (!ctx || !ctx->encode_tiled) return NAR_ERR_NO_DEVICE;
  if (!A || !D || !epi || (!B && epi->precision != 4)) return NAR_ERR_INVALID;
  if (M <= 0 || N <= 0 || K <= 0) return NAR_OK;     // empty problem: nothing to do
  if ((ldd & 3) != 0 || (reinterpret_cast<uintptr_t>(D) & 15u) != 0) return NAR_ERR_INVALID;
  if (epi->bias && (reinterpret_cast<uintptr_t>(epi->bias) & 15u) != 0) return NAR_ERR_INVALID;
  if (epi->dact && (!epi->aux || (epi->ld_aux & 3) != 0 || (reinterpret_cast<uintptr_t>(epi->aux) & 15u) != 0)) return NAR_ERR_INVALID;
  if (epi->precision != 1 && epi->precision != 3 && epi->precision != 4) return NAR_ERR_INVALID;
  const bool bf16 = epi->precision == 4;
  if (bf16 && (!a_kmajor || !epi->b_bf16 || epi->accumulate || epi->split_k > 1)) return NAR_ERR_INVALID;
  const bool blo = epi->precision == 3 && epi->b_lo != nullptr;
  const int mode = bf16 ? 4 : (epi->precision == 1 ? 0 : (blo ? (a_kmajor ? 3 : 2) : 1));
  // 256x256 CTA tiles for the big single-pass GEMMs (L2-bound with 128x128 tiles); 128x128 otherwise
  const bool big = (double)M * (double)N * (double)K >= 2e9;
  // (256x128 tiles with A in TMEM are implemented and validated but measured ~10 % slower than 128x128 for MODE 3)
  // big single-pass GEMMs: 0 = 256x256 tiles, one CTA per SM; 1 = 256x128 tiles, 2 = 128x256 tiles, two CTAs per SM
  // (measured, 24000x1024x1024: dgrad 135 / 122 / 111 us, dgrad + activation derivative 178 / 143 / 134 us, wgrad
  // 143 / 144 / 135 us: one 128x256x8 MMA reads 12 KB of shared memory per 2x the math of a 128x128x8 one (8 KB), and
  // the second resident CTA hides the epilogue)
  static int big_env = -1;
  if (big_env < 0) { const char* e = getenv("NAR_GEMM_BIG_OCC2"); big_env = e ? atoi(e) : 2; }
  const bool bigtile = mode == 0 && M >= 256 && N >= 256 && big;
  const int TM = (bigtile && big_env != 2) ? 2 : 1;
  const int TN = (bigtile && big_env != 1) ? 2 : 1;
  const int64_t n_tiles = (N + BN * TN - 1) / (BN * TN), m_tiles = (M + BM * TM - 1) / (BM * TM);
  if (n_tiles * m_tiles > 0x7fffffffLL) return NAR_ERR_UNSUPPORTED;
  const int k_tiles = (int)((K + BK - 1) / BK);
  int split = epi->split_k;
  if (split <= 0) {          // auto: about two waves of CTAs, at least 8 k-tiles per split
    split = 1;
    if (epi->accumulate) {
      const int64_t want = (2 * (int64_t)ctx->sm_count + n_tiles * m_tiles - 1) / (n_tiles * m_tiles);
      const int64_t cap = k_tiles / 8 > 1 ? k_tiles / 8 : 1;
      split = (int)(want < cap ? want : cap);
      if (split < 1) split = 1;
    }
  }
  if (split > k_tiles) split = k_tiles;
  if (split > 1 && !epi->accumulate) return NAR_ERR_INVALID;
  int per = (k_tiles + split - 1) / split;
  split = (k_tiles + per - 1) / per;          // no empty splits
  // Optional (NAR_GEMM_CLUSTER=1): clusters of 2 CTAs along M share their B tiles through TMA multicast.  Validated
  // bit-identical, but measured 3-6 % SLOWER on B200 for every shape of the step (24000x1024x1024: 3xTF32 forward 379 ->
  // 389 us, single pass 194 -> 201 us, 256x256 dgrad 137 -> 142 us): halving the B traffic out of L2 does not help, the
  // lock-step of the CTA pair costs a little.  Off by default.
  static int cluster_env = -1;
  if (cluster_env < 0) { const char* e = getenv("NAR_GEMM_CLUSTER"); cluster_env = e ? atoi(e) : 0; }
  const bool cluster = cluster_env != 0 && (mode == 0 || mode == 3) && m_tiles >= 2;
  const int64_t m_tiles_launch = cluster ? (m_tiles + 1) / 2 * 2 : m_tiles;      // an odd tail tile gets a partner that stores nothing
  CUtensorMap ta, tb;
  int rc = make_operand_map(ctx, &ta, A, M, K, lda, a_kmajor != 0, TM);
  if (rc) return rc;
  if (bf16) rc = make_bf16_plane_map(ctx, &tb, epi->b_bf16, N, k_tiles, epi->ld_bf16);
  else rc = make_operand_map(ctx, &tb, B, N, K, ldb, b_kmajor != 0, TN, cluster ? 64 : 128);
  if (rc) return rc;
  CUtensorMap tbl = tb;
  if (blo) {
    rc = make_operand_map(ctx, &tbl, epi->b_lo, N, K, ldb, b_kmajor != 0, TN, cluster ? 64 : 128);
    if (rc) return rc;
  }
  Params p;
  p.M = M; p.N = N; p.K = K; p.D = D; p.ldd = ldd; p.bias = epi->bias; p.aux = epi->aux; p.ld_aux = epi->ld_aux;
  p.act = epi->act; p.dact = epi->dact; p.accumulate = epi->accumulate; p.k_tiles_per_split = per;
  p.n_tiles = (int)n_tiles;
  p.cluster = cluster ? 1 : 0;
  dim3 grid((unsigned)(n_tiles * m_tiles_launch), (unsigned)split, 1);
  cudaStream_t st = as_stream(stream);
  const bool amn = !a_kmajor, bmn = !b_kmajor;
  if (mode == 4) {
    const bool occ2_bf = occ_env_early() != 0 && n_tiles * m_tiles_launch >= (int64_t)ctx->sm_count;
    if (occ2_bf) return launch<false, false, 4, 1, 1, 2>(ta, tb, tbl, p, grid, st);
    return launch<false, false, 4, 1, 1, 1>(ta, tb, tbl, p, grid, st);
  }
  static int occ_env = -1;
  if (occ_env < 0) { const char* e = getenv("NAR_GEMM_OCC2"); occ_env = e ? atoi(e) : 1; }
  // two CTAs per SM (half-depth rings) pay off when there are CTAs to pair up; a grid smaller than the GPU is latency
  // bound on its serial k-loop instead: deep rings (one CTA per SM, 4-6 stages of TMA prefetch) serve it better
  const bool occ2 = occ_env != 0 && !cluster && (n_tiles * m_tiles_launch * split >= (int64_t)ctx->sm_count || occ_env == 2);
#define NAR_GEMM_CASE(a, b) \
  if (amn == a && bmn == b) { \
    if (mode == 0 && TM == 1 && TN == 2) return launch<a, b, 0, 1, 2, 2>(ta, tb, tbl, p, grid, st); \
    if (mode == 0 && TM == 1 && occ2) return launch<a, b, 0, 1, 1, 2>(ta, tb, tbl, p, grid, st); \
    if (mode == 3 && TM == 1 && occ2) return launch<false, b, 3, 1, 1, 2>(ta, tb, tbl, p, grid, st); \
    if (mode == 0 && TM == 1) return launch<a, b, 0, 1, 1>(ta, tb, tbl, p, grid, st); \
    if (mode == 0 && TM == 2 && TN == 1) return launch<a, b, 0, 2, 1, 2>(ta, tb, tbl, p, grid, st); \
    if (mode == 0 && TM == 2) return launch<a, b, 0, 2, 2>(ta, tb, tbl, p, grid, st); \
    if (mode == 1) return launch<a, b, 1, 1, 1>(ta, tb, tbl, p, grid, st); \
    if (mode == 2) return launch<a, b, 2, 1, 1>(ta, tb, tbl, p, grid, st); \
    if (mode == 3 && TM == 1) return launch<false, b, 3, 1, 1>(ta, tb, tbl, p, grid, st); \
    if (mode == 3 && TM == 2) return launch<false, b, 3, 2, 1>(ta, tb, tbl, p, grid, st); \
  }
  NAR_GEMM_CASE(false, false) NAR_GEMM_CASE(false, true) NAR_GEMM_CASE(true, false) NAR_GEMM_CASE(true, true)
#undef NAR_GEMM_CASE
  return NAR_ERR_INVALID;
}
