// gemm_tcgen05.cu — the tensor-core path of tnb200_tensordot for bf16 / f16 / f32(tf32).
//
// One persistent, warp-specialised kernel per launch (B200, sm_100a):
//   warp 0  : TMA producer  — cp.async.bulk.tensor (3-D tiled maps built over the *strided
//             operand views*, so the tensordot's transpose is performed by the TMA engine),
//             SWIZZLE_128B tiles into an N-stage shared-memory ring, mbarrier complete_tx.
//   warp 1  : MMA issuer    — one elected thread issues tcgen05.mma (cta_group::1, M=128,
//             N=BN<=256, K=32 bytes) from shared-memory descriptors into a TMEM accumulator;
//             tcgen05.commit releases ring slots / publishes the accumulator.
//   warp 2  : TMEM allocator (512 columns = two accumulator buffers, so the epilogue of tile i
//             overlaps the main loop of tile i+1).
//   warps 4-7: epilogue     — tcgen05.ld (32 lanes x 32 columns per warp), convert, vectorised
//             stores into the row-major output.
// Either operand may be K-major (unit stride along the contracted mode) or MN-major (unit
// stride along the free mode): both are native UMMA layouts, selected in the instruction
// descriptor — no operand is ever transposed in memory on this path.
// Ragged edges in M, N, K are handled by TMA out-of-bounds zero fill + predicated stores.
#include "gemm.cuh"
#include <cuda.h>
#include <mutex>
#include <vector>

namespace tnb {

// ------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (visible as a CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>  // 0: f16/bf16 (kind::f16), 1: tf32
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, sm_100 version = 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)(layout & 7) << 61; // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

struct TcParams {
  int64_t M, N, K, batch;
  int BN, num_kb, stages, out_kind;   // out_kind: 0 bf16, 1 f16, 2 f32
  int64_t tiles_m, tiles_n, num_tiles;
  int a_mn, b_mn;                     // operand is MN-major
  // multi-mode operands: extent of the INNER free / contracted mode (0 = the group is one mode).
  // TMA dims are always (K-major)  [k_in, f_in, f_out, k_out, batch]
  //                     (MN-major) [f_in, k_in, k_out, f_out, batch]
  uint32_t a_fe, a_ke, b_fe, b_ke;
  uint32_t idesc;
  void* C; int64_t c_sm, c_sb;        // row stride / batch stride of the output tile rows
  int64_t c_cs;                       // column stride: 1 (normal) or the original row stride (swap-AB: the
                                      // kernel computes C^T tiles and stores them transposed)
  int vec_ok;
  int c_hint;                         // 0: default L2 policy; 1: evict_last (a chained intermediate: keep it until its consumer has read it)
};

constexpr int kBM = 128;
constexpr int kRowBytes = 128;        // one swizzle row = BK elements
constexpr int kThreads = 256;


__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_global_v4_hint(void* dst, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}

// Store one 32-column chunk of an accumulator row held by this lane (row = TMEM lane).  Normal mode: the
// lane's 32 values are contiguous in C (vector stores).  swap-AB (c_cs != 1): the lane's row is an ORIGINAL
// column, so consecutive lanes write consecutive addresses (coalesced across the warp).
__device__ __forceinline__ void epilogue_store_chunk(const TcParams& p, const uint32_t* r, int64_t bi, int64_t row, int64_t col0) {
  if (p.c_cs != 1) {
    if (row < p.M) {
      const int ncol = (p.N - col0) >= 32 ? 32 : (int)(p.N - col0 > 0 ? p.N - col0 : 0);
      const int64_t base = bi * p.c_sb + row * p.c_sm + col0 * p.c_cs;
      if (p.out_kind == 2) {
        float* dst = (float*)p.C + base;
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c < ncol) dst[(int64_t)c * p.c_cs] = __uint_as_float(r[c]);
      } else {
        uint16_t* dst = (uint16_t*)p.C + base;
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c < ncol) {
          const float v = __uint_as_float(r[c]);
          uint16_t h;
          if (p.out_kind == 0) { __nv_bfloat16 b = __float2bfloat16_rn(v); h = *(uint16_t*)&b; }
          else { __half b = __float2half_rn(v); h = *(uint16_t*)&b; }
          dst[(int64_t)c * p.c_cs] = h;
        }
      }
    }
  } else if (row < p.M && col0 < p.N) {
    const int64_t off = bi * p.c_sb + row * p.c_sm + col0;
    const int ncol = (p.N - col0) >= 32 ? 32 : (int)(p.N - col0);
    if (p.out_kind == 2) {
      float* dst = (float*)p.C + off;
      if (ncol == 32 && p.vec_ok) {
#pragma unroll
        for (int v = 0; v < 8; ++v)
          ((float4*)dst)[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                          __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c < ncol) dst[c] = __uint_as_float(r[c]);
      }
    } else {
      uint16_t* dst = (uint16_t*)p.C + off;
      uint32_t pk[16];
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        float lo = __uint_as_float(r[2 * v]), hi = __uint_as_float(r[2 * v + 1]);
        if (p.out_kind == 0) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); pk[v] = *(uint32_t*)&h; }
        else { __half2 h = __floats2half2_rn(lo, hi); pk[v] = *(uint32_t*)&h; }
      }
      if (ncol == 32 && p.vec_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) ((uint4*)dst)[v] = make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c < ncol) dst[c] = (uint16_t)(pk[c >> 1] >> ((c & 1) * 16));
      }
    }
  }
}
constexpr int kSlabRow = 144;                 // 128 data bytes + 16 pad: conflict-free 16-byte rows
constexpr int kSlabBytes = 32 * kSlabRow;     // one warp's staging slab (32 accumulator rows)

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Drain one accumulator (BN columns) of this warp's 32 TMEM lanes.
// Fast path (row-major output, aligned, N a multiple of the 16-byte vector): 128 bytes of every row are
// converted, staged in a per-warp shared-memory slab and written back so that 8 consecutive lanes cover one
// row's 128 bytes — every global store instruction writes 4 full cache lines (the direct path writes
// 32 half-empty sectors).  Two TMEM chunks are always in flight in separate register sets.
__device__ __forceinline__ void epilogue_drain(const TcParams& p, uint32_t taddr0, int BN, int64_t bi, int64_t row0,
                                               int lane, int64_t n0, uint32_t slab) {
  const int64_t row = row0 + lane;
  const int es = p.out_kind == 2 ? 4 : 2;
  const bool staged = p.c_cs == 1 && p.vec_ok && (p.N % (16 / es) == 0);
  if (!staged) {
    for (int j = 0; j < BN / 32; j += 2) {
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr0 + j * 32, ra);
      tmem_ld32(taddr0 + (j + 1) * 32, rb);
      tmem_ld_wait();
      epilogue_store_chunk(p, ra, bi, row, n0 + (int64_t)j * 32);
      epilogue_store_chunk(p, rb, bi, row, n0 + (int64_t)(j + 1) * 32);
    }
    return;
  }
  const uint32_t my_row = slab + (uint32_t)lane * kSlabRow;
  const int cols_per_pass = 128 / es;                       // 64 (16-bit) or 32 (fp32) columns = 128 bytes per row
  for (int j = 0; j < BN / 32; j += 2) {
    uint32_t ra[32], rb[32];
    tmem_ld32(taddr0 + j * 32, ra);
    tmem_ld32(taddr0 + (j + 1) * 32, rb);
    tmem_ld_wait();
    const int npass = es == 2 ? 1 : 2;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      if (pass >= npass) break;
      if (es == 2) {
        uint32_t pk[32];
#pragma unroll
        for (int v = 0; v < 16; ++v) {
          float lo = __uint_as_float(ra[2 * v]), hi = __uint_as_float(ra[2 * v + 1]);
          float lo2 = __uint_as_float(rb[2 * v]), hi2 = __uint_as_float(rb[2 * v + 1]);
          if (p.out_kind == 0) {
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); pk[v] = *(uint32_t*)&h;
            __nv_bfloat162 g = __floats2bfloat162_rn(lo2, hi2); pk[16 + v] = *(uint32_t*)&g;
          } else {
            __half2 h = __floats2half2_rn(lo, hi); pk[v] = *(uint32_t*)&h;
            __half2 g = __floats2half2_rn(lo2, hi2); pk[16 + v] = *(uint32_t*)&g;
          }
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) st_shared_v4(my_row + v * 16, pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
      } else {
        const uint32_t* r = pass == 0 ? ra : rb;
#pragma unroll
        for (int v = 0; v < 8; ++v) st_shared_v4(my_row + v * 16, r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
      }
      __syncwarp();
      const int64_t col0 = n0 + (int64_t)j * 32 + (es == 2 ? 0 : pass * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i, r = c >> 3, part = c & 7;
        const uint4 v = ld_shared_v4(slab + (uint32_t)r * kSlabRow + (uint32_t)part * 16);
        const int64_t grow = row0 + r, gcol = col0 + (int64_t)part * (16 / es);
        if (grow < p.M && gcol < p.N) {
          char* dst = (char*)p.C + (bi * p.c_sb + grow * p.c_sm + gcol) * es;
          if (p.c_hint) st_global_v4_hint(dst, v, l2_policy_evict_last());
          else *(uint4*)dst = v;
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;   // peer bit cleared: transaction bytes land on CTA 0's barrier
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4,
                                                     uint64_t policy) {
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4,
                                                   uint16_t cta_mask) {
  // multicast to every CTA of `cta_mask` (same CTA-relative destination); with cta_group::2 the transaction bytes are
  // signalled on the barrier of each destination's PAIR LEADER (peer bit cleared)
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mask(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// TMA producer, executed by ALL 32 lanes of warp 0: lane l owns "load slot" l of a stage — one
// 128-byte-wide box of A (slots 0..nA-1) or of B (slots nA..nA+nB-1).  Per-tile work (tile decode,
// free-mode coordinates of the slot) is done once per tile; inside the k loop a lane only advances
// its contracted-mode coordinates incrementally (no division) and issues its own TMA, so the boxes
// of a stage are issued in parallel instead of by one thread in sequence.
template <int KIND, bool TWO>
__device__ __forceinline__ void tma_producer(const CUtensorMap* tmA, const CUtensorMap* tmB, const TcParams& p, uint8_t* smem,
                                             uint32_t bar_base, int S, int stage_bytes, int b_rows, uint32_t rank,
                                             int64_t first, int64_t step, int lane) {
  constexpr int ES = KIND == 0 ? 2 : 4;
  constexpr int BK = kRowBytes / ES;
  constexpr int CHUNK = kRowBytes / ES;
  constexpr int A_BYTES = kBM * kRowBytes;
  const int nA = p.a_mn ? kBM / CHUNK : 1;
  const int nB = p.b_mn ? b_rows / CHUNK : 1;
  const bool mine = lane < nA + nB;
  const bool is_a = lane < nA;
  const int c = is_a ? lane : lane - nA;                       // chunk index inside the operand tile
  const bool mn = is_a ? (p.a_mn != 0) : (p.b_mn != 0);
  const uint32_t fe = is_a ? p.a_fe : p.b_fe, ke = is_a ? p.a_ke : p.b_ke;
  const CUtensorMap* map = is_a ? tmA : tmB;
  const uint32_t dst_off = (is_a ? 0u : (uint32_t)A_BYTES) + (mn ? (uint32_t)c * (BK * kRowBytes) : 0u);
  const uint32_t tiles_n = (uint32_t)p.tiles_n, tiles_m = (uint32_t)p.tiles_m;
  int s = 0; uint32_t ph = 0;
  for (int64_t tile64 = first; tile64 < p.num_tiles; tile64 += step) {
    uint32_t t = (uint32_t)tile64;
    const uint32_t ni = t % tiles_n; t /= tiles_n;
    const uint32_t mi = t % tiles_m;
    const int bi = (int)(t / tiles_m);
    int f;   // first free-mode (row) index of this lane's box
    if (is_a) f = (int)mi * (TWO ? 2 * kBM : kBM) + (TWO ? (int)rank * kBM : 0) + (mn ? c * CHUNK : 0);
    else f = (int)ni * p.BN + (TWO ? (int)rank * b_rows : 0) + (mn ? c * CHUNK : 0);
    const int f_in = fe ? (int)((uint32_t)f % fe) : f, f_out = fe ? (int)((uint32_t)f / fe) : 0;
    int k_in = 0, k_out = 0;
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait(bar_base + 8u * (S + s), ph ^ 1);               // slot free (every lane observes it)
      const uint32_t full = bar_base + 8u * s;
      if (lane == 0) {
        if (!TWO) mbar_expect_tx(full, (uint32_t)stage_bytes);
        else if (rank == 0) mbar_expect_tx(full, (uint32_t)(2 * stage_bytes));
        else mbar_arrive_remote(full, 0);
      }
      __syncwarp();
      if (mine) {
        const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes) + dst_off;
        if (!mn) { if (TWO) tma_load_5d_2sm(dst, map, full, k_in, f_in, f_out, k_out, bi); else tma_load_5d(dst, map, full, k_in, f_in, f_out, k_out, bi); }
        else     { if (TWO) tma_load_5d_2sm(dst, map, full, f_in, k_in, k_out, f_out, bi); else tma_load_5d(dst, map, full, f_in, k_in, k_out, f_out, bi); }
      }
      k_in += BK;
      if (ke && (uint32_t)k_in >= ke) { k_in = 0; ++k_out; }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ TcParams p) {
  constexpr int ES = KIND == 0 ? 2 : 4;          // operand element bytes
  constexpr int BK = kRowBytes / ES;             // 64 (16-bit) or 32 (tf32) elements per k-block
  constexpr int CHUNK = kRowBytes / ES;          // MN elements per 128-byte row of an MN-major tile
  constexpr int A_BYTES = kBM * kRowBytes;       // 16 KB per stage
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int BN = p.BN;
  const int B_BYTES = BN * kRowBytes;
  const int STAGE_BYTES = A_BYTES + B_BYTES;
  const int S = p.stages;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
  // bars: full[S], empty[S], tmem_full[2], tmem_empty[2]
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const uint32_t epi_slab = (smem_u32(tmem_slot) + 16 + 15) & ~15u;   // 4 warps x kSlabBytes of epilogue staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_kb;
  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp, see tma_producer)
    tma_producer<KIND, false>(&tmA, &tmB, p, smem, bar_base, S, STAGE_BYTES, BN, 0u, first, step, lane);
  } else if (warp == 1 && lane == 0) {
    // ===================================================== MMA issuer
    // descriptor fields per majorness (see DESIGN.md "UMMA operand layouts")
    const uint32_t a_layout = (p.a_mn && KIND == 1) ? 1u : 2u;
    const uint32_t b_layout = (p.b_mn && KIND == 1) ? 1u : 2u;
    const uint32_t a_lbo = p.a_mn ? BK * kRowBytes : 0u, b_lbo = p.b_mn ? BK * kRowBytes : 0u;
    const uint32_t a_sbo = (p.a_mn && KIND == 1) ? 512u : 1024u;
    const uint32_t b_sbo = (p.b_mn && KIND == 1) ? 512u : 1024u;
    // bytes the start address advances per MMA (K = 32 bytes of the contracted mode)
    const uint32_t a_kstep = p.a_mn ? (32u / ES) * kRowBytes : 32u;
    const uint32_t b_kstep = p.b_mn ? (32u / ES) * kRowBytes : 32u;
    int s = 0; uint32_t ph = 0;
    int acc = 0; uint32_t acc_ph = 0;
    for (int64_t tile = first; tile < p.num_tiles; tile += step) {
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_layout);
          const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_layout);
          tc_mma<KIND>(d_tmem, ad, bd, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(s));                 // frees the ring slot when these MMAs retire
        if (kb == num_kb - 1) tc_commit(tfull_bar(acc));
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int q = warp & 3;                      // TMEM lane quadrant of this warp
    int acc = 0; uint32_t acc_ph = 0;
    for (int64_t tile = first; tile < p.num_tiles; tile += step) {
      int64_t t = tile;
      const int ni = (int)(t % p.tiles_n); t /= p.tiles_n;
      const int mi = (int)(t % p.tiles_m);
      const int64_t bi = t / p.tiles_m;
      const int64_t row = (int64_t)mi * kBM + q * 32 + lane;
      const int64_t n0 = (int64_t)ni * BN;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      epilogue_drain(p, taddr0, BN, bi, row - lane, lane, n0, epi_slab + (uint32_t)q * kSlabBytes);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =====================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 columns); the MMA issued
// by the leader CTA (cluster rank 0) reads both halves of B across the pair, so the per-CTA
// L2->smem traffic per flop drops by 1.5x versus the 1-CTA 128 x 256 tile (the 1-CTA kernel
// is L2-bandwidth bound on K=512 problems: profiles/r1_ncu_full_tcgen05_bf16_flagship_batch64_v1.csv).
//   * TMA loads of both CTAs complete on the LEADER's full barrier (cta_group::2 + peer-bit mask);
//   * tcgen05.commit multicasts the slot release / accumulator-ready signal to both CTAs;
//   * both epilogues arrive (remotely for the peer) on the leader's tmem_empty barrier.
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ TcParams p) {
  constexpr int ES = KIND == 0 ? 2 : 4;
  constexpr int BK = kRowBytes / ES;
  constexpr int CHUNK = kRowBytes / ES;
  constexpr int A_BYTES = kBM * kRowBytes;       // this CTA's 128 rows of A
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int BN = p.BN;                           // pair tile is 256 x BN
  const int BNH = BN / 2;                        // this CTA's share of the B tile
  const int B_BYTES = BNH * kRowBytes;
  const int STAGE_BYTES = A_BYTES + B_BYTES;
  const int S = p.stages;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const uint32_t epi_slab = (smem_u32(tmem_slot) + 16 + 15) & ~15u;   // 4 warps x kSlabBytes of epilogue staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    // full: leader's expect_tx arrive + the peer's remote arrive; empty / tmem_full: one multicast commit;
    // tmem_empty: 4 epilogue warps of each CTA
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_kb;
  const int64_t first = blockIdx.x >> 1, step = gridDim.x >> 1;   // pair index

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs, whole warp)
    tma_producer<KIND, true>(&tmA, &tmB, p, smem, bar_base, S, STAGE_BYTES, BNH, rank, first, step, lane);
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================================================== MMA issuer (leader CTA only)
    const uint32_t a_layout = (p.a_mn && KIND == 1) ? 1u : 2u;
    const uint32_t b_layout = (p.b_mn && KIND == 1) ? 1u : 2u;
    const uint32_t a_lbo = p.a_mn ? BK * kRowBytes : 0u, b_lbo = p.b_mn ? BK * kRowBytes : 0u;
    const uint32_t a_sbo = (p.a_mn && KIND == 1) ? 512u : 1024u;
    const uint32_t b_sbo = (p.b_mn && KIND == 1) ? 512u : 1024u;
    const uint32_t a_kstep = p.a_mn ? (32u / ES) * kRowBytes : 32u;
    const uint32_t b_kstep = p.b_mn ? (32u / ES) * kRowBytes : 32u;
    int s = 0; uint32_t ph = 0;
    int acc = 0; uint32_t acc_ph = 0;
    for (int64_t tile = first; tile < p.num_tiles; tile += step) {
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_layout);
          const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_layout);
          tc_mma_2sm<KIND>(d_tmem, ad, bd, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit_2sm(empty_bar(s));
        if (kb == num_kb - 1) tc_commit_2sm(tfull_bar(acc));
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs: their own 128 rows)
    const int q = warp & 3;
    int acc = 0; uint32_t acc_ph = 0;
    for (int64_t tile = first; tile < p.num_tiles; tile += step) {
      int64_t t = tile;
      const int ni = (int)(t % p.tiles_n); t /= p.tiles_n;
      const int mi = (int)(t % p.tiles_m);
      const int64_t bi = t / p.tiles_m;
      const int64_t row = (int64_t)mi * 2 * kBM + (int64_t)rank * kBM + q * 32 + lane;
      const int64_t n0 = (int64_t)ni * BN;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      epilogue_drain(p, taddr0, BN, bi, row - lane, lane, n0, epi_slab + (uint32_t)q * kSlabBytes);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote(tempty_bar(acc), 0); }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int es_of(int dt) { return dt == TNB200_F32 ? 4 : 2; }

static bool tf32_mn_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TNB200_TF32_MN"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Which UMMA majorness can serve this view?  0: K-major, 1: MN-major, -1: neither (needs a repack).
// Conditions (es = element bytes, R = 128/es elements per swizzle row):
//   * unit stride on the innermost contracted mode (K-major) or innermost free mode (MN-major);
//   * every other stride and the base pointer 16-byte aligned; rank <= 5 (<= 2 modes per group);
//   * two contracted modes: inner extent % R == 0 (a k-block never straddles the mode boundary);
//   * two free modes: MN-major: inner extent % R == 0;  K-major: inner extent % 256 == 0, or a
//     power of two <= 64 (so every tile size 64/128/256 either divides it or is a multiple of it).
static int view_major(int dtype, const OperandView& v, int64_t ext_f, int64_t ext_k, int64_t batch) {
  if (dtype != TNB200_F32 && dtype != TNB200_F16 && dtype != TNB200_BF16) return -1;
  const int es = es_of(dtype);
  const int64_t R = kRowBytes / es;
  if (((uintptr_t)v.ptr) & 15) return -1;
  if (v.nF > 2 || v.nK > 2) return -1;
  if (ext_f >= (1LL << 31) || ext_k >= (1LL << 31) || batch >= (1LL << 31)) return -1;
  auto ok16 = [&](int64_t s) { return s > 0 && (s * es) % 16 == 0 && s * es < (1LL << 40); };
  if (batch > 1 && !ok16(v.sb)) return -1;
  if (v.nK == 2 && v.ke[1] % R != 0) return -1;
  const bool k_unit = ext_k == 1 || v.nK == 0 || v.ks[v.nK - 1] == 1;
  const bool f_unit = ext_f == 1 || v.nF == 0 || v.fs[v.nF - 1] == 1;
  if (k_unit) {  // K-major candidate: all free strides + outer K stride must be 16B multiples
    bool ok = true;
    for (int i = 0; i < v.nF; ++i) ok = ok && (v.fe[i] == 1 || ok16(v.fs[i]));
    if (v.nK == 2) ok = ok && ok16(v.ks[0]);
    if (v.nF == 2) { int64_t e = v.fe[1]; ok = ok && (e % 256 == 0 || (e <= 64 && (e & (e - 1)) == 0)); }
    if (ok) return 0;
  }
  if (f_unit) {
    bool ok = true;
    for (int i = 0; i < v.nK; ++i) ok = ok && (v.ke[i] == 1 || ok16(v.ks[i]));
    if (v.nF == 2) ok = ok && ok16(v.fs[0]) && (v.fe[1] % R == 0);
    if (dtype == TNB200_F32 && !tf32_mn_enabled()) ok = false;
    if (ok) return 1;
  }
  return -1;
}

bool tcgen05_view_ok(int dtype, const OperandView& v, int64_t ext_f, int64_t ext_k, int64_t batch) {
  return view_major(dtype, v, ext_f, ext_k, batch) >= 0;
}

// Build the rank-5 tensor map of one operand for tiles of `tile` free rows.
static int encode_operand(CUtensorMap* map, int dtype, const OperandView& v, int64_t ext_f, int64_t ext_k, int64_t batch,
                          int tile, bool& mn_major, uint32_t& fe_in, uint32_t& ke_in) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return TNB200_ERR_UNSUPPORTED;
  const int major = view_major(dtype, v, ext_f, ext_k, batch);
  if (major < 0) return TNB200_ERR_UNSUPPORTED;
  mn_major = major == 1;
  const int es = es_of(dtype);
  const int R = kRowBytes / es;
  CUtensorMapDataType cdt = dtype == TNB200_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                            : (dtype == TNB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  // (extent, stride) of inner / outer mode of each group; missing modes are extent 1
  int64_t f_in_e = v.nF ? v.fe[v.nF - 1] : 1, f_in_s = v.nF ? v.fs[v.nF - 1] : 0;
  int64_t f_out_e = v.nF == 2 ? v.fe[0] : 1, f_out_s = v.nF == 2 ? v.fs[0] : 0;
  int64_t k_in_e = v.nK ? v.ke[v.nK - 1] : 1, k_in_s = v.nK ? v.ks[v.nK - 1] : 0;
  int64_t k_out_e = v.nK == 2 ? v.ke[0] : 1, k_out_s = v.nK == 2 ? v.ks[0] : 0;
  fe_in = v.nF == 2 ? (uint32_t)f_in_e : 0u;
  ke_in = v.nK == 2 ? (uint32_t)k_in_e : 0u;
  int64_t de[5], ds[5];   // extents, element strides (ds[0] is the unit-stride dim)
  cuuint32_t box[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  if (!mn_major) {
    de[0] = k_in_e; ds[0] = 1;
    de[1] = f_in_e; ds[1] = f_in_s; de[2] = f_out_e; ds[2] = f_out_s;
    de[3] = k_out_e; ds[3] = k_out_s; de[4] = batch; ds[4] = v.sb;
    box[0] = R;
    if (v.nF == 2 && f_in_e < tile) {
      if (tile % f_in_e) return TNB200_ERR_UNSUPPORTED;
      box[1] = (cuuint32_t)f_in_e; box[2] = (cuuint32_t)(tile / f_in_e);
    } else {
      if (v.nF == 2 && f_in_e % tile) return TNB200_ERR_UNSUPPORTED;
      box[1] = tile;
    }
  } else {
    de[0] = f_in_e; ds[0] = 1;
    de[1] = k_in_e; ds[1] = k_in_s; de[2] = k_out_e; ds[2] = k_out_s;
    de[3] = f_out_e; ds[3] = f_out_s; de[4] = batch; ds[4] = v.sb;
    box[0] = R; box[1] = R;   // R elements of the free mode (128 B) x BK = R contracted rows
  }
  cuuint64_t dims[5], strides[4];
  int64_t natural = 16;   // bytes: a packed stride for extent-1 (never addressed) dims
  for (int d = 0; d < 5; ++d) {
    dims[d] = (cuuint64_t)(de[d] < 1 ? 1 : de[d]);
    int64_t bytes = es;
    if (d > 0) {
      bytes = ds[d] * es;
      if (de[d] <= 1) bytes = (natural + 15) / 16 * 16;
      else if (bytes <= 0 || bytes % 16) return TNB200_ERR_UNSUPPORTED;
      strides[d - 1] = (cuuint64_t)bytes;
    }
    if ((int64_t)dims[d] * bytes > natural) natural = (int64_t)dims[d] * bytes;
  }
  CUtensorMapSwizzle sw = (mn_major && dtype == TNB200_F32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, cdt, 5, const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return TNB200_ERR_UNSUPPORTED;
  return 0;
}

struct TcPrep {
  TcParams p;
  CUtensorMap tmA, tmB;
  bool use2;
  int stage_bytes, b_rows;
};

// Tile selection, tensor maps and instruction descriptor of one GEMM.  chain_mode: the problem is one step of a
// chained launch (gemm_tcgen05_chain_kernel): always 2-CTA pair tiles with the widest BN the problem fills.
static int tc_prepare(const GemmProblem& g, bool chain_mode, TcPrep& o) {
  TcParams& p = o.p;
  const int es = es_of(g.dtype);
  const int sms = num_sms();
  const int64_t tiles_m = (g.M + kBM - 1) / kBM;
  int BN = 64;   // multiples of 64 so that an MN-major B tile is a whole number of 128-byte chunks
  if (g.swapped) {
    // tiny N: 32 columns suffice unless the (MN-major) B tile needs whole 128-byte chunks
    const bool b_unit_f = g.B.nF == 0 || g.B.fs[g.B.nF - 1] == 1;
    const bool b_unit_k = g.K == 1 || g.B.nK == 0 || g.B.ks[g.B.nK - 1] == 1;
    BN = (g.N <= 32 && (b_unit_k || (b_unit_f && es == 4))) ? 32 : 64;
  } else {
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      int bn = cands[i];
      if (bn > 64 && bn / 2 >= g.N) continue;          // tile mostly empty
      int64_t tiles = tiles_m * ((g.N + bn - 1) / bn) * g.batch;
      if (tiles >= sms || bn == 64) { BN = bn; break; }
    }
    if (chain_mode) BN = g.N >= 256 ? 256 : 128;
  }
  // 2-CTA pairs (256 x BN tiles) for problems large enough to fill the GPU with pair tiles
  static int force2 = -2;
  if (force2 == -2) { const char* e = getenv("TNB200_2CTA"); force2 = e ? (e[0] == '0' ? 0 : 1) : -1; }
  bool use2 = false;
  if (g.M >= 256 && BN >= 128 && !g.swapped) {
    const int64_t pair_tiles = ((g.M + 2 * kBM - 1) / (2 * kBM)) * ((g.N + BN - 1) / BN) * g.batch;
    use2 = force2 == 1 || (force2 == -1 && BN == 256 && pair_tiles >= sms / 2);
  }
  if (force2 == 0) use2 = false;
  if (chain_mode) {
    if (g.swapped || g.M < 256 || g.N < 128) return TNB200_ERR_UNSUPPORTED;
    use2 = true;
  }
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
  p.BN = BN;
  const int bk = kRowBytes / es;
  p.num_kb = (int)((g.K + bk - 1) / bk);
  p.tiles_m = use2 ? (g.M + 2 * kBM - 1) / (2 * kBM) : tiles_m;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.num_tiles = p.tiles_m * p.tiles_n * g.batch;
  const int b_rows = use2 ? BN / 2 : BN;                   // B rows staged per CTA per k-block
  const int stage_bytes = kBM * kRowBytes + b_rows * kRowBytes;
  int stages = (196 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > p.num_kb + 1 && p.num_tiles <= sms) stages = p.num_kb + 1 > 2 ? p.num_kb + 1 : 2;
  p.stages = stages;
  p.out_kind = g.dtype == TNB200_BF16 ? 0 : (g.dtype == TNB200_F16 ? 1 : 2);
  p.C = g.C; p.c_sm = g.c_sm; p.c_sb = g.c_sb; p.c_cs = g.swapped ? g.c_sn : 1;
  if (g.swapped && p.c_cs == 1) p.c_cs = 2;   // degenerate (M == 1): force the transposed-store path; stride unused
  p.c_hint = 0;
  p.vec_ok = !g.swapped && (((uintptr_t)g.C) % 16 == 0) && ((g.c_sm * es) % 16 == 0) && ((g.c_sb * es) % 16 == 0);
    bool a_mn = false, b_mn = false;
  int rc = encode_operand(&o.tmA, g.dtype, g.A, g.M, g.K, g.batch, kBM, a_mn, p.a_fe, p.a_ke);
  if (rc) return rc;
  rc = encode_operand(&o.tmB, g.dtype, g.B, g.N, g.K, g.batch, b_rows, b_mn, p.b_fe, p.b_ke);
  if (rc) return rc;
  p.a_mn = a_mn; p.b_mn = b_mn;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B format, majors, N>>3, M>>4
  const uint32_t fmt = g.dtype == TNB200_BF16 ? 1u : (g.dtype == TNB200_F16 ? 0u : 2u);
  const uint32_t mma_m = use2 ? 256u : 128u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((mma_m >> 4) << 24);
  o.use2 = use2; o.stage_bytes = stage_bytes; o.b_rows = b_rows;
  return 0;
}

int gemm_tcgen05(const GemmProblem& g, cudaStream_t st) {
  if (g.dtype != TNB200_F32 && g.dtype != TNB200_F16 && g.dtype != TNB200_BF16) return TNB200_ERR_UNSUPPORTED;
  {
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("TNB200_NO_TCGEN05"); disabled = (e && e[0] == '1') ? 1 : 0; }
    if (disabled) return TNB200_ERR_UNSUPPORTED;
  }
  if (g.conjA || g.conjB) return TNB200_ERR_UNSUPPORTED;
  if (!g.swapped && g.c_sn != 1 && g.N > 1) return TNB200_ERR_UNSUPPORTED;
  if (g.swapped && g.c_sm != 1 && g.M > 1) return TNB200_ERR_UNSUPPORTED;
  if (g.M >= (1LL << 31) || g.N >= (1LL << 31) || g.K >= (1LL << 31)) return TNB200_ERR_UNSUPPORTED;
  if (!tcgen05_view_ok(g.dtype, g.A, g.M, g.K, g.batch) || !tcgen05_view_ok(g.dtype, g.B, g.N, g.K, g.batch))
    return TNB200_ERR_UNSUPPORTED;
  // swap-AB: a tiny M under a large N would waste the 128-row MMA; compute C^T = B^T A^T instead
  // (the big free dimension rides the 128 tile rows, the tiny one a 32/64-column tile) and let the
  // epilogue store the tile transposed — lanes then write consecutive addresses.
  if (g.M <= 64 && g.N >= 128 && !g.swapped) {
    GemmProblem t = g;
    t.swapped = true;
    t.M = g.N; t.N = g.M; t.A = g.B; t.B = g.A;
    t.c_sm = g.c_sn; t.c_sn = g.c_sm;          // row stride of C^T = column stride of C (1)
    return gemm_tcgen05(t, st);
  }
  TcPrep prep;
  {
    int rc = tc_prepare(g, false, prep);
    if (rc) return rc;
  }
  const TcParams& p = prep.p;
  const CUtensorMap& tmA = prep.tmA;
  const CUtensorMap& tmB = prep.tmB;
  const bool use2 = prep.use2;
  const int stages = p.stages, stage_bytes = prep.stage_bytes;
  const int sms = num_sms();
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 4) * 8 + 32 + 4 * kSlabBytes + 1024;
  const int kind = g.dtype == TNB200_F32 ? 1 : 0;
  static bool attr_set[4] = {false, false, false, false};
  const int ai = kind + (use2 ? 2 : 0);
  if (!attr_set[ai]) {
    cudaError_t e;
    if (!use2) e = kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                             : cudaFuncSetAttribute(gemm_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    else e = kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                       : cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("tcgen05: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return TNB200_ERR_CUDA; }
    attr_set[ai] = true;
  }
  if (!use2) {
    const int64_t grid = p.num_tiles < sms ? p.num_tiles : sms;
    if (kind == 0) gemm_tcgen05_kernel<0><<<(unsigned)grid, kThreads, smem, st>>>(tmA, tmB, p);
    else gemm_tcgen05_kernel<1><<<(unsigned)grid, kThreads, smem, st>>>(tmA, tmB, p);
  } else {
    const int64_t pairs = p.num_tiles < sms / 2 ? p.num_tiles : sms / 2;
    if (kind == 0) gemm_tcgen05_2cta_kernel<0><<<(unsigned)(2 * pairs), kThreads, smem, st>>>(tmA, tmB, p);
    else gemm_tcgen05_2cta_kernel<1><<<(unsigned)(2 * pairs), kThreads, smem, st>>>(tmA, tmB, p);
  }
  TNB_LAUNCH_CHECK();
  count_launch();
  if (use2) set_kernel_name(g.dtype == TNB200_BF16 ? "tcgen05_2cta_bf16" : (g.dtype == TNB200_F16 ? "tcgen05_2cta_f16" : "tcgen05_2cta_tf32"));
  else set_kernel_name(g.dtype == TNB200_BF16 ? "tcgen05_bf16" : (g.dtype == TNB200_F16 ? "tcgen05_f16" : "tcgen05_tf32"));
  return 0;
}

int gemm_chain_destroy(void* handle);

// =====================================================================================
// Chained GEMMs in ONE persistent launch (cta_group::2 pair tiles only).
//
// A contraction path often contains long runs of dependent GEMMs (the MPS "zipper": E' = A^T (E A) per site).
// Launched one by one, every step pays the persistent kernel's prologue + drain, and every intermediate makes
// a round trip through HBM (the cfg-2 bulk step is HBM-bound: 184 flop/B).  Here all steps of such a run are
// tiles of ONE kernel: the tile sequence is ordered so that a small group of samples is carried through the
// whole run of steps before the next group starts (intermediates are produced and consumed while still in
// L2), and inter-step dependencies are tracked per (step, sample) with release/acquire counters in global
// memory: an epilogue warp publishes its part of an output tile with red.release, the TMA producer of a
// dependent tile spins with ld.acquire + fence.proxy.async before its first load.  Tiles are assigned to
// CTA pairs round-robin in sequence order and all CTAs are co-resident (grid <= SM count), so a tile only
// ever waits for tiles that are earlier in the sequence: no deadlock.
struct alignas(64) ChainStepDev {
  CUtensorMap tmA, tmB;
  CUtensorMap tmAh;               // A with a 64-row box (K-major operands): one half of a CTA's slab, for the 4-CTA multicast kernel
  TcParams p;
  int dep_a, dep_b;               // chain step that produces operand a / b (-1: available before the launch)
  uint32_t need_a, need_b;        // counter value of that step's (sample) entry when it is complete
  int tiles_per_sample;
  uint32_t tx_bytes;              // bytes landing on the leader's full barrier per k-block (both CTAs)
  int hint_a, hint_b;             // L2 policy of the operand loads: 0 default, 1 evict_first (streamed external operand), 2 evict_last (intermediate)
};
struct ChainSeg { long long tile0; int step, sample0, nsamples, pad; };
struct ChainParams {
  const ChainStepDev* steps;
  const ChainSeg* segs;
  uint32_t* done;                 // [nsteps][batch] completion counters (zeroed before every launch)
  long long num_tiles;
  int nsegs, batch, stages, stage_bytes;
  int debug_no_mma;               // measurement aid (TNB200_CHAIN_NOMMA=1): stream operands, skip the MMAs
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin on a dependency counter (a scheduling bug traps instead of hanging the box)
__device__ __forceinline__ void chain_wait(const uint32_t* ctr, uint32_t need) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    if (ld_acquire_u32(ctr) >= need) return;
    __nanosleep(64);
  }
  __trap();
}

struct ChainTile { int step, bi, mi, ni; };
__device__ __forceinline__ const ChainStepDev* chain_decode(const ChainParams& cp, long long tile, int& cursor, ChainTile& t) {
  while (cursor + 1 < cp.nsegs && tile >= cp.segs[cursor + 1].tile0) ++cursor;
  const ChainSeg sg = cp.segs[cursor];
  const ChainStepDev* sd = cp.steps + sg.step;
  uint32_t local = (uint32_t)(tile - sg.tile0);
  const uint32_t tn = (uint32_t)sd->p.tiles_n, tm = (uint32_t)sd->p.tiles_m;
  t.ni = (int)(local % tn); local /= tn;
  t.mi = (int)(local % tm);
  t.bi = sg.sample0 + (int)(local / tm);
  t.step = sg.step;
  return sd;
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_chain_kernel(const __grid_constant__ ChainParams cp) {
  constexpr int ES = KIND == 0 ? 2 : 4;
  constexpr int BK = kRowBytes / ES;
  constexpr int CHUNK = kRowBytes / ES;
  constexpr int A_BYTES = kBM * kRowBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STAGE_BYTES = cp.stage_bytes;
  const int S = cp.stages;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const uint32_t epi_slab = (smem_u32(tmem_slot) + 16 + 15) & ~15u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long first = blockIdx.x >> 1, step = gridDim.x >> 1;   // pair index / number of pairs

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs, whole warp)
    int s = 0; uint32_t ph = 0;
    int cursor = 0;
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const int BN = sd->p.BN, b_rows = BN / 2;
      const int a_mn = sd->p.a_mn, b_mn = sd->p.b_mn;
      const int nA = a_mn ? kBM / CHUNK : 1;
      const int nB = b_mn ? b_rows / CHUNK : 1;
      const bool mine = lane < nA + nB;
      const bool is_a = lane < nA;
      const int c = is_a ? lane : lane - nA;
      const bool mn = is_a ? (a_mn != 0) : (b_mn != 0);
      const uint32_t fe = is_a ? sd->p.a_fe : sd->p.b_fe, ke = is_a ? sd->p.a_ke : sd->p.b_ke;
      const CUtensorMap* map = is_a ? &sd->tmA : &sd->tmB;
      const uint32_t dst_off = (is_a ? 0u : (uint32_t)A_BYTES) + (mn ? (uint32_t)c * (BK * kRowBytes) : 0u);
      int f;
      if (is_a) f = t.mi * 2 * kBM + (int)rank * kBM + (mn ? c * CHUNK : 0);
      else f = t.ni * BN + (int)rank * b_rows + (mn ? c * CHUNK : 0);
      const int f_in = fe ? (int)((uint32_t)f % fe) : f, f_out = fe ? (int)((uint32_t)f / fe) : 0;
      const int num_kb = sd->p.num_kb;
      const uint32_t tx = sd->tx_bytes;
      const int hint = is_a ? sd->hint_a : sd->hint_b;
      const uint64_t policy = hint == 2 ? l2_policy_evict_last() : l2_policy_evict_first();
      // operands produced by earlier steps of this launch: wait until every tile of (that step, this sample) is out
      if (lane == 0) {
        if (sd->dep_a >= 0) chain_wait(cp.done + (size_t)sd->dep_a * cp.batch + t.bi, sd->need_a);
        if (sd->dep_b >= 0) chain_wait(cp.done + (size_t)sd->dep_b * cp.batch + t.bi, sd->need_b);
      }
      __syncwarp();
      asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy writes (other SMs' epilogues) -> our TMA reads
      int k_in = 0, k_out = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t full = full_bar(s);
        if (lane == 0) {
          if (leader) mbar_expect_tx(full, tx);
          else mbar_arrive_remote(full, 0);
        }
        __syncwarp();
        if (mine) {
          const uint32_t dst = smem_u32(smem + (size_t)s * STAGE_BYTES) + dst_off;
          if (hint) {
            if (!mn) tma_load_5d_2sm_hint(dst, map, full, k_in, f_in, f_out, k_out, t.bi, policy);
            else     tma_load_5d_2sm_hint(dst, map, full, f_in, k_in, k_out, f_out, t.bi, policy);
          } else {
            if (!mn) tma_load_5d_2sm(dst, map, full, k_in, f_in, f_out, k_out, t.bi);
            else     tma_load_5d_2sm(dst, map, full, f_in, k_in, k_out, f_out, t.bi);
          }
        }
        k_in += BK;
        if (ke && (uint32_t)k_in >= ke) { k_in = 0; ++k_out; }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================================================== MMA issuer (leader CTA only)
    int s = 0; uint32_t ph = 0;
    int acc = 0; uint32_t acc_ph = 0;
    int cursor = 0;
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const int a_mn = sd->p.a_mn, b_mn = sd->p.b_mn, num_kb = sd->p.num_kb;
      const uint32_t idesc = sd->p.idesc;
      const uint32_t a_layout = (a_mn && KIND == 1) ? 1u : 2u;
      const uint32_t b_layout = (b_mn && KIND == 1) ? 1u : 2u;
      const uint32_t a_lbo = a_mn ? BK * kRowBytes : 0u, b_lbo = b_mn ? BK * kRowBytes : 0u;
      const uint32_t a_sbo = (a_mn && KIND == 1) ? 512u : 1024u;
      const uint32_t b_sbo = (b_mn && KIND == 1) ? 512u : 1024u;
      const uint32_t a_kstep = a_mn ? (32u / ES) * kRowBytes : 32u;
      const uint32_t b_kstep = b_mn ? (32u / ES) * kRowBytes : 32u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_layout);
          const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_layout);
          if (!cp.debug_no_mma) tc_mma_2sm<KIND>(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit_2sm(empty_bar(s));
        if (kb == num_kb - 1) tc_commit_2sm(tfull_bar(acc));
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs: their own 128 rows)
    // Publishing a tile part needs a device-scope release AFTER its stores are performed: done right after the
    // drain it would stall the warp for the full store round trip (ncu: MEMBAR + CCTL.IVALL were the top stalls of
    // the epilogue warps).  Instead the publication of tile i is deferred until tile i+1's accumulator is ready —
    // by then the stores have long completed and the fence is free — or done immediately when the warp would
    // otherwise idle (a pending publication must never wait behind a dependency: that could deadlock).
    const int q = warp & 3;
    int acc = 0; uint32_t acc_ph = 0;
    int cursor = 0;
    uint32_t* pending = nullptr;
    auto publish = [&]() {
      if (pending != nullptr) {
        __threadfence();
        __syncwarp();
        if (lane == 0) red_release_add_u32(pending, 1u);
        pending = nullptr;
      }
    };
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const TcParams p = sd->p;                              // per-step output description (registers / local)
      const int64_t row = (int64_t)t.mi * 2 * kBM + (int64_t)rank * kBM + q * 32 + lane;
      const int64_t n0 = (int64_t)t.ni * p.BN;
      if (!mbar_test(tfull_bar(acc), acc_ph)) publish();     // idle anyway: publish now
      mbar_wait(tfull_bar(acc), acc_ph);
      publish();                                             // previous tile's stores are long done: cheap
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      epilogue_drain(p, taddr0, p.BN, t.bi, row - lane, lane, n0, epi_slab + (uint32_t)q * kSlabBytes);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote(tempty_bar(acc), 0); }
      pending = cp.done + (size_t)t.step * cp.batch + t.bi;
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
    publish();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- 4-CTA clusters: two cta_group::2 pairs side by side in N share their A slabs by TMA multicast.
// The 2-CTA chain kernel is bound by the L2 -> SM operand feed (128 flop per L2 byte: 16 KB of A + 16 KB of B per CTA and
// k-block).  Here a cluster holds pair 0 = ranks {0, 1} on tile (mi, 2j) and pair 1 = ranks {2, 3} on tile (mi, 2j + 1):
// both pairs need the same 256 x BK block of A.  Rank r loads only HALF of its 128-row slab (half = r >> 1) and multicasts
// it to the CTA of the other pair that stages the same slab (mask {r & 1, (r & 1) + 2}); every CTA still receives its
// whole slab, but issues 8 KB + 16 KB instead of 16 KB + 16 KB per k-block: 171 flop per L2 byte.
//   * full barriers: unchanged accounting — each pair leader expects 2 x stage bytes; multicast bytes are signalled on the
//     leader of every destination CTA (cta_group::2, peer bit cleared);
//   * empty barriers: a stage is written by BOTH pairs' producers, so it is released by BOTH pairs' MMA commits
//     (count 2, tcgen05.commit multicast to all four CTAs): the two pairs advance k-block by k-block together;
//   * accumulator barriers stay pair-local (commit mask 0b0011 / 0b1100).
// Tiles are dealt to clusters two at a time (2 st, 2 st + 1): needs an even number of N tiles in every step.
template <int KIND>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_chain4_kernel(const __grid_constant__ ChainParams cp) {
  constexpr int ES = KIND == 0 ? 2 : 4;
  constexpr int BK = kRowBytes / ES;
  constexpr int CHUNK = kRowBytes / ES;
  constexpr int A_BYTES = kBM * kRowBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STAGE_BYTES = cp.stage_bytes;
  const int S = cp.stages;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const uint32_t epi_slab = (smem_u32(tmem_slot) + 16 + 15) & ~15u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0..3
  const uint32_t pr = rank & 1;                     // rank inside the cta_group::2 pair
  const uint32_t pair_in = rank >> 1;               // which pair of the cluster = which half of the A slab this CTA loads
  const uint32_t pair_leader = rank & ~1u;
  const bool leader = pr == 0;
  const uint16_t pair_mask = (uint16_t)(3u << pair_leader);
  const uint16_t a_mask = (uint16_t)((1u << pr) | (1u << (pr + 2)));

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 2); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long first = 2 * (long long)(blockIdx.x >> 2) + pair_in, step = 2 * (long long)(gridDim.x >> 2);

  if (warp == 0) {
    // ===================================================== TMA producer (all four CTAs, whole warp)
    int s = 0; uint32_t ph = 0;
    int cursor = 0;
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const int BN = sd->p.BN, b_rows = BN / 2;
      const int a_mn = sd->p.a_mn, b_mn = sd->p.b_mn;
      const int nA = a_mn ? (kBM / CHUNK) / 2 : 1;            // boxes of MY half of the slab
      const int nB = b_mn ? b_rows / CHUNK : 1;
      const bool mine = lane < nA + nB;
      const bool is_a = lane < nA;
      const bool mn = is_a ? (a_mn != 0) : (b_mn != 0);
      const uint32_t fe = is_a ? sd->p.a_fe : sd->p.b_fe, ke = is_a ? sd->p.a_ke : sd->p.b_ke;
      const CUtensorMap* map = is_a ? (a_mn ? &sd->tmA : &sd->tmAh) : &sd->tmB;
      int f;
      uint32_t dst_off;
      if (is_a) {
        if (mn) { const int c = (int)pair_in * nA + lane; f = t.mi * 2 * kBM + (int)pr * kBM + c * CHUNK; dst_off = (uint32_t)c * (BK * kRowBytes); }
        else { f = t.mi * 2 * kBM + (int)pr * kBM + (int)pair_in * (kBM / 2); dst_off = pair_in * (uint32_t)(A_BYTES / 2); }
      } else {
        const int c = lane - nA;
        f = t.ni * BN + (int)pr * b_rows + (mn ? c * CHUNK : 0);
        dst_off = (uint32_t)A_BYTES + (mn ? (uint32_t)c * (BK * kRowBytes) : 0u);
      }
      const int f_in = fe ? (int)((uint32_t)f % fe) : f, f_out = fe ? (int)((uint32_t)f / fe) : 0;
      const int num_kb = sd->p.num_kb;
      const uint32_t tx = sd->tx_bytes;
      if (lane == 0) {
        if (sd->dep_a >= 0) chain_wait(cp.done + (size_t)sd->dep_a * cp.batch + t.bi, sd->need_a);
        if (sd->dep_b >= 0) chain_wait(cp.done + (size_t)sd->dep_b * cp.batch + t.bi, sd->need_b);
      }
      __syncwarp();
      asm volatile("fence.proxy.async;" ::: "memory");
      int k_in = 0, k_out = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);                      // both pairs have consumed this stage (here AND in the twin CTA)
        const uint32_t full = full_bar(s);
        if (lane == 0) {
          if (leader) mbar_expect_tx(full, tx);
          else mbar_arrive_remote(full, pair_leader);
        }
        __syncwarp();
        if (mine) {
          const uint32_t dst = smem_u32(smem + (size_t)s * STAGE_BYTES) + dst_off;
          if (is_a) {
            if (!mn) tma_load_5d_2sm_mc(dst, map, full, k_in, f_in, f_out, k_out, t.bi, a_mask);
            else     tma_load_5d_2sm_mc(dst, map, full, f_in, k_in, k_out, f_out, t.bi, a_mask);
          } else {
            if (!mn) tma_load_5d_2sm(dst, map, full, k_in, f_in, f_out, k_out, t.bi);
            else     tma_load_5d_2sm(dst, map, full, f_in, k_in, k_out, f_out, t.bi);
          }
        }
        k_in += BK;
        if (ke && (uint32_t)k_in >= ke) { k_in = 0; ++k_out; }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================================================== MMA issuer (the leader of each pair)
    int s = 0; uint32_t ph = 0;
    int acc = 0; uint32_t acc_ph = 0;
    int cursor = 0;
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const int a_mn = sd->p.a_mn, b_mn = sd->p.b_mn, num_kb = sd->p.num_kb;
      const uint32_t idesc = sd->p.idesc;
      const uint32_t a_layout = (a_mn && KIND == 1) ? 1u : 2u;
      const uint32_t b_layout = (b_mn && KIND == 1) ? 1u : 2u;
      const uint32_t a_lbo = a_mn ? BK * kRowBytes : 0u, b_lbo = b_mn ? BK * kRowBytes : 0u;
      const uint32_t a_sbo = (a_mn && KIND == 1) ? 512u : 1024u;
      const uint32_t b_sbo = (b_mn && KIND == 1) ? 512u : 1024u;
      const uint32_t a_kstep = a_mn ? (32u / ES) * kRowBytes : 32u;
      const uint32_t b_kstep = b_mn ? (32u / ES) * kRowBytes : 32u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_layout);
          const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_layout);
          if (!cp.debug_no_mma) tc_mma_2sm<KIND>(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit_2sm_mask(empty_bar(s), (uint16_t)0xF);      // this pair is done with stage s in all four CTAs' view
        if (kb == num_kb - 1) tc_commit_2sm_mask(tfull_bar(acc), pair_mask);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (all CTAs: their own 128 rows)
    const int q = warp & 3;
    int acc = 0; uint32_t acc_ph = 0;
    int cursor = 0;
    uint32_t* pending = nullptr;
    auto publish = [&]() {
      if (pending != nullptr) {
        __threadfence();
        __syncwarp();
        if (lane == 0) red_release_add_u32(pending, 1u);
        pending = nullptr;
      }
    };
    for (long long tile = first; tile < cp.num_tiles; tile += step) {
      ChainTile t;
      const ChainStepDev* sd = chain_decode(cp, tile, cursor, t);
      const TcParams p = sd->p;
      const int64_t row = (int64_t)t.mi * 2 * kBM + (int64_t)pr * kBM + q * 32 + lane;
      const int64_t n0 = (int64_t)t.ni * p.BN;
      if (!mbar_test(tfull_bar(acc), acc_ph)) publish();
      mbar_wait(tfull_bar(acc), acc_ph);
      publish();
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      epilogue_drain(p, taddr0, p.BN, t.bi, row - lane, lane, n0, epi_slab + (uint32_t)q * kSlabBytes);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote(tempty_bar(acc), pair_leader); }
      pending = cp.done + (size_t)t.step * cp.batch + t.bi;
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
    publish();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------- chain: host side
struct ChainHandle {
  ChainStepDev* d_steps = nullptr;
  ChainSeg* d_segs = nullptr;
  uint32_t* d_done = nullptr;
  ChainParams cp;
  int kind = 0, nsteps = 0;
  int clusters = 0;
  int cl = 2;                     // CTAs per cluster: 2 (one pair) or 4 (two pairs sharing A by multicast)
  size_t smem = 0, done_bytes = 0;
  unsigned grid = 0;
  double flops = 0.0;
};

int gemm_chain_create(int nsteps, const GemmProblem* probs, const int* dep_a, const int* dep_b, void** handle) {
  *handle = nullptr;
  if (nsteps < 1) return TNB200_ERR_INVALID;
  {
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("TNB200_NO_CHAIN"); disabled = (e && e[0] == '1') ? 1 : 0; }
    if (disabled) return TNB200_ERR_UNSUPPORTED;
  }
  const int dtype = probs[0].dtype;
  const int64_t batch = probs[0].batch;
  std::vector<ChainStepDev> steps(nsteps);
  int max_stage = 0;
  double flops = 0.0;
  bool cl4_ok = true;
  { const char* e = getenv("TNB200_CHAIN_CL"); if (!(e && e[0] == '4')) cl4_ok = false; }   // opt-in until validated: TNB200_CHAIN_CL=4
  for (int i = 0; i < nsteps; ++i) {
    const GemmProblem& g = probs[i];
    if (g.dtype != dtype || g.batch != batch || g.conjA || g.conjB) return TNB200_ERR_UNSUPPORTED;
    if (g.dtype != TNB200_F32 && g.dtype != TNB200_F16 && g.dtype != TNB200_BF16) return TNB200_ERR_UNSUPPORTED;
    if (g.c_sn != 1 || g.M >= (1LL << 31) || g.N >= (1LL << 31) || g.K >= (1LL << 31)) return TNB200_ERR_UNSUPPORTED;
    if (dep_a[i] >= i || dep_b[i] >= i) return TNB200_ERR_INVALID;
    TcPrep prep;
    int rc = tc_prepare(g, true, prep);
    if (rc) return rc;
    ChainStepDev& sd = steps[i];
    sd.tmA = prep.tmA; sd.tmB = prep.tmB; sd.p = prep.p;
    sd.tmAh = prep.tmA;
    if (!prep.p.a_mn) {           // K-major A: a second map whose box is half a slab (64 rows) for the multicast kernel
      bool mn2 = false; uint32_t fe2 = 0, ke2 = 0;
      if (encode_operand(&sd.tmAh, g.dtype, g.A, g.M, g.K, g.batch, kBM / 2, mn2, fe2, ke2) != 0 || mn2 || fe2 != prep.p.a_fe || ke2 != prep.p.a_ke)
        cl4_ok = false;
    }
    if (prep.p.tiles_n % 2) cl4_ok = false;
    sd.dep_a = dep_a[i]; sd.dep_b = dep_b[i];
    sd.tiles_per_sample = (int)(prep.p.tiles_m * prep.p.tiles_n);
    sd.tx_bytes = (uint32_t)(2 * prep.stage_bytes);
    sd.need_a = sd.need_b = 0;
    if (prep.stage_bytes > max_stage) max_stage = prep.stage_bytes;
    flops += 2.0 * (double)g.M * (double)g.N * (double)g.K * (double)batch;
  }
  // L2 residency policy, TNB200_CHAIN_L2 = bit mask (measurement knob; default 0 = hardware default policy):
  //   1: operands produced by an earlier step of the run are loaded evict_last
  //   2: operands that only stream through (site tensors) are loaded evict_first
  //   4: results consumed by a later step are stored evict_last
  // Measured on cfg 2 (bf16, 74 networks): all three together cost 7 % (3.76 ms vs 3.52 ms per launch) — see DESIGN.md.
  {
    const char* e = getenv("TNB200_CHAIN_L2");
    const int mask = e ? atoi(e) : 0;
    std::vector<char> consumed(nsteps, 0);
    for (int i = 0; i < nsteps; ++i) {
      if (dep_a[i] >= 0) consumed[dep_a[i]] = 1;
      if (dep_b[i] >= 0) consumed[dep_b[i]] = 1;
    }
    for (int i = 0; i < nsteps; ++i) {
      steps[i].hint_a = dep_a[i] >= 0 ? ((mask & 1) ? 2 : 0) : ((mask & 2) ? 1 : 0);
      steps[i].hint_b = dep_b[i] >= 0 ? ((mask & 1) ? 2 : 0) : ((mask & 2) ? 1 : 0);
      steps[i].p.c_hint = ((mask & 4) && consumed[i]) ? 1 : 0;
    }
  }
  for (int i = 0; i < nsteps; ++i) {
    if (steps[i].dep_a >= 0) steps[i].need_a = 8u * (uint32_t)steps[steps[i].dep_a].tiles_per_sample;   // 2 CTAs x 4 epilogue warps per tile
    if (steps[i].dep_b >= 0) steps[i].need_b = 8u * (uint32_t)steps[steps[i].dep_b].tiles_per_sample;
  }
  // ---- tile sequence.  The batch is cut into rounds of G samples and every round is carried through ALL steps
  // before the next one starts, so a step's results are consumed soon after they are produced; a round must be
  // wide enough that a dependent tile is >= 2 full waves of tiles behind its producers (TMA runs ~1 tile ahead
  // of the MMA, the epilogue ~1 tile behind: measured on cfg 2, G = 9 -> 8.4 ms, G = 37 -> 3.56 ms,
  // G = 74 (one round) -> 3.79 ms).  `rot` > 1 interleaves rot sub-groups inside a round (kept for experiments).
  const int sms = num_sms();
  const int pairs = sms / 2;
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  int min_tps = steps[0].tiles_per_sample;
  for (int i = 1; i < nsteps; ++i) if (steps[i].tiles_per_sample < min_tps) min_tps = steps[i].tiles_per_sample;
  if (batch * min_tps < pairs && !env_int("TNB200_CHAIN_FORCE", 0))
    return TNB200_ERR_UNSUPPORTED;      // too few tiles per step to hide the producer->consumer latency: launch step by step
  int G = env_int("TNB200_CHAIN_G", (2 * pairs + min_tps - 1) / min_tps);
  if (G < 1) G = 1;
  int rot = env_int("TNB200_CHAIN_ROT", 1);
  if (rot < 1) rot = 1;
  std::vector<ChainSeg> segs;
  long long tile0 = 0;
  for (int64_t r0 = 0; r0 < batch; r0 += (int64_t)G * rot) {
    const int64_t r1 = r0 + (int64_t)G * rot < batch ? r0 + (int64_t)G * rot : batch;
    for (int i = 0; i < nsteps; ++i)
      for (int sub = 0; sub < rot; ++sub) {
        const int64_t s0 = r0 + (int64_t)sub * G, s1 = s0 + G < r1 ? s0 + G : r1;
        if (s0 >= s1) continue;
        ChainSeg sg;
        sg.tile0 = tile0; sg.step = i; sg.sample0 = (int)s0; sg.nsamples = (int)(s1 - s0); sg.pad = 0;
        segs.push_back(sg);
        tile0 += (long long)(s1 - s0) * steps[i].tiles_per_sample;
      }
  }
  int stages = (196 * 1024) / max_stage;
  if (stages > 8) stages = 8;
  if (stages < 2) return TNB200_ERR_UNSUPPORTED;
  ChainHandle* h = new ChainHandle();
  h->kind = dtype == TNB200_F32 ? 1 : 0;
  h->nsteps = nsteps;
  h->flops = flops;
  h->done_bytes = sizeof(uint32_t) * (size_t)nsteps * (size_t)batch;
  cudaError_t e = cudaMalloc(&h->d_steps, sizeof(ChainStepDev) * steps.size());
  if (e == cudaSuccess) e = cudaMalloc(&h->d_segs, sizeof(ChainSeg) * segs.size());
  if (e == cudaSuccess) e = cudaMalloc(&h->d_done, h->done_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_steps, steps.data(), sizeof(ChainStepDev) * steps.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_segs, segs.data(), sizeof(ChainSeg) * segs.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("chain: device table allocation failed: %s", cudaGetErrorString(e));
    cudaFree(h->d_steps); cudaFree(h->d_segs); cudaFree(h->d_done);
    delete h;
    return TNB200_ERR_CUDA;
  }
  h->cp.steps = h->d_steps; h->cp.segs = h->d_segs; h->cp.done = h->d_done;
  h->cp.num_tiles = tile0; h->cp.nsegs = (int)segs.size(); h->cp.batch = (int)batch;
  h->cp.stages = stages; h->cp.stage_bytes = max_stage;
  h->cp.debug_no_mma = env_int("TNB200_CHAIN_NOMMA", 0);
  h->smem = (size_t)stages * max_stage + (2 * stages + 4) * 8 + 32 + 4 * kSlabBytes + 1024;
  const long long np = tile0 < pairs ? tile0 : pairs;
  h->grid = (unsigned)(2 * np);
  // two pairs per cluster sharing A by multicast when every step has an even number of N tiles and the grid is full
  {
    // every CTA pair of the launch must be resident (tiles wait on earlier tiles): ask the runtime how many 2-CTA clusters
    // fit (MPS / green-context SM limits, other resident kernels) instead of assuming SMs / 2
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * np)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    cudaError_t qe = h->kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                                  : cudaFuncSetAttribute(gemm_tcgen05_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (qe == cudaSuccess)
      qe = h->kind == 0 ? cudaOccupancyMaxActiveClusters(&ncl, gemm_tcgen05_chain_kernel<0>, &cfg)
                        : cudaOccupancyMaxActiveClusters(&ncl, gemm_tcgen05_chain_kernel<1>, &cfg);
    if (qe != cudaSuccess) { cudaGetLastError(); ncl = 0; }
    if (ncl >= 1 && ncl < np) h->grid = (unsigned)(2 * ncl);
  }
  h->cl = (cl4_ok && np == pairs && (pairs % 2) == 0 && (tile0 % 2) == 0) ? 4 : 2;
  if (h->cl == 4) {
    // every cluster must be resident (tiles wait on each other): a cluster lives inside one GPC, so fewer than SMs / 4
    // clusters may fit — ask the runtime, and keep the 2-CTA kernel when too many SMs would stay idle
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    cudaError_t qe = h->kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_chain4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                                  : cudaFuncSetAttribute(gemm_tcgen05_chain4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (qe == cudaSuccess)
      qe = h->kind == 0 ? cudaOccupancyMaxActiveClusters(&ncl, gemm_tcgen05_chain4_kernel<0>, &cfg)
                        : cudaOccupancyMaxActiveClusters(&ncl, gemm_tcgen05_chain4_kernel<1>, &cfg);
    if (qe != cudaSuccess) { cudaGetLastError(); ncl = 0; }
    if (ncl > pairs / 2) ncl = (int)(pairs / 2);
    const int min_cl = env_int("TNB200_CHAIN_CL4_MIN", (int)(pairs / 2) * 3 / 4);
    if (ncl >= min_cl && ncl >= 1) h->grid = (unsigned)(4 * ncl);
    else h->cl = 2;
    h->clusters = ncl;
    if (env_int("TNB200_CHAIN_VERBOSE", 0)) fprintf(stderr, "[tnb200] chain: 4-CTA clusters resident %d of %d -> cl = %d, grid = %u\n", ncl, (int)(pairs / 2), h->cl, h->grid);
  }
  static bool attr_set[4] = {false, false, false, false};
  const int ai = h->kind + (h->cl == 4 ? 2 : 0);
  if (!attr_set[ai]) {
    if (h->cl == 4)
      e = h->kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_chain4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                       : cudaFuncSetAttribute(gemm_tcgen05_chain4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    else
      e = h->kind == 0 ? cudaFuncSetAttribute(gemm_tcgen05_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                       : cudaFuncSetAttribute(gemm_tcgen05_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("chain: cannot raise dynamic smem: %s", cudaGetErrorString(e)); gemm_chain_destroy(h); return TNB200_ERR_CUDA; }
    attr_set[ai] = true;
  }
  *handle = h;
  return 0;
}

int gemm_chain_launch(void* handle, cudaStream_t st) {
  ChainHandle* h = (ChainHandle*)handle;
  if (!h) return TNB200_ERR_INVALID;
  TNB_CHECK_CUDA(cudaMemsetAsync(h->d_done, 0, h->done_bytes, st));
  if (h->cl == 4) {
    if (h->kind == 0) gemm_tcgen05_chain4_kernel<0><<<h->grid, kThreads, h->smem, st>>>(h->cp);
    else gemm_tcgen05_chain4_kernel<1><<<h->grid, kThreads, h->smem, st>>>(h->cp);
  } else {
    if (h->kind == 0) gemm_tcgen05_chain_kernel<0><<<h->grid, kThreads, h->smem, st>>>(h->cp);
    else gemm_tcgen05_chain_kernel<1><<<h->grid, kThreads, h->smem, st>>>(h->cp);
  }
  TNB_LAUNCH_CHECK();
  count_launch();
  set_kernel_name(h->kind == 0 ? (h->cl == 4 ? "tcgen05_chain4_16" : "tcgen05_chain_16") : (h->cl == 4 ? "tcgen05_chain4_tf32" : "tcgen05_chain_tf32"));
  return 0;
}

int gemm_chain_destroy(void* handle) {
  ChainHandle* h = (ChainHandle*)handle;
  if (!h) return 0;
  cudaFree(h->d_steps); cudaFree(h->d_segs); cudaFree(h->d_done);
  delete h;
  return 0;
}

}  // namespace tnb
