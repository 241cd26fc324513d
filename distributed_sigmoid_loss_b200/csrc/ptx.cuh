// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM),
// cluster addressing. Nothing here is library code; every wrapper is one PTX instruction (or a bounded
// spin around one) so that the kernels in siglip_kernels.cu read as the hardware sequence they are.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace siglip {

// ---------------------------------------------------------------------------------------------
// Debug record: written to host-mapped pinned memory just before a __trap() so the host can say
// WHICH wait timed out even though the context is dead afterwards.
// ---------------------------------------------------------------------------------------------
struct DebugRecord {
  unsigned int code;    // 0 = ok; else site id of the timed-out wait
  unsigned int block;
  unsigned int thread;
  unsigned int aux0;
  unsigned int aux1;
  unsigned int aux2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// Hand-back of a TMEM accumulator stage: the arrive only has to follow this warp's tcgen05.ld's (tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync make that so), it publishes no memory writes — a relaxed arrive avoids the
// MEMBAR / ERRBAR sequence a release at cluster scope costs per warp and tile (ncu: 9 % of the loss kernel's stalls).
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

#ifndef SIGLIP_WAIT_TIMEOUT_NS
#define SIGLIP_WAIT_TIMEOUT_NS 4000000000ull  // 4 s: any legitimate in-kernel wait is < 100 ms
#endif

__device__ __forceinline__ long long clock_cycles() {
  long long c;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
  return c;
}

// Bounded wait: a protocol bug must not hang the GPU. On timeout the site id is published to the
// host-mapped debug record and the kernel traps (the launch then fails loudly on the host).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, DebugRecord* dbg, uint32_t site,
                                          uint32_t aux0 = 0, uint32_t aux1 = 0, uint32_t sleep_ns = 0,
                                          long long* waited = nullptr) {
  if (mbar_try_wait(bar, parity)) return;
  const long long c0 = waited ? clock_cycles() : 0;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);  // long waits (epilogue warps during a K loop): do not burn issue slots
    if ((++spins & 0x3ffu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > SIGLIP_WAIT_TIMEOUT_NS) {
        if (dbg != nullptr) {
          dbg->block = blockIdx.x;
          dbg->thread = threadIdx.x;
          dbg->aux0 = aux0;
          dbg->aux1 = aux1;
          dbg->aux2 = parity;
          dbg->code = site;
          __threadfence_system();
        }
        __trap();
      }
    }
  }
  if (waited) *waited += clock_cycles() - c0;
}

// The same wait for a warp whose 32 lanes all poll (warp-uniform loops of the TMA producer / MMA issuer): the warp
// votes on the outcome, so the control flow around the wait is uniform for the compiler as well and the loop's
// counters, barrier addresses and descriptors can stay in uniform registers.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, DebugRecord* dbg, uint32_t site,
                                               uint32_t aux0 = 0, uint32_t aux1 = 0, long long* waited = nullptr) {
  if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
  const long long c0 = waited ? clock_cycles() : 0;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if ((++spins & 0x3ffu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > SIGLIP_WAIT_TIMEOUT_NS) {
        if (dbg != nullptr) {
          dbg->block = blockIdx.x;
          dbg->thread = threadIdx.x;
          dbg->aux0 = aux0;
          dbg->aux1 = aux1;
          dbg->aux2 = parity;
          dbg->code = site;
          __threadfence_system();
        }
        __trap();
      }
    }
  }
  if (waited) *waited += clock_cycles() - c0;
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2-D tiled load global -> shared, completion (bytes) signalled on `bar`.
// kCG == 2: the barrier may live in the peer CTA of the pair (shared::cluster address).
template <int kCG>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  if constexpr (kCG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}

// L2 eviction-priority policies for streamed vs re-used operands (createpolicy; whole line range, fraction 1.0).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// tma_load_2d with an L2 cache policy for the lines it touches.
template <int kCG>
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1,
                                                 uint64_t policy) {
  if constexpr (kCG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
  }
}

// Same load, delivered to the same shared-memory offset of every CTA in `cta_mask` of the cluster; each
// destination CTA's barrier at offset `bar` receives the complete_tx. One L2 read feeds all destinations.
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}

// cta_group::2 flavour of the multicast load (2x2 clusters: two MMA pairs share an operand tile). `bar` is THIS CTA's
// barrier offset with the pair bit cleared (bit 24 of a shared::cluster address selects the CTA within an MMA pair):
// in every destination CTA the complete_tx lands on the barrier of that destination's pair LEADER.
__device__ __forceinline__ void tma_load_2d_mcast_2sm(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads
// ---------------------------------------------------------------------------------------------
template <int kCG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (kCG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCG == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
template <int kCG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (kCG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// Same with 8-bit operands (kind::f8f6f4; here e4m3 x e4m3 -> fp32): K = 32 per instruction. Measurement path only.
template <int kCG>
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kCG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// All previously issued MMAs of this thread -> arrive(1) on `bar` when they retire.
// kCG == 2: arrive on the same barrier offset in both CTAs of the pair.
template <int kCG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (kCG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  } else {
    const uint16_t mask = 0x3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            bar),
        "h"(mask)
        : "memory");
  }
}

// cta_group::1 commit whose arrive is delivered to the same barrier offset in every CTA of `cta_mask`
// (frees a multicast-fed smem stage in all CTAs that write into it).
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// cta_group::2 commit with an explicit cluster CTA mask (2x2 clusters: a stage is released in all four CTAs, an
// accumulator is published to the two CTAs of one pair)
__device__ __forceinline__ void umma_commit_2sm_mask(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Descriptors (bit layouts: PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor")
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle, sm_100 version field = 1.
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1           [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (1 = MN)  [16] B major (1 = MN)    [17,23) N >> 3          [24,29) M >> 4
//   ab_f16: both operands hold IEEE fp16 instead of bf16 (mixing fp16 with bf16 in one MMA faults on sm_100a)
//   ab_f16 == 2: kind::f8f6f4 with e4m3 operands (format code 0 in both fields)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major,
                                                       int ab_f16 = 0) {
  return (1u << 4) | ((ab_f16 ? 0u : 1u) << 7) | ((ab_f16 ? 0u : 1u) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 — two fp32 lanes per instruction on the FMA pipe)
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---------------------------------------------------------------------------------------------
// MUFU approximations (each one SFU instruction)
// ---------------------------------------------------------------------------------------------
// three-input maximum (FMNMX3 on sm_100)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace siglip
