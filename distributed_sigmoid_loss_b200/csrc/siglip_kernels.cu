// sm_100a kernels of the distributed sigmoid (SigLIP) loss hot path.
//
// What the reference does per text chunk (distributed_sigmoid_loss.py:22-33, rwightman_sigmoid_loss.py:49-66):
//     logits = img @ txt_chunk.T * exp(t') + b ; loss = -logsigmoid(labels * logits).sum()
// and, through autograd, two more contractions (G @ txt, G.T @ img) for the gradients.
//
// Here every contraction is a tile loop on the tcgen05 tensor pipe, one persistent warp-specialised kernel template:
//   * 20 warps: 16 epilogue warps (4 per TMEM lane quarter), a TMA producer warp and an MMA warp (warp-uniform loops,
//     one elected lane issues), 2 auxiliary warps (TMEM allocation, NVSwitch peer pulls / folds, operand conversion);
//   * operands staged by TMA into a 128B-swizzled shared-memory ring, 2 x 256-column fp32 accumulators in TMEM so the
//     MMA of tile n+1 overlaps the epilogue of tile n;
//   * kModeLoss: the epilogue turns the S tile into softplus / sigma terms, reduces the three scalar sums and
//     (training) writes the sigma tile as the scaled-fp16 operand of the gradient contractions through TMA stores —
//     the logits never exist in HBM;
//   * kModeOut: the epilogue scales the accumulator by grad_out * exp(t') / B, adds the fp32 positive-pair rank-1
//     term and writes fp32 or bf16 gradients. Two problems (dimg and dtxt) share one launch so that the tile count
//     fills the 148 SMs evenly.
// Instantiated for cta_group::1 (128x256 tiles) and cta_group::2 (256x256 tiles per SM pair, the default), each
// optionally with the B tile TMA-multicast across two vertically adjacent tiles of a cluster.
#include "siglip_kernels.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

namespace siglip {

namespace {

constexpr int kBlockM = 128;   // accumulator rows per CTA (= TMEM lanes)
constexpr int kTileN = 256;    // accumulator columns per tile (= UMMA N)
constexpr int kBlockK = 64;    // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 16;                       // 4 per TMEM lane quarter -> 4 resident per SM sub-partition
constexpr int kEpiColGroups = kNumEpiWarps / 4;        // column groups of the 256-column accumulator
constexpr int kEpiCols = 256 / kEpiColGroups;          // columns per epilogue warp (64)
constexpr int kSlabsPerWarp = kEpiCols / 32;           // 32-column TMEM loads per warp per tile (2)
constexpr int kProducerWarp = kNumEpiWarps;
constexpr int kMmaWarp = kNumEpiWarps + 1;
constexpr int kAllocWarp = kNumEpiWarps + 2;           // this warp and the next also run the optional peer pull
constexpr int kNumThreads = (kNumEpiWarps + 4) * 32;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;

constexpr int kStagingBytesPerWarp = 2048;  // one 32x32 16-bit slab, 64-byte rows, 64B-swizzled (TMA store source)

template <int kCG, int kMode, int kStagesT>
struct Cfg {
  static constexpr int kTileM = kBlockM * kCG;
  static constexpr int kBRows = kTileN / kCG;                      // B-operand rows held by each CTA
  static constexpr int kABytes = kBlockM * kBlockK * 2;            // 16 KiB
  static constexpr int kBBytes = kBRows * kBlockK * 2;             // 32 / 16 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kStagesT;
  static constexpr int kStagingBytes = (kMode == kModeLoss) ? kNumEpiWarps * kStagingBytesPerWarp : 0;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

constexpr float kLog2e = 1.4426950408889634f;

// log1p(e) / e on [0, 1], degree-7 interpolant at Chebyshev nodes; max relative error 3.2e-7 in fp32 Horner form.
// (lg2.approx has 2^-22 ABSOLUTE error near 1, i.e. ~4e-3 relative on log1p(4.5e-5) — not usable here.)
__device__ __forceinline__ float log1p_over_e(float e) {
  float p = -0.00837115291506052f;
  p = fmaf(p, e, 0.04349390044808388f);
  p = fmaf(p, e, -0.1068500280380249f);
  p = fmaf(p, e, 0.1768747717142105f);
  p = fmaf(p, e, -0.24474774301052094f);
  p = fmaf(p, e, 0.3327192962169647f);
  p = fmaf(p, e, -0.49997174739837646f);
  p = fmaf(p, e, 0.9999997615814209f);
  return p;
}

// two fp32 -> one 32-bit word of two 16-bit floats (low half = lo). kF16: IEEE fp16, else bf16.
template <bool kF16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
  uint32_t r;
  if constexpr (kF16) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// sum += x with a running compensation term (Kahan): the error of each fp32 addition is carried into the next one.
// Written with explicit intrinsics so that fast-math style reassociation cannot remove the compensation.
__device__ __forceinline__ void kahan_add(float& sum, float& comp, float x) {
  const float y = __fsub_rn(x, comp);
  const float t = __fadd_rn(sum, y);
  comp = __fsub_rn(__fsub_rn(t, sum), y);
  sum = t;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Bounded wait of the two auxiliary warps (64 threads) of a CTA on `n` consecutive flags written by peer GPUs
// (st.release.sys): ONE thread per CTA polls (acquire, system scope) with a back-off — thousands of threads hammering
// one L2 line slowed the MMA operand traffic of the whole kernel — then the 64 threads meet on a named barrier.
// Traps after timeout_ns (SIGLIP_OPT_PEER_TIMEOUT_MS: minutes by default, like a process-group timeout — a peer may
// legitimately be late by a checkpoint save or an evaluation pass).
__device__ __forceinline__ void wait_peer_flags(const volatile unsigned int* flags, int n, unsigned int value,
                                                unsigned long long timeout_ns, DebugRecord* dbg, unsigned int site) {
  if (threadIdx.x == kAllocWarp * 32) {
    uint64_t t0 = 0;
    uint32_t spins = 0;
    for (int f = 0; f < n; ++f) {
      while (true) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + f) : "memory");
        if (v >= value) break;
        __nanosleep(200);
        if ((++spins & 0xffu) == 0) {
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          if (now - t0 > timeout_ns) {
            if (dbg != nullptr) {
              dbg->block = blockIdx.x;
              dbg->thread = static_cast<unsigned int>(f);
              dbg->aux0 = v;
              dbg->aux1 = value;
              dbg->code = site;
              __threadfence_system();
            }
            __trap();
          }
        }
      }
    }
    __threadfence();  // order the peer data reads of the other 63 threads after the observed flags
  }
  asm volatile("bar.sync 2, 64;" ::: "memory");
}

// Every CTA has finished its share of something: the last one to arrive (ticket) publishes. Called by ONE thread per
// CTA after a barrier that covers the CTA's writers; returns true on the last CTA. The ticket is left at zero.
__device__ __forceinline__ bool last_cta_arrives(unsigned int* ticket) {
  __threadfence_system();                     // this CTA's writes (possibly read by peers over NVLink) before the ticket
  const unsigned int t = atomicAdd(ticket, 1u);
  if (t != gridDim.x - 1) return false;
  atomicExch(ticket, 0u);                     // ready for the next launch
  __threadfence_system();                     // the other CTAs' writes (observed through the ticket) before the signal
  return true;
}

__device__ __forceinline__ void release_store_sys(unsigned int* flag, unsigned int value) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// 16-byte load of peer (NVLink-mapped) or streaming data: no L1 allocation, data is touched once
__device__ __forceinline__ uint4 ld_peer_16(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// t' as the module holds it: fp64 like the reference's parameter (distributed_sigmoid_loss.py:11), or fp32
__device__ __forceinline__ float load_t_prime(const KernelParams& p) {
  return p.tprime_f64 ? static_cast<float>(*reinterpret_cast<const double*>(p.t_prime)) : *p.t_prime;
}

struct TileCoord {
  int prob;
  int m_blk;
  int n_blk;
  int part;    // -1: whole tile; 0 .. sk_parts-1: this work item is one K-slice of a split tile (0 = the owner, which
               // adds the other slices' partial accumulators and runs the epilogue)
  int slot;    // split tiles only: index into the partial-accumulator workspace / arrival counters
};

// Work unit of a cluster: kMC vertically adjacent tiles (same n block, consecutive m blocks) — one per CTA (pair)
// of the cluster, so that the B operand tile is common and can be TMA-multicast.
template <int kMC>
__device__ __forceinline__ int cluster_tiles(const Problem& pr) {
  return ((pr.tiles_m + kMC - 1) / kMC) * pr.tiles_n;
}

// Column tiles of the out kernel (N-major B operand, no multicast) may be 128 wide instead of 256 (Problem::tile_n):
// twice the tiles of half the work each, for shapes whose 256-wide tiles fill the last wave badly.
template <int kMode, int kMC>
__device__ __forceinline__ int tile_stride_n(const Problem& pr) {
  if constexpr (kMode == kModeOut && kMC == 1) {
    return (pr.tile_n == kTileN / 2 && pr.b_mn) ? kTileN / 2 : kTileN;
  } else {
    return kTileN;
  }
}

// Columns the MMA of column tile n_blk computes: 256, or 128 for narrow tiles / a short last tile of the out kernel.
template <int kMode, int kMC>
__device__ __forceinline__ int tile_cols(const Problem& pr, int n_blk) {
  if constexpr (kMode == kModeOut && kMC == 1) {
    if (tile_stride_n<kMode, kMC>(pr) == kTileN / 2) return kTileN / 2;
    return (pr.b_mn && pr.N - n_blk * kTileN <= kTileN / 2) ? kTileN / 2 : kTileN;
  } else {
    return kTileN;
  }
}

// Work item -> tile. Items 0 .. sk_first-1 are whole tiles in schedule order; with split-K (out kernel, kMC == 1) the
// remaining sk_tiles tiles — the ragged last wave — appear sk_parts times, slice-major, so that the slices of one tile
// run on different clusters at the same time.
template <int kMC>
__device__ __forceinline__ TileCoord decode_tile(const KernelParams& p, int t, int mc_rank) {
  TileCoord c;
  c.part = -1;
  c.slot = 0;
  if (kMC == 1 && p.sk_parts > 1 && t >= p.sk_first) {
    const int q = t - p.sk_first;
    c.part = q / p.sk_tiles;
    c.slot = q - c.part * p.sk_tiles;
    t = p.sk_first + c.slot;
  }
  const int t0 = cluster_tiles<kMC>(p.prob[0]);
  c.prob = (t >= t0) ? 1 : 0;
  const int tt = c.prob ? t - t0 : t;
  const int tn = p.prob[c.prob].tiles_n;
  const int mrow = tt / tn;
  c.m_blk = mrow * kMC + mc_rank;   // may be >= tiles_m for the last row when tiles_m is odd: fully masked tile
  c.n_blk = tt - mrow * tn;
  return c;
}

// k-blocks [kb0, kb1) of a work item: everything, or the item's slice of a split tile
__device__ __forceinline__ void item_k_range(const KernelParams& p, const TileCoord& tc, int num_kb, int& kb0,
                                             int& kb1) {
  if (tc.part < 0) {
    kb0 = 0;
    kb1 = num_kb;
  } else {
    kb0 = static_cast<int>(static_cast<long long>(num_kb) * tc.part / p.sk_parts);
    kb1 = static_cast<int>(static_cast<long long>(num_kb) * (tc.part + 1) / p.sk_parts);
  }
}

// -------------------------------------------------------------------------------------------------
// Epilogue of the loss kernel: one 32-column slab of one accumulator row per thread.
//
// Per element (s = <img_i, txt_j>, z = t*s + b, reference distributed_sigmoid_loss.py:24-33):
//   negative pair: term = softplus(z),  g = dterm/dz = sigma(z)
//   positive pair: term = softplus(-z), g = -sigma(-z)            (own chunk diagonal only)
// Sums kept per thread: sum term, sum g, sum g*s (-> loss, dbias, dt').
//
// Fast path (whole warp slab has z < kFastZ, i.e. e = exp(z) < 2^-6, which is where a SigLIP batch lives:
// bias ~ -10): 1 MUFU (ex2) + ~6 packed FMA-pipe instructions per element, sigma and log1p by their series.
// General path: any z, exp(-|z|) + degree-7 log1p polynomial + rcp.
// -------------------------------------------------------------------------------------------------
constexpr float kFastZ = -4.2f;  // e < 0.015 < 2^-6: series truncated after e^2, relative error < e^3 = 3.4e-6

// The sigma slab goes to HBM through shared memory + one TMA store per warp: 4 conflict-free 16-byte
// st.shared per thread instead of 4 strided 16-byte global stores (32 cache lines per instruction).
struct GStore {
  const CUtensorMap* tmap;  // 16-bit [B, B] tensor, box {32 cols, 32 rows}, SWIZZLE_64B
  uint32_t stage;           // this warp's 2 KiB staging buffer (shared::cta address, 512-byte aligned)
  int row0;                 // first row of this warp's 32-row block
  int lane;
  uint64_t policy;          // L2 evict_first: the sigma lines are not read again before they have left the L2
};

__device__ __forceinline__ void store_g_slab(const GStore& gs, int col0, const uint32_t (&packed)[16]) {
  // the previous TMA store of this warp must have finished READING the staging buffer
  if (gs.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  const uint32_t row_addr = gs.stage + static_cast<uint32_t>(gs.lane) * 64u;
  const uint32_t sw = (static_cast<uint32_t>(gs.lane) >> 1) & 3u;   // 64B swizzle: 16-byte chunk ^= (row / 2) % 4
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((c ^ sw) << 4)), "r"(packed[4 * c + 0]),
                 "r"(packed[4 * c + 1]), "r"(packed[4 * c + 2]), "r"(packed[4 * c + 3])
                 : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
  __syncwarp();
  if (gs.lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(gs.tmap)),
                 "r"(gs.stage), "r"(col0), "r"(gs.row0), "l"(gs.policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// Fast path: every z of the slab is < kFastZ, so e = exp(z) < 2^-6 and both sigma(z) = e / (1 + e) and log1p(e) are
// evaluated by their alternating series on the FMA pipe (truncation < 3.4e-6 relative at the edge of the path, < 1e-8
// for the z ~ -10 of a SigLIP batch; the tolerance is 1e-3), two elements per instruction
// (FFMA2 / FMUL2 / FADD2). One MUFU (ex2) per element instead of two: the epilogue of this kernel is bound by the
// MUFU and FMA pipes, not by the tensor pipe it has to keep up with.
template <bool kF16>
__device__ __forceinline__ void loss_slab_fast(const uint32_t (&v)[32], float tl, float bl, int col0, bool store_g,
                                               const GStore& gst, float gscale, float& acc_sp, float& acc_g,
                                               float& acc_gs) {
  uint32_t packed[16];
  const f32x2 tl2 = pack2(tl, tl), bl2 = pack2(bl, bl);
  // sigma is produced already multiplied by the power-of-two scale of the 16-bit operand (exact), and the two sums
  // that use it are un-scaled once per slab
  const f32x2 gs_p = pack2(gscale, gscale), gs_n = pack2(-gscale, -gscale), one = pack2(1.0f, 1.0f);
  const f32x2 c2 = pack2(0.33333334f, 0.33333334f), c1 = pack2(-0.5f, -0.5f);
  f32x2 a_sp = pack2(0.f, 0.f), a_g = pack2(0.f, 0.f), a_gs = pack2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const f32x2 s = pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    const f32x2 t1 = fma2(s, tl2, bl2);                   // z * log2(e)
    float t1a, t1b;
    unpack2(t1, t1a, t1b);
    const f32x2 e = pack2(ex2_approx(t1a), ex2_approx(t1b));   // exp(z), z < kFastZ
    f32x2 q = fma2(e, gs_p, gs_n);                         // S sigma(z) / e = S (1 - e + e^2)   [- e^3 < 3.8e-6 dropped]
    q = fma2(e, q, gs_p);
    const f32x2 g = mul2(e, q);                            // S sigma(z)
    f32x2 l = fma2(e, c2, c1);                             // log1p(e) / e = 1 - e/2 + e^2/3     [- e^3/4 < 1e-6 dropped]
    l = fma2(e, l, one);
    a_sp = fma2(e, l, a_sp);
    a_g = add2(a_g, g);
    a_gs = fma2(g, s, a_gs);
    float g0, g1;
    unpack2(g, g0, g1);
    packed[j] = pack_16x2<kF16>(g0, g1);
  }
  const float inv_s = 1.0f / gscale;
  float x0, x1;
  unpack2(a_sp, x0, x1);
  acc_sp += x0 + x1;
  unpack2(a_g, x0, x1);
  acc_g = fmaf(x0 + x1, inv_s, acc_g);
  unpack2(a_gs, x0, x1);
  acc_gs = fmaf(x0 + x1, inv_s, acc_gs);
  if (store_g) store_g_slab(gst, col0, packed);
}

template <bool kEdge, bool kDiag, bool kF16>
__device__ __forceinline__ void loss_slab(const uint32_t (&v)[32], float t, float b, int row, int col0, int nrows,
                                          int ncols, bool store_g, const GStore& gst, float gscale, float* g_diag,
                                          float& acc_sp, float& acc_g, float& acc_gs, bool on_diag = true) {
  uint32_t packed[16];
  float g_prev = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float s = __uint_as_float(v[j]);
    const float z = fmaf(s, t, b);
    const float e = ex2_approx(-fabsf(z) * kLog2e);    // exp(-|z|) in (0, 1]
    const float l = e * log1p_over_e(e);               // log1p(exp(-|z|))
    const float r = rcp_approx(1.0f + e);              // sigma(|z|)
    const float sig_z = (z >= 0.f) ? r : e * r;        // sigma(z)
    float sp = fmaxf(z, 0.f) + l;                      // softplus(z): negative pair (label -1)
    float g = sig_z;                                   // d softplus(z) / dz
    float g_store = sig_z;
    bool valid = true;
    if constexpr (kEdge) valid = (row < nrows) && (col0 + j < ncols);
    if constexpr (kDiag) {
      if (on_diag && row == col0 + j) {                // positive pair (label +1): softplus(-z), -sigma(-z)
        sp = fmaxf(-z, 0.f) + l;
        g = -((z >= 0.f) ? e * r : r);                 // sigma(-z) without the 1 - sigma(z) cancellation
        g_store = 0.f;                                 // the 16-bit operand carries negatives only
        if (store_g && valid) g_diag[row] = g;
      }
    }
    if constexpr (kEdge) {
      sp = valid ? sp : 0.f;
      g = valid ? g : 0.f;
      g_store = valid ? g_store : 0.f;
    }
    acc_sp += sp;
    acc_g += g;
    acc_gs = fmaf(g, s, acc_gs);
    if (j & 1) {
      packed[j >> 1] = pack_16x2<kF16>(g_prev, g_store * gscale);
    } else {
      g_prev = g_store * gscale;
    }
  }
  if (store_g) store_g_slab(gst, col0, packed);
}

// Epilogue of the out kernel: one 32-column slab.
//   val = scale * (acc * acc_scale + fix * x[row, col]) (+ add_src[row, col]);  written as fp32 or bf16
// Every global load of the slab is issued before the first store: the stores may alias the loads as far as the compiler
// knows, and a load -> store -> load -> store chain (8 exposed L2 round trips per warp and tile) was most of the 15 us
// between the last MMA and the end of the kernel.
__device__ __forceinline__ void out_slab(const uint32_t (&v)[32], float scale, int row, int col0, const Problem& pr,
                                         float fix) {
  if (row >= pr.M) return;
  const float as = pr.acc_scale;   // undoes the power-of-two scaling of the fp16 operands (1 for bf16)
  const float xs = (pr.fix_mat_scale != 0.f) ? pr.fix_mat_scale : 1.0f;
  const float fixs = fix * xs;
  const __nv_bfloat16* xrow = pr.fix_mat ? pr.fix_mat + static_cast<long long>(row) * pr.ldx : nullptr;
  const float* arow = pr.add_src ? pr.add_src + static_cast<long long>(row) * pr.ld_add : nullptr;
#pragma unroll
  for (int h = 0; h < 2; ++h) {         // two halves of 16 columns: all loads of a half are in flight before its stores
    uint4 xb[2];
    float4 ad[4];
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {
      const int c = col0 + 16 * h + 8 * qq;
      const bool in = c < pr.N;
      xb[qq] = (xrow != nullptr && in) ? __ldg(reinterpret_cast<const uint4*>(xrow + c)) : make_uint4(0u, 0u, 0u, 0u);
      ad[2 * qq] = (arow != nullptr && in) ? *reinterpret_cast<const float4*>(arow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      ad[2 * qq + 1] =
          (arow != nullptr && in) ? *reinterpret_cast<const float4*>(arow + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {    // 8 columns at a time (N % 8 == 0 is enforced by the host)
      const int q = 2 * h + qq;
      const int c = col0 + 8 * q;
      if (c < pr.N) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * q + e]) * as;
        if (xrow != nullptr) {
          const uint32_t xw[4] = {xb[qq].x, xb[qq].y, xb[qq].z, xb[qq].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0, x1;
            if (pr.fix_f16) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw[e]));
              x0 = f.x;
              x1 = f.y;
            } else {
              x0 = __uint_as_float(xw[e] << 16);
              x1 = __uint_as_float(xw[e] & 0xffff0000u);
            }
            o[2 * e + 0] = fmaf(fixs, x0, o[2 * e + 0]);
            o[2 * e + 1] = fmaf(fixs, x1, o[2 * e + 1]);
          }
        }
        const float4 a0 = ad[2 * qq], a1 = ad[2 * qq + 1];   // zeros without a running sum
        o[0] = fmaf(o[0], scale, a0.x); o[1] = fmaf(o[1], scale, a0.y);
        o[2] = fmaf(o[2], scale, a0.z); o[3] = fmaf(o[3], scale, a0.w);
        o[4] = fmaf(o[4], scale, a1.x); o[5] = fmaf(o[5], scale, a1.y);
        o[6] = fmaf(o[6], scale, a1.z); o[7] = fmaf(o[7], scale, a1.w);
        if (pr.out_bf16) {
          __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(pr.out) + static_cast<long long>(row) * pr.ldo;
          *reinterpret_cast<uint4*>(orow + c) =
              make_uint4(pack_16x2<false>(o[0], o[1]), pack_16x2<false>(o[2], o[3]), pack_16x2<false>(o[4], o[5]),
                         pack_16x2<false>(o[6], o[7]));
        } else {
          float* orow = reinterpret_cast<float*>(pr.out) + static_cast<long long>(row) * pr.ldo;
          *reinterpret_cast<float4*>(orow + c) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(orow + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// The kernel
// -------------------------------------------------------------------------------------------------
template <int kCG, int kMode, int kStagesT, int kMC>
__global__ void __launch_bounds__(kNumThreads, 1)
siglip_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmG, const __grid_constant__ KernelParams p) {
  using C = Cfg<kCG, kMode, kStagesT>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stage bases
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + C::kStages * C::kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  // barrier map (8 bytes each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + kAccStages + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::kStages + 2 * kAccStages);
  const uint32_t red_smem = tmem_ptr_smem + 16;  // 8 warps x 3 doubles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.aux_trace != nullptr && threadIdx.x == 0) {
    const unsigned long long now = globaltimer_ns();
    if (blockIdx.x == 0) p.aux_trace[4] = now;                                       // kernel entry (CTA 0)
    atomicMax(p.aux_trace + 12, now);                                                // ... of the last CTA to start
    atomicMin(p.aux_trace + 13, now);                                                // ... of the first
  }
  static_assert(kMC == 1 || kMC == 2, "operand multicast across 1 or 2 tiles");
  // Cluster layout: kCG consecutive CTAs form one MMA pair; kMC pairs (or single CTAs) on vertically adjacent tiles
  // share the B operand tile by TMA multicast. cg=2, mc=2 is the 2x2 cluster cuBLAS' nvjet kernels use.
  constexpr int kClusterSize = kCG * kMC;
  const uint32_t crank = (kClusterSize > 1) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = (kCG == 2) ? (crank & 1u) : 0u;                 // rank inside the MMA pair
  const int mc_rank = (kMC > 1) ? static_cast<int>(crank / kCG) : 0;         // which of the cluster's tiles
  const uint32_t leader_rank = crank - cta_rank;                             // cluster rank of this pair's leader
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << kClusterSize) - 1u);  // every CTA of the cluster
  const int cluster_id = blockIdx.x / kClusterSize;
  const int num_clusters = gridDim.x / kClusterSize;
  const int whole_tiles = cluster_tiles<kMC>(p.prob[0]) + (p.nprob > 1 ? cluster_tiles<kMC>(p.prob[1]) : 0);
  const int total_tiles = (kMC == 1 && p.sk_parts > 1) ? p.sk_first + p.sk_tiles * p.sk_parts : whole_tiles;

  if (warp == kProducerWarp && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB0);
    if (p.nprob > 1) {
      prefetch_tmap(&tmA1);
      prefetch_tmap(&tmB1);
    }
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kMC);   // every CTA that multicasts into this stage must see it released by all readers
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), kNumEpiWarps * kCG);
    }
    fence_mbar_init();
  }
  if (warp == kAllocWarp) {
    tmem_alloc<kCG>(tmem_ptr_smem, kTmemCols);
  }
  tc_fence_before();
  if constexpr (kClusterSize > 1) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touched no global memory and
  // may run while the previous kernel of the stream drains its last tiles; from here on its results are needed (and the
  // buffers it read are overwritten). The next kernel of the stream may start ITS set-up as soon as SMs free up.
  if (p.pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (p.aux_trace != nullptr && threadIdx.x == 0) {
    const unsigned long long now = globaltimer_ns();
    if (blockIdx.x == 0) p.aux_trace[5] = now;                                       // set-up done (CTA 0)
    atomicMax(p.aux_trace + 14, now);                                                // ... on the last CTA
  }

  if (warp == kProducerWarp) {
    // ===================================== TMA producer =====================================
    // The whole warp runs the loop (warp-uniform control flow keeps descriptors and barrier addresses in uniform
    // registers); one elected lane issues the TMA instructions.
    {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      // L2 priorities: the embeddings (A and B of the loss kernel, B of the gradient kernel: 32 MiB each) are re-read
      // by every tile wave and must survive the 512 MiB sigma operand streaming through the 126 MB L2 once per pass.
      // The sigma operand itself keeps normal priority: the four column tiles of a row panel (and both CTAs of a pair)
      // read the same lines a little apart in time and rely on finding them in L2 (evict_first there: 2.6 GB of DRAM
      // reads per launch instead of 1.5).
      const uint64_t pol_b = l2_policy_evict_last();
      const uint64_t pol_a = (kMode == kModeLoss) ? pol_b : l2_policy_evict_normal();
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const TileCoord tc = decode_tile<kMC>(p, t, mc_rank);
        const Problem& pr = p.prob[tc.prob];
        const CUtensorMap* tmA = tc.prob ? &tmA1 : &tmA0;
        const CUtensorMap* tmB = tc.prob ? &tmB1 : &tmB0;
        const int m_idx = tc.m_blk * C::kTileM + static_cast<int>(cta_rank) * kBlockM;
        // a last column tile with <= 128 columns left runs as a 128-wide MMA (out mode, N-major B, no multicast):
        // D = 1152 is 4.5 tiles of 256 — without this 10 % of the gradient MMA work would be padding
        const int n_cur = tile_cols<kMode, kMC>(pr, tc.n_blk);
        const int b_rows = n_cur / kCG;                               // B rows this CTA holds for the tile
        const int n_idx = tc.n_blk * tile_stride_n<kMode, kMC>(pr) + static_cast<int>(cta_rank) * b_rows;
        const uint32_t stage_tx = static_cast<uint32_t>(C::kABytes + b_rows * kBlockK * 2) * kCG;
        // elements per 128-byte swizzle row: 64 16-bit values, or 128 8-bit ones (fp8 experiment, K-major only)
        const int kblk = (pr.ab_f16 == 2) ? 2 * kBlockK : kBlockK;
        const int num_kb = (pr.K + kblk - 1) / kblk;
        const int a_mn = pr.a_mn, b_mn = pr.b_mn;
        int kb0, kb1;
        item_k_range(p, tc, num_kb, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.dbg, 1, t, kb, 0, p.wait_stats ? &w_empty : nullptr);
          if (elect_one_sync()) {
            const uint32_t sA = smem_base + stage * C::kStageBytes;
            const uint32_t sB = sA + C::kABytes;
            uint32_t fb = full_bar(stage);
            if (cta_rank == 0) mbar_arrive_expect_tx(fb, stage_tx);
            const uint32_t fb_local = fb;
            if constexpr (kCG == 2) fb = mapa_shared(fb, leader_rank);   // the pair's leader owns the full barriers
            const int k_idx = kb * kblk;
            if (!a_mn) {
              tma_load_2d_hint<kCG>(tmA, fb, sA, k_idx, m_idx, pol_a);  // box {64 k, 128 rows}
            } else {
#pragma unroll
              for (int h = 0; h < kBlockM / 64; ++h)        // boxes {64 rows, 64 k}
                tma_load_2d_hint<kCG>(tmA, fb, sA + h * 8192, m_idx + 64 * h, k_idx, pol_a);
            }
            if constexpr (kMC == 1) {
              if (!b_mn) {
                tma_load_2d_hint<kCG>(tmB, fb, sB, k_idx, n_idx, pol_b);  // box {64 k, kBRows rows}
              } else {
#pragma unroll
                for (int h = 0; h < C::kBRows / 64; ++h)
                  if (h * 64 < b_rows) tma_load_2d_hint<kCG>(tmB, fb, sB + h * 8192, n_idx + 64 * h, k_idx, pol_b);
              }
            } else if constexpr (kCG == 1) {
              // this CTA fetches 1/kMC of the common B tile and multicasts it to every CTA of the cluster
              constexpr int kPart = C::kBRows / kMC;  // rows of B per CTA
              if (!b_mn) {
                tma_load_2d_mcast(tmB, fb, sB + mc_rank * (kPart * kBlockK * 2), k_idx, n_idx + mc_rank * kPart,
                                  kMcMask);         // box {64 k, kPart rows}
              } else {
#pragma unroll
                for (int h = 0; h < kPart / 64; ++h) {
                  const int hh = mc_rank * (kPart / 64) + h;
                  tma_load_2d_mcast(tmB, fb, sB + hh * 8192, n_idx + 64 * hh, k_idx, kMcMask);
                }
              }
            } else {
              // 2x2 cluster: the CTAs with the same rank-in-pair of both pairs hold the same 128 B rows; each of them
              // fetches 64 of those rows and multicasts them to both. The bytes are counted on each pair leader's
              // full barrier (pair bit cleared in the barrier address).
              constexpr int kPart = C::kBRows / kMC;  // 64 rows
              const uint16_t mask = static_cast<uint16_t>(0x5u << cta_rank);   // CTAs {cta_rank, cta_rank + 2}
              if (!b_mn) {
                tma_load_2d_mcast_2sm(tmB, fb_local, sB + mc_rank * (kPart * kBlockK * 2), k_idx,
                                      n_idx + mc_rank * kPart, mask);          // box {64 k, 64 rows}
              } else {
                tma_load_2d_mcast_2sm(tmB, fb_local, sB + mc_rank * 8192, n_idx + 64 * mc_rank, k_idx, mask);
              }
            }
          }
          __syncwarp();
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.wait_stats && lane == 0) p.wait_stats[8ll * blockIdx.x + 0] = static_cast<unsigned long long>(w_empty);
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    // Warp-uniform loop; one elected lane issues tcgen05.mma / tcgen05.commit.
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      long long w_full = 0, w_tmem = 0;
      const bool prof = p.wait_stats != nullptr;
      const long long c_start = clock_cycles();
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const TileCoord tc = decode_tile<kMC>(p, t, mc_rank);
        const Problem& pr = p.prob[tc.prob];
        const uint32_t idesc = make_idesc_bf16(C::kTileM, tile_cols<kMode, kMC>(pr, tc.n_blk), pr.a_mn, pr.b_mn, pr.ab_f16);
        // K-major: 8-row groups 1024 B apart (SBO), K advance 32 B inside the swizzle row.
        // MN-major: 64-element MN blocks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), K advance 16 rows.
        const uint32_t a_lbo = pr.a_mn ? 8192u : 16u, b_lbo = pr.b_mn ? 8192u : 16u;
        const uint32_t a_adv = pr.a_mn ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
        const uint32_t b_adv = pr.b_mn ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
        // descriptors of stage 0; later stages add stage * kStageBytes >> 4 to the start-address field
        const uint64_t adesc0 = make_smem_desc_sw128(smem_base, a_lbo, 1024u);
        const uint64_t bdesc0 = make_smem_desc_sw128(smem_base + C::kABytes, b_lbo, 1024u);
        const bool fp8 = (pr.ab_f16 == 2);
        const int num_kb = (pr.K + (fp8 ? 2 * kBlockK : kBlockK) - 1) / (fp8 ? 2 * kBlockK : kBlockK);
        int kb0, kb1;
        item_k_range(p, tc, num_kb, kb0, kb1);
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, p.dbg, 2, t, as, 0, prof ? &w_tmem : nullptr);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kTileN);
        if (p.aux_trace != nullptr && blockIdx.x == 0 && t == cluster_id) {     // diagnostic, first tile only
          mbar_wait(full_bar(stage), phase, p.dbg, 3, t, kb0, 0, nullptr);
          if (lane == 0) p.aux_trace[6] = globaltimer_ns();                      // first operands have landed
        }
        // The k loop is the critical path of the kernel (one elected lane feeds the tensor pipe): keep it free of
        // anything that is not the four MMAs and the two commits. The 8-bit measurement variant gets its own copy.
        // Everything the four MMAs of a k block need is computed here, in warp-uniform code (uniform registers), from
        // 32-bit arithmetic on the low descriptor word (the start-address field cannot carry out of it: shared memory
        // addresses >> 4 stay below 2^14); the elected branch holds nothing but the issue. This warp shares its
        // scheduler with four epilogue warps: in the loss kernel (busy epilogue) every instruction of this loop shows
        // up as tensor-pipe idle time (138 instead of 128 cycles per MMA with the ~96-instruction loop this replaces).
        const uint32_t a_hi = static_cast<uint32_t>(adesc0 >> 32), b_hi = static_cast<uint32_t>(bdesc0 >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(adesc0), b_lo0 = static_cast<uint32_t>(bdesc0);
        auto desc64 = [](uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; };
        auto issue_tile = [&](auto is_fp8) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait_warp(full_bar(stage), phase, p.dbg, 3, t, kb, prof ? &w_full : nullptr);
            tc_fence_after();
            const uint32_t so = static_cast<uint32_t>(stage) * static_cast<uint32_t>(C::kStageBytes >> 4);
            const uint32_t al = a_lo0 + so, bl = b_lo0 + so;
            const uint32_t acc0 = static_cast<uint32_t>(kb != kb0);
            const bool last = (kb == kb1 - 1);
            const uint32_t ebar = empty_bar(stage);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t ad = desc64(a_hi, al + static_cast<uint32_t>(k) * a_adv);
                const uint64_t bd = desc64(b_hi, bl + static_cast<uint32_t>(k) * b_adv);
                if constexpr (decltype(is_fp8)::value) {   // kind::f8f6f4: 32 e4m3 values (32 bytes) per instruction
                  umma_f8<kCG>(tmem_d, ad, bd, idesc, k == 0 ? acc0 : 1u);
                } else {
                  umma_bf16<kCG>(tmem_d, ad, bd, idesc, k == 0 ? acc0 : 1u);
                }
              }
              if constexpr (kMC > 1 && kCG == 1) {
                umma_commit_mcast(ebar, kMcMask);  // stage is free in every CTA that writes into it
              } else if constexpr (kMC > 1) {
                umma_commit_2sm_mask(ebar, kMcMask);   // all four CTAs of the 2x2 cluster
              } else {
                umma_commit<kCG>(ebar);  // frees the smem stage (both CTAs) when the MMAs retire
              }
              if (last) {                            // accumulator ready for the epilogue warps of this pair
                if constexpr (kMC > 1 && kCG == 2) {
                  umma_commit_2sm_mask(tmem_full_bar(as), static_cast<uint16_t>(0x3u << leader_rank));
                } else {
                  umma_commit<kCG>(tmem_full_bar(as));
                }
              }
            }
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        };
        if (fp8)
          issue_tile(std::true_type{});
        else
          issue_tile(std::false_type{});
        if (++as == kAccStages) {
          as = 0;
          aphase ^= 1u;
        }
      }
      if (p.aux_trace != nullptr && lane == 0) {
        const unsigned long long now = globaltimer_ns();
        if (blockIdx.x == 0) p.aux_trace[7] = now;                    // last MMA issued (CTA 0)
        atomicMax(p.aux_trace + 9, now);                              // ... latest / earliest over the issuing CTAs
        atomicMin(p.aux_trace + 10, now);
      }
      if (prof && lane == 0) {
        p.wait_stats[8ll * blockIdx.x + 1] = static_cast<unsigned long long>(w_full);
        p.wait_stats[8ll * blockIdx.x + 2] = static_cast<unsigned long long>(w_tmem);
        p.wait_stats[8ll * blockIdx.x + 3] = static_cast<unsigned long long>(clock_cycles() - c_start);
      }
    }
  } else if (warp < kNumEpiWarps) {
    // ===================================== epilogue =====================================
    const int q = warp & 3;        // TMEM lane quarter this warp may touch
    const int cgrp = warp >> 2;    // which kEpiCols columns of the 256-column accumulator
    const int row_in_cta = q * 32 + lane;
    const float t_exact = expf(load_t_prime(p));
    const float bias = (kMode == kModeLoss) ? *p.bias : 0.f;
    // loss kernel: z = t_eff * acc + b with t_eff = t * s_scale (the accumulator is 2^8 <img, txt> for fp16 x 16 operands)
    const float s_scale = (kMode == kModeLoss && p.s_scale != 0.f) ? p.s_scale : 1.0f;
    const float t_eff = t_exact * s_scale;
    const float tl = t_eff * kLog2e, bl = bias * kLog2e;
    // per-thread running sums over the tiles of this CTA: compensated fp32 (Kahan) — a DADD per sum, thread and tile
    // was 11 % of the loss kernel's stall samples (the fp64 pipe of this part is narrow); fp64 only at the very end
    float s_sp = 0.f, s_g = 0.f, s_gs = 0.f, c_sp = 0.f, c_g = 0.f, c_gs = 0.f;
    long long w_epi = 0;
    if constexpr (kMode == kModeOut) {
      // backward of the two scalars: saved (upstream gradient 1) * grad_out, by one thread of the launch
      if (blockIdx.x == 0 && threadIdx.x == 0 && p.sc_saved != nullptr) {
        const float g = (p.grad_out != nullptr) ? *p.grad_out : 1.0f;
        if (p.sc_dt_prime) {
          if (p.tprime_f64)
            *reinterpret_cast<double*>(p.sc_dt_prime) = static_cast<double>(p.sc_saved[0] * g);
          else
            *p.sc_dt_prime = p.sc_saved[0] * g;
        }
        if (p.sc_dbias) *p.sc_dbias = p.sc_saved[1] * g;
      }
    }
    const uint64_t g_store_policy = l2_policy_evict_first();
    const long long epi_start = clock_cycles();
    int as = 0;
    uint32_t aphase = 0;
    bool p1_ready = false;
    uint32_t empty_remote[kAccStages];
#pragma unroll
    for (int a = 0; a < kAccStages; ++a) {
      empty_remote[a] = (kCG == 2) ? mapa_shared(tmem_empty_bar(a), leader_rank) : tmem_empty_bar(a);
    }
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const TileCoord tc = decode_tile<kMC>(p, t, mc_rank);
      const Problem& pr = p.prob[tc.prob];
      const int row = tc.m_blk * C::kTileM + static_cast<int>(cta_rank) * kBlockM + row_in_cta;
      const int col_base = tc.n_blk * tile_stride_n<kMode, kMC>(pr) + cgrp * kEpiCols;
      mbar_wait(tmem_full_bar(as), aphase, p.dbg, 4, t, as, p.epi_sleep_ns,
                (p.wait_stats != nullptr && warp == 0) ? &w_epi : nullptr);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(as * kTileN + cgrp * kEpiCols) +
                             (static_cast<uint32_t>(q * 32) << 16);
      float acc_sp = 0.f, acc_g = 0.f, acc_gs = 0.f;

      bool edge = false, diag = false;
      GStore gst;
      gst.tmap = &tmG;
      gst.stage = staging_base + static_cast<uint32_t>(warp) * kStagingBytesPerWarp;
      gst.row0 = tc.m_blk * C::kTileM + static_cast<int>(cta_rank) * kBlockM + q * 32;
      gst.lane = lane;
      gst.policy = g_store_policy;
      float scale = 0.f, fix = 0.f;
      if constexpr (kMode == kModeLoss) {
        const int tile_m0 = tc.m_blk * C::kTileM, tile_n0 = tc.n_blk * kTileN;
        edge = (tile_m0 + C::kTileM > pr.M) || (tile_n0 + kTileN > pr.N);
        diag = p.own_chunk && (tile_m0 < tile_n0 + kTileN) && (tile_n0 < tile_m0 + C::kTileM);
      } else {
        scale = t_exact * p.inv_b * (p.grad_out != nullptr ? *p.grad_out : 1.0f);
        fix = (pr.fix_vec != nullptr && row < pr.M) ? pr.fix_vec[row] : 0.f;
        if (tc.prob == 1 && p.p1_wait_flag != nullptr && !p1_ready) {
          // the dtxt tiles of the LAST gradient launch add the folded sum of the peers' contributions: the fold that
          // completes it runs in this very launch (auxiliary warps of all CTAs) and must have finished everywhere
          if (lane == 0) {
            uint64_t t0 = 0;
            uint32_t spins = 0;
            while (true) {
              unsigned int v;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.p1_wait_flag) : "memory");
              if (v >= p.p1_wait_value) break;
              __nanosleep(500);
              if ((++spins & 0xffu) == 0) {
                const uint64_t now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > p.peer_timeout_ns) {
                  if (p.dbg != nullptr) {
                    p.dbg->block = blockIdx.x;
                    p.dbg->thread = threadIdx.x;
                    p.dbg->aux0 = v;
                    p.dbg->aux1 = p.p1_wait_value;
                    p.dbg->code = 8;
                    __threadfence_system();
                  }
                  __trap();
                }
              }
            }
          }
          __syncwarp();
          p1_ready = true;
        }
      }

      auto slab = [&](uint32_t(&v)[32], int c) {
        const int col0 = col_base + c * 32;
        if constexpr (kMode == kModeLoss) {
          const bool sg = p.store_g != 0;
          // Only the 32x32 slabs that touch the diagonal of a diagonal tile (rows and columns are 32-aligned: row0 ==
          // col0) hold positive pairs, and only the slabs that cross the matrix border of an edge tile need masking:
          // 8 of the 64 slabs of a diagonal 256x256 tile. Every other slab of such a tile takes the fast path like the
          // rest of the matrix — a whole diagonal tile on the general path had a 3-4x longer epilogue, which ended up
          // on the critical path of the clusters that own one (most visible at small B: 16 of 256 tiles at B = 4096).
          const bool slab_diag = diag && (gst.row0 == col0);
          const bool slab_edge = edge && ((gst.row0 + 32 > pr.M) || (col0 + 32 > pr.N));
          if (!slab_edge && !slab_diag) {
            // z is monotone in s (t > 0): the slab is "all very negative" iff max s is
            float smax = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
            for (int j = 2; j < 32; j += 2) smax = max3(smax, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            const bool fast = __all_sync(0xffffffffu, fmaf(smax, t_eff, bias) < kFastZ);
            if (fast)
              loss_slab_fast<true>(v, tl, bl, col0, sg, gst, p.g_scale, acc_sp, acc_g, acc_gs);
            else
              loss_slab<false, false, true>(v, t_eff, bias, row, col0, pr.M, pr.N, sg, gst, p.g_scale, p.g_diag, acc_sp, acc_g, acc_gs);
          } else {
            // border and / or diagonal slabs (a few per chunk): ONE masked variant — two more copies of the unrolled
            // general path only made the kernel's code larger (the loss and gradient kernels alternate and share the
            // instruction caches)
            loss_slab<true, true, true>(v, t_eff, bias, row, col0, pr.M, pr.N, sg, gst, p.g_scale, p.g_diag, acc_sp, acc_g, acc_gs, slab_diag);
          }
        } else {
          if (tc.part < 0) {
            out_slab(v, scale, row, col0, pr, fix);
          } else {
            // split tile: 32x32 fp32 slab of this thread's row <-> workspace, lanes contiguous (512 B per warp access)
            const int slab_id = cgrp * kSlabsPerWarp + c;
            float4* ws = reinterpret_cast<float4*>(p.sk_ws) +
                         ((static_cast<size_t>(tc.slot) * (p.sk_parts - 1)) * kCG + cta_rank) * (8 * 8 * kBlockM) +
                         static_cast<size_t>(slab_id) * (8 * kBlockM) + row_in_cta;
            const size_t part_stride = static_cast<size_t>(kCG) * (8 * 8 * kBlockM);   // float4 per slice
            if (tc.part > 0) {
              float4* dst = ws + static_cast<size_t>(tc.part - 1) * part_stride;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                dst[j4 * kBlockM] = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                                __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
            } else {
#pragma unroll 1
              for (int sp = 0; sp < p.sk_parts - 1; ++sp) {     // fixed order: slice 0 (mine) + 1 + 2 + ...
                const float4* src = ws + static_cast<size_t>(sp) * part_stride;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 a = __ldcg(src + j4 * kBlockM);
                  v[4 * j4] = __float_as_uint(__uint_as_float(v[4 * j4]) + a.x);
                  v[4 * j4 + 1] = __float_as_uint(__uint_as_float(v[4 * j4 + 1]) + a.y);
                  v[4 * j4 + 2] = __float_as_uint(__uint_as_float(v[4 * j4 + 2]) + a.z);
                  v[4 * j4 + 3] = __float_as_uint(__uint_as_float(v[4 * j4 + 3]) + a.w);
                }
              }
              out_slab(v, scale, row, col0, pr, fix);
            }
          }
        }
      };
      if constexpr (kMode == kModeOut) {
        if (tc.part == 0) {
          // owner slice: the other slices' partial accumulators must be in the workspace (they run at the same time
          // on other clusters and wait for nothing, so this cannot deadlock on a co-resident persistent grid)
          if (lane == 0) {
            const unsigned int need = static_cast<unsigned int>((p.sk_parts - 1) * kCG * kNumEpiWarps);
            const unsigned int* ctr = p.sk_counters + 2 * tc.slot;
            uint64_t t0 = 0;
            uint32_t spins = 0;
            while (true) {
              unsigned int cv;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cv) : "l"(ctr) : "memory");
              if (cv >= need) break;
              __nanosleep(100);
              if ((++spins & 0x3ffu) == 0) {
                const uint64_t now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > SIGLIP_WAIT_TIMEOUT_NS) {
                  if (p.dbg != nullptr) {
                    p.dbg->block = blockIdx.x;
                    p.dbg->thread = threadIdx.x;
                    p.dbg->aux0 = cv;
                    p.dbg->aux1 = need;
                    p.dbg->code = 4;
                    __threadfence_system();
                  }
                  __trap();
                }
              }
            }
          }
          __syncwarp();
        }
      }

      // kSlabsPerWarp slabs of 32 columns. With four epilogue warps per SM sub-partition the TMEM load latency
      // of one warp is covered by the arithmetic of the others.
      uint32_t v[32];
      // column groups beyond a short (128-column) tile have nothing to read (warp-uniform)
      const bool cols_active = cgrp * kEpiCols < tile_cols<kMode, kMC>(pr, tc.n_blk);
#pragma unroll
      for (int c = 0; c < kSlabsPerWarp; ++c) {
        if (cols_active) {
          tmem_ld_32x32(taddr + 32 * c, v);
          tmem_ld_wait();
        }
        if (c == kSlabsPerWarp - 1) {
          // every TMEM read of this warp for this accumulator stage has landed: hand the stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kCG == 2 && cta_rank != 0) {
              mbar_arrive_cluster_relaxed(empty_remote[as]);
            } else {
              mbar_arrive_relaxed(tmem_empty_bar(as));
            }
          }
        }
        if (cols_active) slab(v, c);
      }

      if constexpr (kMode == kModeLoss) {
        kahan_add(s_sp, c_sp, acc_sp);
        kahan_add(s_g, c_g, acc_g);
        kahan_add(s_gs, c_gs, acc_gs);
      } else {
        if (tc.part > 0) {
          // my slabs of the partial accumulator are written: arrive (release) on the tile's counter
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(p.sk_counters + 2 * tc.slot, 1u);
        } else if (tc.part == 0) {
          // every owner warp has consumed the partials: the last one re-arms the counters for the next launch
          __syncwarp();
          if (lane == 0) {
            const unsigned int seen = atomicAdd(p.sk_counters + 2 * tc.slot + 1, 1u);
            if (seen == static_cast<unsigned int>(kCG * kNumEpiWarps) - 1u) {
              p.sk_counters[2 * tc.slot] = 0u;
              p.sk_counters[2 * tc.slot + 1] = 0u;
            }
          }
        }
      }
      if (++as == kAccStages) {
        as = 0;
        aphase ^= 1u;
      }
    }
    if (p.aux_trace != nullptr && threadIdx.x == 0) atomicMax(p.aux_trace + 11, globaltimer_ns());   // tiles done
    if (p.wait_stats != nullptr && threadIdx.x == 0) {
      p.wait_stats[8ll * blockIdx.x + 6] = static_cast<unsigned long long>(w_epi);
      p.wait_stats[8ll * blockIdx.x + 7] = static_cast<unsigned long long>(clock_cycles() - epi_start);
    }
    if constexpr (kMode == kModeLoss) {
      // all sigma slabs of this warp must be in global memory before the kernel ends
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      // fixed-order reduction: lanes -> warp -> epilogue warps -> one slot per CTA (summed later in slot order)
      double d_sp = warp_sum(static_cast<double>(s_sp) - static_cast<double>(c_sp));
      double d_g = warp_sum(static_cast<double>(s_g) - static_cast<double>(c_g));
      // sum g * acc -> sum g * <img, txt>
      double d_gs = warp_sum(static_cast<double>(s_gs) - static_cast<double>(c_gs)) * static_cast<double>(s_scale);
      double* red = reinterpret_cast<double*>(smem_raw + (red_smem - smem_u32(smem_raw)));
      if (lane == 0) {
        red[warp * 3 + 0] = d_sp;
        red[warp * 3 + 1] = d_g;
        red[warp * 3 + 2] = d_gs;
      }
      named_barrier_sync(1, kNumEpiWarps * 32);
      if (warp == 0) {
        unsigned int ticket = 0;
        // 16 warps x 3 sums: lanes 0..15 take one warp's triple each, fixed shuffle tree (same order every run)
        double s0 = (lane < kNumEpiWarps) ? red[lane * 3 + 0] : 0.0;
        double s1 = (lane < kNumEpiWarps) ? red[lane * 3 + 1] : 0.0;
        double s2 = (lane < kNumEpiWarps) ? red[lane * 3 + 2] : 0.0;
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
          double* slot = p.partials + 4ll * blockIdx.x;
          if (p.accumulate_partials) {
            s0 += slot[0];
            s1 += slot[1];
            s2 += slot[2];
          }
          slot[0] = s0;
          slot[1] = s1;
          slot[2] = s2;
          if (p.fin_counter != nullptr) {
            __threadfence();                              // the slot before the ticket
            ticket = atomicAdd(p.fin_counter, 1u);
          }
        }
        if (p.fin_counter != nullptr) {
          ticket = __shfl_sync(0xffffffffu, ticket, 0);
          if (ticket == gridDim.x - 1) {
            // last CTA of the forward's last loss kernel: every slot is final. One warp, slot order and a fixed
            // shuffle tree => bitwise reproducible for a fixed grid (the former finalize kernel, without its launch)
            __threadfence();
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            for (unsigned int i = lane; i < gridDim.x; i += 32) {
              f0 += __ldcg(p.partials + 4ll * i + 0);
              f1 += __ldcg(p.partials + 4ll * i + 1);
              f2 += __ldcg(p.partials + 4ll * i + 2);
            }
            f0 = warp_sum(f0);
            f1 = warp_sum(f1);
            f2 = warp_sum(f2);
            if (lane == 0) {
              const double inv_b = static_cast<double>(p.inv_b);
              if (p.fin_loss) *p.fin_loss = static_cast<float>(f0 * inv_b);
              if (p.fin_dbias) *p.fin_dbias = static_cast<float>(f1 * inv_b);
              if (p.fin_dt_prime)
                *p.fin_dt_prime = static_cast<float>(static_cast<double>(expf(load_t_prime(p))) * f2 * inv_b);
              *p.fin_counter = 0u;                        // ready for the next forward
            }
          }
        }
      }
    }
  } else {
    // ===================== the two auxiliary warps: NVSwitch peer pull, conversions, fold =====================
    // A text chunk is read ONCE from its owner's buffer (P2P over NVLink) into local HBM while the previous chunk's
    // tiles compute (replaces distributed_utils.py:10-27 neighbour_exchange / the all_gather at
    // distributed_sigmoid_loss.py:35); MMA operands are then fed from local memory only. The peers' dtxt
    // contributions are folded into a local fp32 accumulator the same way (the reduce-scatter of all_gather's backward,
    // torch functional.py:343-354, spread over the steps instead of exposed at the end). Buffer hand-over between the
    // ranks is by flags: waits before a job, release-stores once every CTA has finished its share of it.
    const int aux_tid = static_cast<int>(threadIdx.x) - kAllocWarp * 32;
    const unsigned long long nthreads = static_cast<unsigned long long>(gridDim.x) * 64ull;
    const unsigned long long tid0 = static_cast<unsigned long long>(blockIdx.x) * 64ull + aux_tid;
    const bool tracer = (p.aux_trace != nullptr && blockIdx.x == 0 && aux_tid == 0);
    if (tracer) p.aux_trace[0] = globaltimer_ns();
    unsigned long long waited_until = 0;
#pragma unroll 1
    for (int j = 0; j < p.naux; ++j) {
      const AuxJob& job = p.aux[j];
      if (job.kind == kAuxNone) continue;
      if (job.wait_flags != nullptr) {
        wait_peer_flags(job.wait_flags, job.wait_n, job.wait_value, p.peer_timeout_ns, p.dbg, job.site);
        if (tracer) waited_until = globaltimer_ns();
      }
      const unsigned long long n16 = job.n16;
      if (job.kind == kAuxCopy) {
        // 8 independent 16-byte loads in flight per thread (~1.2 MB per GPU) to cover the NVLink round trip
        unsigned long long i = tid0;
        for (; i + 7ull * nthreads < n16; i += 8ull * nthreads) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = ld_peer_16(job.src + i + u * nthreads);
#pragma unroll
          for (int u = 0; u < 8; ++u) job.dst[i + u * nthreads] = v[u];
        }
        for (; i < n16; i += nthreads) job.dst[i] = ld_peer_16(job.src + i);
      } else if (job.kind == kAuxCvt) {
        // bf16 -> scaled fp16 copies of the embeddings for the gradient kernel (its sigma operand is fp16, and an
        // MMA cannot mix fp16 with bf16): done here, off the critical path, while the tiles of this chunk compute.
        const float sc = p.cvt_scale;
        const bool plain = job.cvt_copy != 0;
        auto cvt = [&](uint4 v) {
          if (plain) return v;
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = fminf(fmaxf(__uint_as_float(w[q] << 16) * sc, -65504.f), 65504.f);
            const float hi = fminf(fmaxf(__uint_as_float(w[q] & 0xffff0000u) * sc, -65504.f), 65504.f);
            o[q] = pack_16x2<true>(lo, hi);
          }
          return make_uint4(o[0], o[1], o[2], o[3]);
        };
        // 8 independent 16-byte loads in flight per thread: with one the loop was latency-bound (0.26 ms for the two
        // 32 MiB operands of the headline shape, most of the loss kernel's duration, for 128 MiB of traffic)
        unsigned long long i = tid0;
        for (; i + 7ull * nthreads < n16; i += 8ull * nthreads) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = ld_peer_16(job.src + i + u * nthreads);
#pragma unroll
          for (int u = 0; u < 8; ++u) job.dst[i + u * nthreads] = cvt(v[u]);
        }
        for (; i < n16; i += nthreads) job.dst[i] = cvt(ld_peer_16(job.src + i));
      } else if (job.kind == kAuxFold) {
        const bool has_in = job.src2 != nullptr;   // first contribution of a backward pass: plain copy
        const float4* acc_in = reinterpret_cast<const float4*>(job.src2);
        float4* acc_out = reinterpret_cast<float4*>(job.dst);
        unsigned long long i = tid0;
        for (; i + 7ull * nthreads < n16; i += 8ull * nthreads) {
          uint4 rv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) rv[u] = ld_peer_16(job.src + i + u * nthreads);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 a = has_in ? acc_in[i + u * nthreads] : make_float4(0.f, 0.f, 0.f, 0.f);
            a.x += __uint_as_float(rv[u].x);
            a.y += __uint_as_float(rv[u].y);
            a.z += __uint_as_float(rv[u].z);
            a.w += __uint_as_float(rv[u].w);
            acc_out[i + u * nthreads] = a;
          }
        }
        for (; i < n16; i += nthreads) {
          const uint4 rr = ld_peer_16(job.src + i);
          float4 a = has_in ? acc_in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
          a.x += __uint_as_float(rr.x);
          a.y += __uint_as_float(rr.y);
          a.z += __uint_as_float(rr.z);
          a.w += __uint_as_float(rr.w);
          acc_out[i] = a;
        }
      }
      if (job.ticket != nullptr) {
        asm volatile("bar.sync 2, 64;" ::: "memory");      // this CTA's share is written
        if (aux_tid == 0 && last_cta_arrives(job.ticket)) {
          for (int i = 0; i < job.sig_n; ++i) release_store_sys(job.sig_ptrs[i], job.sig_value);
          if (job.done_flag != nullptr)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(job.done_flag), "r"(job.done_value) : "memory");
        }
      }
    }
    if (tracer) {
      p.aux_trace[1] = waited_until;
      p.aux_trace[2] = globaltimer_ns();
    }
  }

  // ===================================== teardown =====================================
  tc_fence_before();
  if constexpr (kClusterSize > 1) {
    cluster_sync_all();   // peers may still multicast into this CTA's smem / arrive on its barriers
  } else {
    __syncthreads();
  }
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, kTmemCols);
  }
  // every warp of this CTA is past its last global write (barrier above): the last CTA of the launch tells the peers
  if (threadIdx.x == 0 && p.aux_trace != nullptr) atomicMin(p.aux_trace + 8, globaltimer_ns());   // first CTA to finish
  if (threadIdx.x == 0 && p.end_ticket != nullptr) {
    if (last_cta_arrives(p.end_ticket)) {
      for (int i = 0; i < p.end_sig_n; ++i) release_store_sys(p.end_sig_ptrs[i], p.end_sig_value);
      if (p.aux_trace != nullptr) p.aux_trace[3] = globaltimer_ns();
    }
  }
}

// -------------------------------------------------------------------------------------------------
// small helper kernels
// -------------------------------------------------------------------------------------------------
__global__ void reduce_slots_kernel(void* __restrict__ out, int out_bf16, const float* const* __restrict__ slots,
                                    int nslots, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = reinterpret_cast<const float4*>(slots[0])[i];
    for (int s = 1; s < nslots; ++s) {
      const float4 v = reinterpret_cast<const float4*>(slots[s])[i];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    if (out_bf16) {
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_16x2<false>(acc.x, acc.y), pack_16x2<false>(acc.z, acc.w));
    } else {
      reinterpret_cast<float4*>(out)[i] = acc;
    }
  }
}

// dst = src * (*g): the whole backward() of the module when the fused step already produced the gradients for an
// upstream gradient of 1. 16-byte vectors when both buffers are 16-byte aligned, element-wise head / tail otherwise.
__device__ __forceinline__ uint4 scale_vec(uint4 v, int is_bf16, float s) {
  if (is_bf16) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      r[q] = pack_16x2<false>(__uint_as_float(w[q] << 16) * s, __uint_as_float(w[q] & 0xffff0000u) * s);
    return make_uint4(r[0], r[1], r[2], r[3]);
  }
  return make_uint4(__float_as_uint(__uint_as_float(v.x) * s), __float_as_uint(__uint_as_float(v.y) * s),
                    __float_as_uint(__uint_as_float(v.z) * s), __float_as_uint(__uint_as_float(v.w) * s));
}

__global__ void scale_kernel(const void* __restrict__ src, void* __restrict__ dst, int is_bf16,
                             const float* __restrict__ g, size_t nbytes, int aligned) {
  const float s = *g;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nvec = aligned ? nbytes / 16 : 0;
  for (size_t i = tid; i < nvec; i += stride)
    reinterpret_cast<uint4*>(dst)[i] = scale_vec(reinterpret_cast<const uint4*>(src)[i], is_bf16, s);
  // elements the vector loop did not cover (everything for unaligned buffers, < 16 bytes otherwise)
  const size_t esz = is_bf16 ? 2 : 4;
  const size_t first = nvec * 16 / esz, nel = nbytes / esz;
  for (size_t i = first + tid; i < nel; i += stride) {
    if (is_bf16) {
      const __nv_bfloat16 x = reinterpret_cast<const __nv_bfloat16*>(src)[i];
      reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(__bfloat162float(x) * s);
    } else {
      reinterpret_cast<float*>(dst)[i] = reinterpret_cast<const float*>(src)[i] * s;
    }
  }
}

__global__ void signal_flags_kernel(unsigned int* const* flag_ptrs, int n, unsigned int value) {
  // everything this stream wrote before the signal must be visible to the peers that observe the flag
  __threadfence_system();
  const int i = threadIdx.x;
  if (i < n) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_ptrs[i]), "r"(value) : "memory");
  }
}

__global__ void wait_flags_kernel(const volatile unsigned int* flags, int n, unsigned int value,
                                  unsigned long long timeout_ns, DebugRecord* dbg) {
  const int i = threadIdx.x;
  if (i < n) {
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (v >= value) break;
      if ((++spins & 0xffu) == 0) {
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          if (dbg != nullptr) {
            dbg->block = i;
            dbg->aux0 = v;
            dbg->aux1 = value;
            dbg->code = 6;
            __threadfence_system();
          }
          __trap();
        }
      }
    }
  }
}

// Mean over the ranks of the two scalar gradients: what wrapping the module in DDP (README.md:20) or
// the toy average_gradients (test_distributed_sigmoid_loss.py:79-83) does with an all_reduce, here as one warp that
// publishes (dt', dbias) in a peer-visible mailbox, signals every rank, waits for every rank and adds the W mailboxes
// in rank order — the same order on every rank, so all ranks end up with bit-identical parameter gradients.
__global__ void allreduce_scalars_kernel(const float* __restrict__ saved, const float* __restrict__ g,
                                         float* mailbox_local, const float* const* __restrict__ mailboxes,
                                         unsigned int* const* __restrict__ signal_ptrs,
                                         const volatile unsigned int* flags_local, int world, unsigned int value,
                                         float* dt_prime, float* dbias, int dtp_f64, unsigned long long timeout_ns,
                                         DebugRecord* dbg) {
  __shared__ float sh[2][32];
  const int i = threadIdx.x;
  const float s = (g != nullptr) ? *g : 1.0f;
  if (i == 0) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mailbox_local), "f"(saved[0] * s) : "memory");
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mailbox_local + 1), "f"(saved[1] * s) : "memory");
    __threadfence_system();
  }
  __syncwarp();
  if (i < world) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(signal_ptrs[i]), "r"(value) : "memory");
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags_local + i) : "memory");
      if (v >= value) break;
      if ((++spins & 0xffu) == 0) {
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          if (dbg != nullptr) {
            dbg->block = i;
            dbg->aux0 = v;
            dbg->aux1 = value;
            dbg->code = 9;
            __threadfence_system();
          }
          __trap();
        }
      }
    }
    float a, b;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(a) : "l"(mailboxes[i]) : "memory");
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(b) : "l"(mailboxes[i] + 1) : "memory");
    sh[0][i] = a;
    sh[1][i] = b;
  }
  __syncwarp();
  if (i == 0) {
    float a = 0.0f, b = 0.0f;
    for (int p = 0; p < world; ++p) {
      a += sh[0][p];
      b += sh[1][p];
    }
    const float inv_w = 1.0f / static_cast<float>(world);
    if (dt_prime) {
      if (dtp_f64)
        *reinterpret_cast<double*>(dt_prime) = static_cast<double>(a * inv_w);
      else
        *dt_prime = a * inv_w;
    }
    if (dbias) *dbias = b * inv_w;
  }
}

// -------------------------------------------------------------------------------------------------
// L2 normalisation fused around the loss (the step the reference's callers run right before it,
// test_distributed_sigmoid_loss.py:99-101). One warp per row; HBM-bound: coalesced 16-byte accesses, fp32 math.
// -------------------------------------------------------------------------------------------------
template <bool kInBf16>
__device__ __forceinline__ void load8(const void* base, size_t idx8, float (&v)[8]) {
  if constexpr (kInBf16) {
    const uint4 w = reinterpret_cast<const uint4*>(base)[idx8];
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[2 * q] = __uint_as_float(u[q] << 16);
      v[2 * q + 1] = __uint_as_float(u[q] & 0xffff0000u);
    }
  } else {
    const float4 a = reinterpret_cast<const float4*>(base)[2 * idx8];
    const float4 b = reinterpret_cast<const float4*>(base)[2 * idx8 + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <bool kOutBf16>
__device__ __forceinline__ void store8(void* base, size_t idx8, const float (&v)[8]) {
  if constexpr (kOutBf16) {
    reinterpret_cast<uint4*>(base)[idx8] = make_uint4(pack_16x2<false>(v[0], v[1]), pack_16x2<false>(v[2], v[3]),
                                                      pack_16x2<false>(v[4], v[5]), pack_16x2<false>(v[6], v[7]));
  } else {
    reinterpret_cast<float4*>(base)[2 * idx8] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(base)[2 * idx8 + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// fp16(v * scale), clamped to the fp16 range: the operand format of the fp32-input path
__device__ __forceinline__ void store8_f16(void* base, size_t idx8, const float (&v)[8], float scale) {
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float lo = fminf(fmaxf(v[2 * q] * scale, -65504.f), 65504.f);
    const float hi = fminf(fmaxf(v[2 * q + 1] * scale, -65504.f), 65504.f);
    o[q] = pack_16x2<true>(lo, hi);
  }
  reinterpret_cast<uint4*>(base)[idx8] = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// xhat[r, :] = bf16(x[r, :] / max(||x[r, :]||, eps)),  inv_norm[r] = 1 / max(||x||, eps)      (F.normalize, eps 1e-12)
// f16_scale > 0: xhat is written as fp16(xhat * f16_scale) instead of bf16 (fp32-input path)
template <bool kInBf16>
__global__ void normalize_fwd_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ xhat,
                                     float* __restrict__ inv_norm, int rows, int d8, float f16_scale) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const size_t row8 = static_cast<size_t>(r) * d8;
    float ss = 0.f;
    for (int c = lane; c < d8; c += 32) {
      float v[8];
      load8<kInBf16>(x, row8 + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[e], v[e], ss);
    }
    ss = warp_sum_f(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[r] = inv;
    for (int c = lane; c < d8; c += 32) {
      float v[8];
      load8<kInBf16>(x, row8 + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= inv;
      if (f16_scale > 0.f)
        store8_f16(xhat, row8 + c, v, f16_scale);
      else
        store8<true>(xhat, row8 + c, v);
    }
  }
}

// dst = bf16(src) or fp16(src * f16_scale): the module's cast of fp32 embeddings to the operand format
__global__ void convert_f32_kernel(const float* __restrict__ src, void* __restrict__ dst, size_t n8, float f16_scale) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    load8<false>(src, i, v);
    if (f16_scale > 0.f)
      store8_f16(dst, i, v, f16_scale);
    else
      store8<true>(dst, i, v);
  }
}

// dx = inv * (dxhat - xhat * <xhat, dxhat>) with xhat = x * inv recomputed in fp32 from the raw input
template <bool kInBf16, bool kGradBf16>
__global__ void normalize_bwd_kernel(const void* __restrict__ x, const float* __restrict__ inv_norm,
                                     const void* __restrict__ dxhat, void* __restrict__ dx, int rows, int d8) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const size_t row8 = static_cast<size_t>(r) * d8;
    const float inv = inv_norm[r];
    float dot = 0.f;
    for (int c = lane; c < d8; c += 32) {
      float v[8], g[8];
      load8<kInBf16>(x, row8 + c, v);
      load8<kGradBf16>(dxhat, row8 + c, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) dot = fmaf(v[e] * inv, g[e], dot);
    }
    dot = warp_sum_f(dot);
    for (int c = lane; c < d8; c += 32) {
      float v[8], g[8], o[8];
      load8<kInBf16>(x, row8 + c, v);
      load8<kGradBf16>(dxhat, row8 + c, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = inv * (g[e] - v[e] * inv * dot);
      store8<kInBf16>(dx, row8 + c, o);
    }
  }
}

template <int kCG, int kMode, int kStages, int kMC>
int launch_impl(const CUtensorMap* tmA0, const CUtensorMap* tmB0, const CUtensorMap* tmA1, const CUtensorMap* tmB1,
                const CUtensorMap* tmG, const KernelParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<kCG, kMode, kStages>;
  constexpr int kClusterSize = kCG * kMC;
  auto kern = siglip_gemm_kernel<kCG, kMode, kStages, kMC>;
  static bool attr_set = false;
  cudaError_t e = cudaSuccess;
  if (!attr_set) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  int total_tiles = ((p.prob[0].tiles_m + kMC - 1) / kMC) * p.prob[0].tiles_n;
  if (p.nprob > 1) total_tiles += ((p.prob[1].tiles_m + kMC - 1) / kMC) * p.prob[1].tiles_n;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterSize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // A persistent kernel must have every cluster co-resident: clusters of 4 do not tile every GPC, so ask the runtime
  // how many fit (once per instantiation) instead of assuming num_sms / cluster size.
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cfg.gridDim = dim3((num_sms / kClusterSize) * kClusterSize);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms / kClusterSize;
    }
    max_clusters = n;
    if (getenv("SIGLIP_DEBUG_WAITSTATS")) printf("[launch] cluster size %d: %d co-resident clusters\n", kClusterSize, n);
  }
  int clusters = max_clusters < num_sms / kClusterSize ? max_clusters : num_sms / kClusterSize;
  KernelParams pl = p;
  pl.sk_parts = 0;
  if (kMode == kModeOut && kMC == 1 && p.sk_request != 0 && p.sk_ws != nullptr && clusters > 1) {
    // Split-K of the ragged last wave: `rem` tiles left after the full waves would occupy rem of `clusters` units for
    // a whole tile time. Cut each of them into S = floor(clusters / rem) K-slices (S * rem <= clusters work items, one
    // per unit): the last wave then lasts ~1/S of a tile. Not worth it when the last wave is more than half full
    // (S would be 1) or the slices get shorter than a handful of k-blocks.
    const int rem = total_tiles % clusters;
    const int min_k = pl.prob[0].K < pl.prob[pl.nprob > 1 ? 1 : 0].K ? pl.prob[0].K : pl.prob[pl.nprob > 1 ? 1 : 0].K;
    const int num_kb = (min_k + kBlockK - 1) / kBlockK;
    if (rem > 0) {
      int S = clusters / rem;
      if (p.sk_request > 1 && S > p.sk_request) S = p.sk_request;
      if (p.sk_request < 0 && S > 4) S = 4;
      while (S > 1 && num_kb / S < 8) --S;
      const size_t need = static_cast<size_t>(rem) * (S - 1) * kCG * (8 * 8 * kBlockM) * sizeof(float4);
      if (S >= 2 && rem <= p.sk_max_tiles && need <= p.sk_ws_bytes) {
        pl.sk_parts = S;
        pl.sk_tiles = rem;
        pl.sk_first = total_tiles - rem;
        total_tiles = pl.sk_first + rem * S;
      }
    }
  }
  if (clusters > total_tiles) clusters = total_tiles;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(clusters * kClusterSize);
  cfg.numAttrs = pl.pdl ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, *tmA0, *tmB0, *tmA1, *tmB1, *tmG, pl);
  return static_cast<int>(e);
}

}  // namespace

int query_max_active_clusters(int cta_group) {
  int n = -1;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(kNumThreads);
  cfg.gridDim = dim3(148);
  if (cta_group == 2) {
    auto kern = siglip_gemm_kernel<2, kModeOut, 7, 1>;
    cfg.dynamicSmemBytes = Cfg<2, kModeOut, 7>::kSmemBytes;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.dynamicSmemBytes);
    cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  } else {
    auto kern = siglip_gemm_kernel<1, kModeOut, 4, 2>;
    cfg.dynamicSmemBytes = Cfg<1, kModeOut, 4>::kSmemBytes;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.dynamicSmemBytes);
    cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  }
  return n;
}

int default_stages(int cta_group, int mode) {
  if (mode == kModeLoss) return cta_group == 2 ? 6 : 4;
  return cta_group == 2 ? 7 : 4;
}

size_t gemm_smem_bytes(int cta_group, int mode) {
  if (mode == kModeLoss)
    return cta_group == 2 ? static_cast<size_t>(Cfg<2, kModeLoss, 6>::kSmemBytes)
                          : static_cast<size_t>(Cfg<1, kModeLoss, 4>::kSmemBytes);
  return cta_group == 2 ? static_cast<size_t>(Cfg<2, kModeOut, 7>::kSmemBytes)
                        : static_cast<size_t>(Cfg<1, kModeOut, 4>::kSmemBytes);
}

#define SIGLIP_LAUNCH(CG, MODE, ST, MC) \
  return launch_impl<CG, MODE, ST, MC>(tmA0, tmB0, tmA1, tmB1, tmG, p, num_sms, stream)

int launch_gemm(int cta_group, int mode, int stages, int mcast, const CUtensorMap* tmA0, const CUtensorMap* tmB0,
                const CUtensorMap* tmA1, const CUtensorMap* tmB1, const CUtensorMap* tmG, const KernelParams& p,
                int num_sms, cudaStream_t stream) {
  if (stages <= 0) stages = default_stages(cta_group, mode);
  if (cta_group == 2 && mcast == 2) {  // 2x2 clusters: two MMA pairs share the B tile by TMA multicast
    if (mode == kModeLoss) {
      switch (stages) {
        case 4: SIGLIP_LAUNCH(2, kModeLoss, 4, 2);
        default: SIGLIP_LAUNCH(2, kModeLoss, 6, 2);
      }
    }
    switch (stages) {
      case 4: SIGLIP_LAUNCH(2, kModeOut, 4, 2);
      default: SIGLIP_LAUNCH(2, kModeOut, 7, 2);
    }
  }
  if (cta_group == 2) {  // 2-CTA MMA pairs, no operand multicast
    if (mode == kModeLoss) {
      switch (stages) {
        case 4: SIGLIP_LAUNCH(2, kModeLoss, 4, 1);
        default: SIGLIP_LAUNCH(2, kModeLoss, 6, 1);
      }
    }
    switch (stages) {
      case 4: SIGLIP_LAUNCH(2, kModeOut, 4, 1);
      default: SIGLIP_LAUNCH(2, kModeOut, 7, 1);
    }
  }
  if (mcast == 2) {  // 1-CTA MMA, clusters of 2 sharing the B tile by TMA multicast
    if (mode == kModeLoss) {
      switch (stages) {
        case 3: SIGLIP_LAUNCH(1, kModeLoss, 3, 2);
        default: SIGLIP_LAUNCH(1, kModeLoss, 4, 2);
      }
    }
    switch (stages) {
      case 3: SIGLIP_LAUNCH(1, kModeOut, 3, 2);
      default: SIGLIP_LAUNCH(1, kModeOut, 4, 2);
    }
  }
  if (mode == kModeLoss) {
    switch (stages) {
      case 3: SIGLIP_LAUNCH(1, kModeLoss, 3, 1);
      default: SIGLIP_LAUNCH(1, kModeLoss, 4, 1);
    }
  }
  switch (stages) {
    case 3: SIGLIP_LAUNCH(1, kModeOut, 3, 1);
    default: SIGLIP_LAUNCH(1, kModeOut, 4, 1);
  }
}
#undef SIGLIP_LAUNCH

int launch_reduce_slots(void* out, int out_bf16, const float* const* slots_dev, int nslots, size_t n, int num_sms,
                        cudaStream_t stream) {
  reduce_slots_kernel<<<num_sms * 4, 256, 0, stream>>>(out, out_bf16, slots_dev, nslots, n / 4);
  return static_cast<int>(cudaGetLastError());
}

int launch_normalize_fwd(const void* x, int in_bf16, __nv_bfloat16* xhat, float* inv_norm, int rows, int D,
                         float f16_scale, int num_sms, cudaStream_t stream) {
  const int grid = num_sms * 8, block = 256;
  if (in_bf16)
    normalize_fwd_kernel<true><<<grid, block, 0, stream>>>(x, xhat, inv_norm, rows, D / 8, f16_scale);
  else
    normalize_fwd_kernel<false><<<grid, block, 0, stream>>>(x, xhat, inv_norm, rows, D / 8, f16_scale);
  return static_cast<int>(cudaGetLastError());
}

int launch_convert_f32(const float* src, void* dst, size_t n, float f16_scale, int num_sms, cudaStream_t stream) {
  convert_f32_kernel<<<num_sms * 4, 256, 0, stream>>>(src, dst, n / 8, f16_scale);
  return static_cast<int>(cudaGetLastError());
}

int launch_normalize_bwd(const void* x, int in_bf16, const float* inv_norm, const void* dxhat, int grad_bf16, void* dx,
                         int rows, int D, int num_sms, cudaStream_t stream) {
  const int grid = num_sms * 8, block = 256;
  if (in_bf16 && grad_bf16)
    normalize_bwd_kernel<true, true><<<grid, block, 0, stream>>>(x, inv_norm, dxhat, dx, rows, D / 8);
  else if (in_bf16)
    normalize_bwd_kernel<true, false><<<grid, block, 0, stream>>>(x, inv_norm, dxhat, dx, rows, D / 8);
  else if (grad_bf16)
    normalize_bwd_kernel<false, true><<<grid, block, 0, stream>>>(x, inv_norm, dxhat, dx, rows, D / 8);
  else
    normalize_bwd_kernel<false, false><<<grid, block, 0, stream>>>(x, inv_norm, dxhat, dx, rows, D / 8);
  return static_cast<int>(cudaGetLastError());
}

int launch_scale(const void* src, void* dst, int is_bf16, const float* g, size_t nbytes, int num_sms,
                 cudaStream_t stream) {
  const int aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  scale_kernel<<<num_sms * 4, 256, 0, stream>>>(src, dst, is_bf16, g, nbytes, aligned);
  return static_cast<int>(cudaGetLastError());
}

int launch_allreduce_scalars(const float* saved, const float* g, float* mailbox_local, const float* const* mailboxes_dev,
                             unsigned int* const* signal_ptrs_dev, const volatile unsigned int* flags_local, int world,
                             unsigned int value, float* dt_prime, float* dbias, int dtp_f64,
                             unsigned long long timeout_ns, DebugRecord* dbg, cudaStream_t stream) {
  allreduce_scalars_kernel<<<1, 32, 0, stream>>>(saved, g, mailbox_local, mailboxes_dev, signal_ptrs_dev, flags_local,
                                                 world, value, dt_prime, dbias, dtp_f64, timeout_ns, dbg);
  return static_cast<int>(cudaGetLastError());
}

int launch_signal_flags(unsigned int* const* flag_ptrs_dev, int n, unsigned int value, cudaStream_t stream) {
  signal_flags_kernel<<<1, 32, 0, stream>>>(flag_ptrs_dev, n, value);
  return static_cast<int>(cudaGetLastError());
}

int launch_wait_flags(const volatile unsigned int* flags, int n, unsigned int value, unsigned long long timeout_ns,
                      DebugRecord* dbg, cudaStream_t stream) {
  wait_flags_kernel<<<1, 32, 0, stream>>>(flags, n, value, timeout_ns, dbg);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace siglip
