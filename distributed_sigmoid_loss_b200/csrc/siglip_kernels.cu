// sm_100a kernels of the distributed sigmoid (SigLIP) loss hot path.
//
// What the reference does per text chunk (distributed_sigmoid_loss.py:22-33, rwightman_sigmoid_loss.py:49-66):
//     logits = img @ txt_chunk.T * exp(t') + b ; loss = -logsigmoid(labels * logits).sum()
// and, through autograd, two more contractions (G @ txt, G.T @ img) for the gradients.
//
// Here every contraction is a tile loop on the tcgen05 tensor pipe:
//   * one persistent, warp-specialised kernel (TMA producer warp / single-thread MMA issuer / 8 epilogue warps),
//   * operands staged by TMA into 128B-swizzled shared memory, accumulators double-buffered in TMEM,
//   * mode kModeLoss: epilogue turns the S tile into softplus / sigma terms, reduces the three scalar sums and
//     (training) writes the sigma tile as the bf16 operand of the gradient contractions — the logits never
//     exist in HBM,
//   * mode kModeOut: epilogue scales the accumulator by exp(t')/B, adds the fp32 positive-pair rank-1 term and
//     writes fp32 gradients. Two problems (dimg and dtxt) share one launch so that the tile count fills the
//     148 SMs evenly.
// 1-CTA (cta_group::1, 128x256 tiles) and 2-CTA (cta_group::2, 256x256 tiles per SM pair) variants are both
// instantiated; the host picks one.
#include "siglip_kernels.cuh"

namespace siglip {

namespace {

constexpr int kBlockM = 128;   // accumulator rows per CTA (= TMEM lanes)
constexpr int kTileN = 256;    // accumulator columns per tile (= UMMA N)
constexpr int kBlockK = 64;    // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kAllocWarp = 10;  // warps 10 and 11 also run the optional peer pull
constexpr int kNumThreads = 384;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;

template <int kCG>
struct Cfg {
  static constexpr int kTileM = kBlockM * kCG;
  static constexpr int kBRows = kTileN / kCG;                      // B-operand rows held by each CTA
  static constexpr int kABytes = kBlockM * kBlockK * 2;            // 16 KiB
  static constexpr int kBBytes = kBRows * kBlockK * 2;             // 32 / 16 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kCG == 1) ? 4 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct TileCoord {
  int prob;
  int m_blk;
  int n_blk;
};

__device__ __forceinline__ TileCoord decode_tile(const KernelParams& p, int t) {
  TileCoord c;
  const int t0 = p.prob[0].tiles_m * p.prob[0].tiles_n;
  c.prob = (t >= t0) ? 1 : 0;
  const int tt = c.prob ? t - t0 : t;
  const int tn = p.prob[c.prob].tiles_n;
  c.m_blk = tt / tn;
  c.n_blk = tt - c.m_blk * tn;
  return c;
}

// -------------------------------------------------------------------------------------------------
// Epilogue of the loss kernel: one 32-column slab of one accumulator row per thread.
// -------------------------------------------------------------------------------------------------
template <bool kEdge, bool kDiag>
__device__ __forceinline__ void loss_slab(const uint32_t (&v)[32], float t, float b, int row, int col0, int nrows,
                                          int ncols, bool store_g, __nv_bfloat16* g_row, float* g_diag,
                                          float& acc_sp, float& acc_g, float& acc_gs) {
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
      if (row == col0 + j) {                           // positive pair (label +1): softplus(-z), -sigma(-z)
        sp = fmaxf(-z, 0.f) + l;
        g = -((z >= 0.f) ? e * r : r);                 // sigma(-z) without the 1 - sigma(z) cancellation
        g_store = 0.f;                                 // the bf16 operand carries negatives only
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
      packed[j >> 1] = pack_bf16x2(g_prev, g_store);
    } else {
      g_prev = g_store;
    }
  }
  if (store_g) {
    uint4* dst = reinterpret_cast<uint4*>(g_row + col0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dst[q] = make_uint4(packed[4 * q + 0], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
  }
}

// Epilogue of the out kernel: one 32-column slab.
__device__ __forceinline__ void out_slab(const uint32_t (&v)[32], float scale, int row, int col0, const Problem& pr,
                                         float fix) {
  if (row >= pr.M) return;
  float* orow = pr.out + static_cast<long long>(row) * pr.ldo;
  const __nv_bfloat16* xrow = pr.fix_mat ? pr.fix_mat + static_cast<long long>(row) * pr.ldx : nullptr;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = col0 + 4 * q;
    if (c < pr.N) {  // N % 4 == 0 is enforced by the host
      float4 o;
      o.x = __uint_as_float(v[4 * q + 0]);
      o.y = __uint_as_float(v[4 * q + 1]);
      o.z = __uint_as_float(v[4 * q + 2]);
      o.w = __uint_as_float(v[4 * q + 3]);
      if (xrow != nullptr) {
        const uint2 xb = *reinterpret_cast<const uint2*>(xrow + c);
        const float x0 = __uint_as_float(xb.x << 16), x1 = __uint_as_float(xb.x & 0xffff0000u);
        const float x2 = __uint_as_float(xb.y << 16), x3 = __uint_as_float(xb.y & 0xffff0000u);
        o.x = fmaf(fix, x0, o.x);
        o.y = fmaf(fix, x1, o.y);
        o.z = fmaf(fix, x2, o.z);
        o.w = fmaf(fix, x3, o.w);
      }
      o.x *= scale;
      o.y *= scale;
      o.z *= scale;
      o.w *= scale;
      float4* dst = reinterpret_cast<float4*>(orow + c);
      if (pr.beta) {
        const float4 old = *dst;
        o.x += old.x;
        o.y += old.y;
        o.z += old.z;
        o.w += old.w;
      }
      *dst = o;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// The kernel
// -------------------------------------------------------------------------------------------------
template <int kCG, int kMode>
__global__ void __launch_bounds__(kNumThreads, 1)
siglip_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ KernelParams p) {
  using C = Cfg<kCG>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stage bases
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  // barrier map (8 bytes each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + kAccStages + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::kStages + 2 * kAccStages);
  const uint32_t red_smem = tmem_ptr_smem + 16;  // 8 warps x 3 doubles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / kCG;
  const int num_clusters = gridDim.x / kCG;
  const int total_tiles = p.prob[0].tiles_m * p.prob[0].tiles_n +
                          (p.nprob > 1 ? p.prob[1].tiles_m * p.prob[1].tiles_n : 0);

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
      mbar_init(empty_bar(s), 1);
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
  if constexpr (kCG == 2) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == kProducerWarp) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_owner_rank = 0;  // the pair's leader CTA owns the "full" barriers
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const TileCoord tc = decode_tile(p, t);
        const Problem& pr = p.prob[tc.prob];
        const CUtensorMap* tmA = tc.prob ? &tmA1 : &tmA0;
        const CUtensorMap* tmB = tc.prob ? &tmB1 : &tmB0;
        const int m_idx = tc.m_blk * C::kTileM + static_cast<int>(cta_rank) * kBlockM;
        const int n_idx = tc.n_blk * kTileN + static_cast<int>(cta_rank) * C::kBRows;
        const int num_kb = (pr.K + kBlockK - 1) / kBlockK;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.dbg, 1, t, kb);
          const uint32_t sA = smem_base + stage * C::kStageBytes;
          const uint32_t sB = sA + C::kABytes;
          uint32_t fb = full_bar(stage);
          if (cta_rank == 0) mbar_arrive_expect_tx(fb, C::kStageBytes * kCG);
          if constexpr (kCG == 2) fb = mapa_shared(fb, full_owner_rank);
          const int k_idx = kb * kBlockK;
          if (!pr.a_mn) {
            tma_load_2d<kCG>(tmA, fb, sA, k_idx, m_idx);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int h = 0; h < kBlockM / 64; ++h)        // boxes {64 rows, 64 k}
              tma_load_2d<kCG>(tmA, fb, sA + h * 8192, m_idx + 64 * h, k_idx);
          }
          if (!pr.b_mn) {
            tma_load_2d<kCG>(tmB, fb, sB, k_idx, n_idx);  // box {64 k, kBRows rows}
          } else {
#pragma unroll
            for (int h = 0; h < C::kBRows / 64; ++h)
              tma_load_2d<kCG>(tmB, fb, sB + h * 8192, n_idx + 64 * h, k_idx);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    if (cta_rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const TileCoord tc = decode_tile(p, t);
        const Problem& pr = p.prob[tc.prob];
        const uint32_t idesc = make_idesc_bf16(C::kTileM, kTileN, pr.a_mn, pr.b_mn);
        // K-major: 8-row groups 1024 B apart (SBO), K advance 32 B inside the swizzle row.
        // MN-major: 64-element MN blocks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), K advance 16 rows.
        const uint32_t a_lbo = pr.a_mn ? 8192u : 16u, b_lbo = pr.b_mn ? 8192u : 16u;
        const uint32_t a_adv = pr.a_mn ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
        const uint32_t b_adv = pr.b_mn ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
        const int num_kb = (pr.K + kBlockK - 1) / kBlockK;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, p.dbg, 2, t, as);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kTileN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.dbg, 3, t, kb);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * C::kStageBytes;
          const uint32_t sB = sA + C::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sA, a_lbo, 1024u);
          const uint64_t bdesc = make_smem_desc_sw128(sB, b_lbo, 1024u);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            umma_bf16<kCG>(tmem_d, adesc + static_cast<uint64_t>(k * a_adv), bdesc + static_cast<uint64_t>(k * b_adv),
                           idesc, static_cast<uint32_t>((kb | k) != 0));
          }
          umma_commit<kCG>(empty_bar(stage));  // frees the smem stage (both CTAs) when the MMAs retire
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit<kCG>(tmem_full_bar(as));   // accumulator ready for the epilogue (both CTAs)
        if (++as == kAccStages) {
          as = 0;
          aphase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp < kNumEpiWarps) {
    // ===================================== epilogue =====================================
    const int q = warp & 3;       // TMEM lane quarter this warp may touch
    const int half = warp >> 2;   // which 128 columns of the 256-column accumulator
    const int row_in_cta = q * 32 + lane;
    const float t_exact = expf(*p.t_prime);
    const float bias = (kMode == kModeLoss) ? *p.bias : 0.f;
    double d_sp = 0.0, d_g = 0.0, d_gs = 0.0;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t empty_remote[kAccStages];
#pragma unroll
    for (int a = 0; a < kAccStages; ++a) {
      empty_remote[a] = (kCG == 2) ? mapa_shared(tmem_empty_bar(a), 0) : tmem_empty_bar(a);
    }
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const TileCoord tc = decode_tile(p, t);
      const Problem& pr = p.prob[tc.prob];
      const int row = tc.m_blk * C::kTileM + static_cast<int>(cta_rank) * kBlockM + row_in_cta;
      const int col_base = tc.n_blk * kTileN + half * 128;
      mbar_wait(tmem_full_bar(as), aphase, p.dbg, 4, t, as);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(as * kTileN + half * 128) +
                             (static_cast<uint32_t>(q * 32) << 16);
      uint32_t va[32], vb[32];
      float acc_sp = 0.f, acc_g = 0.f, acc_gs = 0.f;

      bool edge = false, diag = false;
      __nv_bfloat16* g_row = nullptr;
      float scale = 0.f, fix = 0.f;
      if constexpr (kMode == kModeLoss) {
        const int tile_m0 = tc.m_blk * C::kTileM, tile_n0 = tc.n_blk * kTileN;
        edge = (tile_m0 + C::kTileM > pr.M) || (tile_n0 + kTileN > pr.N);
        diag = p.own_chunk && (tile_m0 < tile_n0 + kTileN) && (tile_n0 < tile_m0 + C::kTileM);
        g_row = p.G + static_cast<long long>(row) * p.ldg;
      } else {
        scale = t_exact * p.inv_b;
        fix = (pr.fix_vec != nullptr && row < pr.M) ? pr.fix_vec[row] : 0.f;
      }

      auto slab = [&](const uint32_t(&v)[32], int c) {
        const int col0 = col_base + c * 32;
        if constexpr (kMode == kModeLoss) {
          const bool sg = p.store_g != 0;
          if (edge) {
            if (diag)
              loss_slab<true, true>(v, t_exact, bias, row, col0, pr.M, pr.N, sg, g_row, p.g_diag, acc_sp, acc_g, acc_gs);
            else
              loss_slab<true, false>(v, t_exact, bias, row, col0, pr.M, pr.N, sg, g_row, p.g_diag, acc_sp, acc_g, acc_gs);
          } else {
            if (diag)
              loss_slab<false, true>(v, t_exact, bias, row, col0, pr.M, pr.N, sg, g_row, p.g_diag, acc_sp, acc_g, acc_gs);
            else
              loss_slab<false, false>(v, t_exact, bias, row, col0, pr.M, pr.N, sg, g_row, p.g_diag, acc_sp, acc_g, acc_gs);
          }
        } else {
          out_slab(v, scale, row, col0, pr, fix);
        }
      };

      // 4 slabs of 32 columns; the TMEM load of slab c+1 is in flight while slab c is processed.
      tmem_ld_32x32(taddr + 0, va);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + 32, vb);
      slab(va, 0);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + 64, va);
      slab(vb, 1);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + 96, vb);
      slab(va, 2);
      tmem_ld_wait();
      // every TMEM read of this warp for this accumulator stage has landed: hand the stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCG == 2 && cta_rank != 0) {
          mbar_arrive_cluster(empty_remote[as]);
        } else {
          mbar_arrive(tmem_empty_bar(as));
        }
      }
      slab(vb, 3);

      if constexpr (kMode == kModeLoss) {
        d_sp += static_cast<double>(acc_sp);
        d_g += static_cast<double>(acc_g);
        d_gs += static_cast<double>(acc_gs);
      }
      if (++as == kAccStages) {
        as = 0;
        aphase ^= 1u;
      }
    }
    if constexpr (kMode == kModeLoss) {
      // fixed-order reduction: lanes -> warp -> 8 warps -> one slot per CTA (summed later in slot order)
      d_sp = warp_sum(d_sp);
      d_g = warp_sum(d_g);
      d_gs = warp_sum(d_gs);
      double* red = reinterpret_cast<double*>(smem_raw + (red_smem - smem_u32(smem_raw)));
      if (lane == 0) {
        red[warp * 3 + 0] = d_sp;
        red[warp * 3 + 1] = d_g;
        red[warp * 3 + 2] = d_gs;
      }
      named_barrier_sync(1, kNumEpiWarps * 32);
      if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < kNumEpiWarps; ++w) {
          s0 += red[w * 3 + 0];
          s1 += red[w * 3 + 1];
          s2 += red[w * 3 + 2];
        }
        double* slot = p.partials + 4ll * blockIdx.x;
        if (p.accumulate_partials) {
          s0 += slot[0];
          s1 += slot[1];
          s2 += slot[2];
        }
        slot[0] = s0;
        slot[1] = s1;
        slot[2] = s2;
      }
    }
  } else {
    // ============================ warps 10, 11: NVSwitch peer pull ============================
    // The next text chunk is read ONCE from its owner's buffer (P2P over NVLink) into local HBM while this
    // chunk's tiles compute (replaces distributed_utils.py:10-27 neighbour_exchange / the all_gather at
    // distributed_sigmoid_loss.py:35). MMA operands are then fed from local memory only.
    if (p.pull_bytes != 0) {
      if (p.pull_wait_flag != nullptr) {
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (true) {
          unsigned int v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.pull_wait_flag) : "memory");
          if (v >= p.pull_wait_value) break;
          if ((++spins & 0xffu) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 20000000000ull) {
              if (p.dbg != nullptr && lane == 0) {
                p.dbg->block = blockIdx.x;
                p.dbg->thread = threadIdx.x;
                p.dbg->aux0 = v;
                p.dbg->aux1 = p.pull_wait_value;
                p.dbg->code = 5;
                __threadfence_system();
              }
              __trap();
            }
          }
        }
      }
      const unsigned long long n16 = p.pull_bytes >> 4;
      const unsigned long long nthreads = static_cast<unsigned long long>(gridDim.x) * 64ull;
      unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * 64ull + (threadIdx.x - kAllocWarp * 32);
      // 4 independent 16-byte loads in flight per thread to cover the ~2 us NVLink round trip
      for (; i + 3ull * nthreads < n16; i += 4ull * nthreads) {
        uint4 a = p.pull_src[i];
        uint4 b = p.pull_src[i + nthreads];
        uint4 c = p.pull_src[i + 2ull * nthreads];
        uint4 d = p.pull_src[i + 3ull * nthreads];
        p.pull_dst[i] = a;
        p.pull_dst[i + nthreads] = b;
        p.pull_dst[i + 2ull * nthreads] = c;
        p.pull_dst[i + 3ull * nthreads] = d;
      }
      for (; i < n16; i += nthreads) p.pull_dst[i] = p.pull_src[i];
    }
  }

  // ===================================== teardown =====================================
  tc_fence_before();
  if constexpr (kCG == 2) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------
// small helper kernels
// -------------------------------------------------------------------------------------------------
__global__ void finalize_kernel(const double* __restrict__ partials, int nparts, const float* __restrict__ t_prime,
                                float inv_b, float* loss, float* dt_prime, float* dbias) {
  // one warp; fixed summation order => bitwise reproducible for a fixed grid
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 32) {
    s0 += partials[4ll * i + 0];
    s1 += partials[4ll * i + 1];
    s2 += partials[4ll * i + 2];
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (threadIdx.x == 0) {
    const double t = exp(static_cast<double>(*t_prime));
    if (loss) *loss = static_cast<float>(s0 * inv_b);
    if (dbias) *dbias = static_cast<float>(s1 * inv_b);
    if (dt_prime) *dt_prime = static_cast<float>(t * s2 * inv_b);
  }
}

__global__ void zero_partials_kernel(double* partials, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) partials[i] = 0.0;
}

__global__ void reduce_slots_kernel(float* __restrict__ out, const float* const* __restrict__ slots, int nslots,
                                    size_t n4) {
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
    reinterpret_cast<float4*>(out)[i] = acc;
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

__global__ void wait_flags_kernel(const volatile unsigned int* flags, int n, unsigned int value, DebugRecord* dbg) {
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
        if (now - t0 > 20000000000ull) {
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

template <int kCG, int kMode>
int launch_impl(const CUtensorMap* tmA0, const CUtensorMap* tmB0, const CUtensorMap* tmA1, const CUtensorMap* tmB1,
                const KernelParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<kCG>;
  auto kern = siglip_gemm_kernel<kCG, kMode>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int total_tiles =
      p.prob[0].tiles_m * p.prob[0].tiles_n + (p.nprob > 1 ? p.prob[1].tiles_m * p.prob[1].tiles_n : 0);
  int grid = (num_sms / kCG) * kCG;
  if (grid > total_tiles * kCG) grid = total_tiles * kCG;
  if (grid < kCG) grid = kCG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, *tmA0, *tmB0, *tmA1, *tmB1, p);
  return static_cast<int>(e);
}

}  // namespace

size_t gemm_smem_bytes(int cta_group) {
  return cta_group == 2 ? static_cast<size_t>(Cfg<2>::kSmemBytes) : static_cast<size_t>(Cfg<1>::kSmemBytes);
}

int launch_gemm(int cta_group, int mode, const CUtensorMap* tmA0, const CUtensorMap* tmB0, const CUtensorMap* tmA1,
                const CUtensorMap* tmB1, const KernelParams& p, int num_sms, cudaStream_t stream) {
  if (cta_group == 2) {
    return mode == kModeLoss ? launch_impl<2, kModeLoss>(tmA0, tmB0, tmA1, tmB1, p, num_sms, stream)
                             : launch_impl<2, kModeOut>(tmA0, tmB0, tmA1, tmB1, p, num_sms, stream);
  }
  return mode == kModeLoss ? launch_impl<1, kModeLoss>(tmA0, tmB0, tmA1, tmB1, p, num_sms, stream)
                           : launch_impl<1, kModeOut>(tmA0, tmB0, tmA1, tmB1, p, num_sms, stream);
}

int launch_finalize(const double* partials, int nparts, const float* t_prime, float inv_b, float* loss,
                    float* dt_prime, float* dbias, cudaStream_t stream) {
  finalize_kernel<<<1, 32, 0, stream>>>(partials, nparts, t_prime, inv_b, loss, dt_prime, dbias);
  return static_cast<int>(cudaGetLastError());
}

int launch_zero_partials(double* partials, int nparts, cudaStream_t stream) {
  const int n = nparts * 4;
  zero_partials_kernel<<<(n + 255) / 256, 256, 0, stream>>>(partials, n);
  return static_cast<int>(cudaGetLastError());
}

int launch_reduce_slots(float* out, const float* const* slots_dev, int nslots, size_t n, int num_sms,
                        cudaStream_t stream) {
  reduce_slots_kernel<<<num_sms * 4, 256, 0, stream>>>(out, slots_dev, nslots, n / 4);
  return static_cast<int>(cudaGetLastError());
}

int launch_signal_flags(unsigned int* const* flag_ptrs_dev, int n, unsigned int value, cudaStream_t stream) {
  signal_flags_kernel<<<1, 32, 0, stream>>>(flag_ptrs_dev, n, value);
  return static_cast<int>(cudaGetLastError());
}

int launch_wait_flags(const volatile unsigned int* flags, int n, unsigned int value, DebugRecord* dbg,
                      cudaStream_t stream) {
  wait_flags_kernel<<<1, 32, 0, stream>>>(flags, n, value, dbg);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace siglip
