// C ABI (include/siglip_b200.h) over the sm_100a kernels: context + workspaces, TMA descriptor encoding,
// the per-step chunk schedule, CUDA-IPC peer bootstrap. Host-side only; no torch types anywhere.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/siglip_b200.h"
#include "siglip_kernels.cuh"

using siglip::DebugRecord;
using siglip::KernelParams;
using siglip::Problem;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      char buf__[512];                                                                                 \
      snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
               __LINE__);                                                                              \
      return fail(SIGLIP_ERR_CUDA, buf__);                                                             \
    }                                                                                                  \
  } while (0)

#define CKI(expr)                                                                                         \
  do {                                                                                                    \
    int e__ = (expr);                                                                                     \
    if (e__ != 0) {                                                                                       \
      char buf__[512];                                                                                    \
      snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #expr,                                      \
               cudaGetErrorString(static_cast<cudaError_t>(e__)), __FILE__, __LINE__);                    \
      return fail(SIGLIP_ERR_CUDA, buf__);                                                                \
    }                                                                                                     \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// bf16 2-D tensor, `inner` contiguous elements per row, rows `row_stride_elems` apart; 128B-swizzled boxes.
int encode_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                   uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(SIGLIP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return fail(SIGLIP_ERR_INVALID, "operand not 16-byte aligned");
  if ((row_stride_elems * 2) % 16 != 0) return fail(SIGLIP_ERR_INVALID, "row stride not a multiple of 16 bytes");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu)",
             static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer,
             (unsigned long long)row_stride_elems);
    return fail(SIGLIP_ERR_CUDA, buf);
  }
  return 0;
}

// Operand tensor map for the mainloop. mn == 0: stored [rows][K]; mn == 1: stored [K][rows].
int encode_operand(CUtensorMap* m, const void* base, int rows, int K, long long ld, int mn, int box_rows_kmajor) {
  if (!mn) return encode_bf16_2d(m, base, (uint64_t)K, (uint64_t)rows, (uint64_t)ld, 64, (uint32_t)box_rows_kmajor);
  return encode_bf16_2d(m, base, (uint64_t)rows, (uint64_t)K, (uint64_t)ld, 64, 64);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// The sigma operand G of the gradient contractions is stored as IEEE fp16 scaled by 2^14: 11 significant bits
// (bf16 has 8; its 2^-9 rounding was the whole 1e-3 error budget when negatives dominate a gradient), and every
// sigma down to 3.7e-9 (logit -19.4) stays a normal number. The out epilogue multiplies the accumulator by 2^-14.
constexpr float kGScale = 16384.0f;
// The embeddings enter the gradient contractions as fp16(x * 16): exact for every bf16 value with 3.8e-6 <= |x| <= 4094
// (L2-normalised embeddings live in [~1e-4, 1]); the conversion runs inside the loss kernel's idle warps.
constexpr float kXScale = 16.0f;
constexpr int kMaxWorld = 32;
// flags[kind][rank]: counters written by peers with st.release.sys
//   0 text of forward #n is in place        1 dtxt contribution of backward #n, gradient slot j is complete
//   2 forward #n finished pulling everyone's text     3 backward #n finished reading everyone's contributions
//   4 (dt', dbias) of backward #n are in the owner's mailbox (SIGLIP_OPT_SYNC_SCALAR_GRADS)
constexpr int kFlagKinds = 5;
// the mailbox (2 floats) lives behind the flag counters, in the same peer-mapped allocation
constexpr int kMailboxOffset = kFlagKinds * kMaxWorld;
constexpr size_t kFlagBytes = (kMailboxOffset + 4) * sizeof(unsigned int);

struct IpcBlob {
  cudaIpcMemHandle_t txt;
  cudaIpcMemHandle_t slots;
  cudaIpcMemHandle_t flags;
  int rank;
  int device;
  int B;
  int D;
};

}  // namespace

struct siglip_ctx {
  int device = 0, rank = 0, world = 1, B = 0, D = 0, Bp = 0;
  int num_sms = 0;
  // options
  int cta_group = 2;
  int overlap_pull = 1;                  // pull the next text chunk inside the loss kernel
  int overlap_reduce = 1;                // fold the peers' dtxt contributions inside the gradient kernels
  int kernel_timing = 0;
  int stages_loss = 0, stages_grad = 0;  // 0 = kernel default
  int mcast = 1;                         // 2: vertically adjacent tiles share the B tile by TMA multicast
  int grad_bf16 = 0;                     // dimg / dtxt outputs are bf16 instead of fp32
  int epi_sleep_grad_ns = 0;             // back-off of the gradient kernel's epilogue warps while a K loop runs
  int epi_sleep_loss_ns = 0;
  int sync_scalar_grads = 0;             // backward returns the mean over ranks of dt' / dbias
  int bidir = 0;                         // visiting order of the text chunks: r, r+1, r-1, r+2, r-2, ...
  int grad_tile_n = 0;                   // column-tile width of the gradient kernel: 0 = choose, 128, 256
  int input_f16 = 0;                     // img / txt are fp16(x * kXScale) instead of bf16 (fp32-input path)
  int saved_f16 = 0;                     // format of the embeddings of the forward saved for backward
  // diagnostics, read from the environment once at context creation (see include/siglip_b200.h)
  bool dbg_no_gstore = false, dbg_no_cvt = false, dbg_loss_waitstats = false;
  // workspaces
  __nv_bfloat16* txt_all = nullptr;      // [world][B, D] bf16; slot `rank` is what the peers pull (world > 1)
  __nv_bfloat16* G[kMaxWorld] = {};      // per step k: [Bp, Bp] sigma operand (fp16 bits x kGScale), diagonal zeroed
  __nv_bfloat16* img16 = nullptr;        // [B, D] fp16 (x kXScale) images: B operand of the dtxt contraction
  __nv_bfloat16* txt16 = nullptr;        // [world][B, D] fp16 (x kXScale) text chunk of step k: B operand of dimg
  float* g_diag = nullptr;               // [Bp] fp32 positive-pair terms -sigma(-z_ii)
  float* slots = nullptr;                // [world][B, D] fp32 dtxt contributions, slot c is for owner c (world > 1)
  float* dimg_acc = nullptr;             // [B, D] fp32 running dimg over the chunks (world > 1)
  float* dtxt_acc = nullptr;             // [B, D] fp32 running sum of the peers' contributions (world > 1)
  double* partials = nullptr;            // [num_sms][4]
  unsigned int* fin_counter = nullptr;   // ticket counter of the loss kernel's last-CTA finalisation
  unsigned int* flags = nullptr;         // [kFlagKinds][kMaxWorld]
  float* scalars = nullptr;              // [16] device scalars: host API staging, saved dt'/dbias of the last forward
  // peers (index = rank); own entries point at local memory
  __nv_bfloat16* peer_txt[kMaxWorld] = {};
  float* peer_slots[kMaxWorld] = {};
  unsigned int* peer_flags[kMaxWorld] = {};
  bool peers_ready = false;
  bool loopback = false;
  const float** reduce_ptrs_dev = nullptr;   // [world] peer_slots[p] + rank*B*D  (reduction-at-the-end variant)
  const float** final_ptrs_dev = nullptr;    // [2] {dtxt_acc, own slot}: the local last add of the progressive variant
  unsigned int** signal_ptrs_dev = nullptr;  // [kFlagKinds][world]
  const float** mailbox_ptrs_dev = nullptr;  // [world] every rank's (dt', dbias) mailbox
  unsigned int n_fwd = 0, n_bwd = 0;         // forward / backward passes issued (flag counters)
  unsigned long long gen = 0;                // generation of the state saved for backward (0 = none)
  DebugRecord* dbg_host = nullptr;
  DebugRecord* dbg_dev = nullptr;
  unsigned long long launches = 0;
  size_t workspace_bytes = 0;
  std::vector<cudaEvent_t> ev_loss, ev_grad;  // start, stop, start, stop, ...
  size_t ev_loss_used = 0, ev_grad_used = 0;
  // host-API staging
  __nv_bfloat16* h_img[2] = {nullptr, nullptr};   // device staging of the host entries, two sets (pipelining)
  __nv_bfloat16* h_txt[2] = {nullptr, nullptr};
  float* h_dimg = nullptr;
  float* h_dtxt = nullptr;
  float* h_pinned = nullptr;                       // pinned host: [2][8] = {t', bias, -, -, loss, dt', dbias, -} per set
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  unsigned long long host_submitted = 0;           // tickets handed out so far
};

namespace {

constexpr int kSavedScalars = 8;  // scalars[8], scalars[9]: dt', dbias of the last forward for upstream gradient 1

int check_dbg(siglip_ctx* c, const char* where) {
  if (c->dbg_host != nullptr && c->dbg_host->code != 0) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: device wait timed out at site %u (block %u thread %u aux %u %u %u)", where,
             c->dbg_host->code, c->dbg_host->block, c->dbg_host->thread, c->dbg_host->aux0, c->dbg_host->aux1,
             c->dbg_host->aux2);
    return fail(SIGLIP_ERR_CUDA, buf);
  }
  return 0;
}

int timing_mark(siglip_ctx* c, std::vector<cudaEvent_t>& evs, size_t& used, cudaStream_t st) {
  if (!c->kernel_timing) return 0;
  if (used == evs.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    evs.push_back(e);
  }
  CK(cudaEventRecord(evs[used++], st));
  return 0;
}

// Sigma operands are allocated on first use: one per step of the chunk schedule (world of them when training).
int ensure_g(siglip_ctx* c, int k) {
  if (c->G[k] != nullptr) return 0;
  const size_t bytes = static_cast<size_t>(c->Bp) * c->Bp * sizeof(__nv_bfloat16);
  CK(cudaMalloc(reinterpret_cast<void**>(&c->G[k]), bytes));
  c->workspace_bytes += bytes;
  return 0;
}

struct PullJob {
  const void* src = nullptr;
  void* dst = nullptr;
  size_t bytes = 0;
  const unsigned int* flag = nullptr;
  unsigned int value = 0;
};

// The loss kernel over the text chunk of step k: S = img @ txt_c^T on tcgen05, fused scale/bias/log-sigmoid/reduce.
// save: also write the sigma operand G[k] (+ g_diag on the own chunk) and the fp16 copies the gradient kernel needs.
struct FinJob {   // last chunk of a forward: the loss kernel's last CTA writes the results
  float* loss = nullptr;
  float* dt_prime = nullptr;
  float* dbias = nullptr;
};

int run_loss_chunk(siglip_ctx* c, int k, const void* img, const __nv_bfloat16* txt_c, const float* t_prime,
                   const float* bias, bool save, const PullJob& pull, const FinJob* fin, cudaStream_t st) {
  const int cg = c->cta_group;
  const int tile_m = 128 * cg;
  const bool own = (k == 0);
  int rc;
  if ((rc = ensure_g(c, save ? k : 0))) return rc;
  __nv_bfloat16* G = c->G[save ? k : 0];
  CUtensorMap tmA, tmB, tmG;
  if ((rc = encode_operand(&tmA, img, c->B, c->D, c->D, 0, 128))) return rc;
  const int mc = c->mcast;
  if ((rc = encode_operand(&tmB, txt_c, c->B, c->D, c->D, 0, 256 / (cg * mc)))) return rc;
  // store map of the sigma operand: [B, B] inside the padded [Bp, Bp] buffer, one 32x32 slab per TMA store
  if ((rc = encode_bf16_2d(&tmG, G, (uint64_t)c->B, (uint64_t)c->B, (uint64_t)c->Bp, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_64B)))
    return rc;
  KernelParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = 1;
  // fp32-input path: both operands are fp16(x * kXScale), the accumulator is kXScale^2 <img, txt>
  p.prob[0].ab_f16 = c->input_f16;
  p.s_scale = c->input_f16 ? 1.0f / (kXScale * kXScale) : 1.0f;
  p.cvt_copy = c->input_f16;
  p.prob[0].M = c->B;
  p.prob[0].N = c->B;
  p.prob[0].K = c->D;
  p.prob[0].tiles_m = ceil_div(c->B, tile_m);
  p.prob[0].tiles_n = ceil_div(c->B, 256);
  p.t_prime = t_prime;
  p.bias = bias;
  p.inv_b = 1.0f / static_cast<float>(c->B);
  p.G = G;
  p.ldg = c->Bp;
  p.g_diag = c->g_diag;
  p.own_chunk = own ? 1 : 0;
  p.store_g = save ? 1 : 0;
  p.g_scale = kGScale;
  p.partials = c->partials;
  p.accumulate_partials = (k > 0) ? 1 : 0;  // the first chunk of a forward overwrites every slot of the grid
  if (fin != nullptr) {
    p.fin_counter = c->fin_counter;
    p.fin_loss = fin->loss;
    p.fin_dt_prime = fin->dt_prime;
    p.fin_dbias = fin->dbias;
  }
  p.dbg = c->dbg_dev;
  p.pull_src = reinterpret_cast<const uint4*>(pull.src);
  p.pull_dst = reinterpret_cast<uint4*>(pull.dst);
  p.pull_bytes = pull.bytes;
  p.pull_wait_flag = pull.flag;
  p.pull_wait_value = pull.value;
  p.epi_sleep_ns = static_cast<unsigned int>(c->epi_sleep_loss_ns);
  if (save && c->dbg_no_gstore) p.store_g = 0;  // timing experiments only (wrong gradients)
  if (save && !c->dbg_no_cvt) {
    const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
    const unsigned long long n16 = chunk_elems * sizeof(__nv_bfloat16) / 16;
    p.cvt_scale = kXScale;
    p.cvt_src[0] = reinterpret_cast<const uint4*>(txt_c);
    p.cvt_dst[0] = reinterpret_cast<uint4*>(c->txt16 + static_cast<size_t>(k) * chunk_elems);
    p.cvt_n16[0] = n16;
    if (own) {
      p.cvt_src[1] = reinterpret_cast<const uint4*>(img);
      p.cvt_dst[1] = reinterpret_cast<uint4*>(c->img16);
      p.cvt_n16[1] = n16;
    }
  }
  unsigned long long* wstats = nullptr;
  if (c->dbg_loss_waitstats) {   // diagnostic: where the roles of the loss kernel spend their cycles
    CK(cudaMalloc(reinterpret_cast<void**>(&wstats), 8 * 256 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(wstats, 0, 8 * 256 * sizeof(unsigned long long), st));
    p.wait_stats = wstats;
  }
  if ((rc = timing_mark(c, c->ev_loss, c->ev_loss_used, st))) return rc;
  CKI(siglip::launch_gemm(cg, siglip::kModeLoss, c->stages_loss, mc, &tmA, &tmB, &tmA, &tmB, &tmG, p, c->num_sms,
                          st));
  if ((rc = timing_mark(c, c->ev_loss, c->ev_loss_used, st))) return rc;
  if (wstats != nullptr) {
    CK(cudaStreamSynchronize(st));
    std::vector<unsigned long long> h(8 * 256);
    CK(cudaMemcpy(h.data(), wstats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nmma = 0, nall = 0;
    for (int b = 0; b < 256; ++b) {
      if (h[8 * b + 7]) nall++;
      if (h[8 * b + 3]) nmma++;
      for (int j = 0; j < 8; ++j) s[j] += static_cast<double>(h[8 * b + j]);
    }
    printf("[loss waitstats] producer empty-wait %.0f cyc/CTA | MMA (%d issuers): loop %.0f cyc, full-wait %.1f%%, "
           "tmem-wait (epilogue not done) %.1f%% | epilogue warp 0 (%d CTAs): loop %.0f cyc, waiting for an accumulator %.1f%%\n",
           s[0] / (nall ? nall : 1), nmma, s[3] / (nmma ? nmma : 1), 100.0 * s[1] / (s[3] > 0 ? s[3] : 1),
           100.0 * s[2] / (s[3] > 0 ? s[3] : 1), nall, s[7] / (nall ? nall : 1), 100.0 * s[6] / (s[7] > 0 ? s[7] : 1));
    fflush(stdout);
    cudaFree(wstats);
  }
  c->launches++;
  return 0;
}

struct FoldJob {   // dtxt_acc = (in ? in : 0) + remote   by the idle warps of a gradient kernel
  const float* in = nullptr;
  const float* remote = nullptr;
  const unsigned int* flag = nullptr;
  unsigned int value = 0;
};

// The two gradient contractions of the chunk of step k in one launch (g = upstream gradient, device scalar or null):
//   prob 0: dimg (+)= g (t/B) (G @ txt_c  [+ g_diag * txt_own])      A = G K-major,  B = txt16[k] N-major
//   prob 1: dtxt_c  = g (t/B) (G^T @ img  [+ g_diag * img])          A = G M-major,  B = img16 N-major
int run_grad_chunk(siglip_ctx* c, int k, const void* img, const __nv_bfloat16* txt_c, const float* t_prime,
                   const float* grad_out, const float* dimg_add, void* dimg_out, bool dimg_bf16, void* dtxt_out,
                   bool dtxt_bf16, const FoldJob& fold, float* sc_dt_prime, float* sc_dbias, cudaStream_t st) {
  const int cg = c->cta_group;
  const int tile_m = 128 * cg;
  const bool own = (k == 0);
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  int rc;
  if ((rc = encode_operand(&tmA0, c->G[k], c->B, c->B, c->Bp, 0, 128))) return rc;
  if ((rc = encode_operand(&tmB0, c->txt16 + static_cast<size_t>(k) * chunk_elems, c->D, c->B, c->D, 1, 0))) return rc;
  if ((rc = encode_operand(&tmA1, c->G[k], c->B, c->B, c->Bp, 1, 0))) return rc;
  if ((rc = encode_operand(&tmB1, c->img16, c->D, c->B, c->D, 1, 0))) return rc;
  KernelParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = 2;
  // Column-tile width: 256, or 128 when that fills the waves of the persistent grid better (small B: B = 4096, D = 768
  // is 96 tiles of 256 columns on 74 SM pairs = 2 waves for 1.3 waves of work, but 3 half-waves with 128 columns).
  // A narrow tile streams the same sigma panel for half the flops and becomes L2->SM bound: it costs 0.66-0.70 of a
  // full tile, measured (tools/tile_width_ab.py), so it pays only when the 256-wide grid leaves most of a wave empty.
  int tile_n = c->grad_tile_n;
  if (tile_n == 0) {
    const long long units = c->num_sms / cg;
    auto cost = [&](int tn) {   // waves x average tile cost (a 256-wide grid already runs a short last column as 128)
      const long long cols = ceil_div(c->D, tn);
      const int rem = c->D - static_cast<int>(cols - 1) * tn;
      const double row_cost = (tn == 128) ? 0.70 * cols : (cols - 1) + (rem <= 128 ? 0.70 : 1.0);
      const long long tiles = 2ll * ceil_div(c->B, tile_m) * cols;
      return static_cast<double>((tiles + units - 1) / units) * row_cost / static_cast<double>(cols);
    };
    tile_n = (c->mcast == 1 && cost(128) < 0.97 * cost(256)) ? 128 : 256;
  }
  if (c->mcast != 1) tile_n = 256;
  for (int i = 0; i < 2; ++i) {
    Problem& pr = p.prob[i];
    pr.M = c->B;
    pr.N = c->D;
    pr.K = c->B;
    pr.tile_n = tile_n;
    pr.tiles_m = ceil_div(c->B, tile_m);
    pr.tiles_n = ceil_div(c->D, tile_n);
    pr.b_mn = 1;
    pr.ab_f16 = 1;
    pr.acc_scale = 1.0f / (kGScale * kXScale);
    pr.ldo = c->D;
    pr.ldx = c->D;
    pr.fix_vec = own ? c->g_diag : nullptr;
    pr.fix_f16 = c->saved_f16;
    pr.fix_mat_scale = c->saved_f16 ? 1.0f / kXScale : 1.0f;
  }
  p.prob[0].a_mn = 0;
  p.prob[0].out = dimg_out;
  p.prob[0].out_bf16 = dimg_bf16 ? 1 : 0;
  p.prob[0].add_src = dimg_add;
  p.prob[0].ld_add = c->D;
  p.prob[0].fix_mat = own ? txt_c : nullptr;
  p.prob[1].a_mn = 1;
  p.prob[1].out = dtxt_out;
  p.prob[1].out_bf16 = dtxt_bf16 ? 1 : 0;
  p.prob[1].fix_mat = own ? reinterpret_cast<const __nv_bfloat16*>(img) : nullptr;
  if (fold.remote != nullptr) {
    p.acc_in = reinterpret_cast<const float4*>(fold.in);
    p.acc_remote = reinterpret_cast<const float4*>(fold.remote);
    p.acc_out = reinterpret_cast<float4*>(c->dtxt_acc);
    p.acc_n4 = static_cast<unsigned long long>(chunk_elems / 4);
    p.acc_wait_flag = fold.flag;
    p.acc_wait_value = fold.value;
  }
  p.epi_sleep_ns = static_cast<unsigned int>(c->epi_sleep_grad_ns);
  if (sc_dt_prime != nullptr || sc_dbias != nullptr) {   // backward of the two scalars rides on this launch
    p.sc_saved = c->scalars + kSavedScalars;
    p.sc_dt_prime = sc_dt_prime;
    p.sc_dbias = sc_dbias;
  }
  p.t_prime = t_prime;
  p.grad_out = grad_out;
  p.inv_b = 1.0f / static_cast<float>(c->B);
  p.dbg = c->dbg_dev;
  if ((rc = timing_mark(c, c->ev_grad, c->ev_grad_used, st))) return rc;
  CKI(siglip::launch_gemm(cg, siglip::kModeOut, c->stages_grad, c->mcast, &tmA0, &tmB0, &tmA1, &tmB1,
                          &tmA0, p, c->num_sms, st));
  if ((rc = timing_mark(c, c->ev_grad, c->ev_grad_used, st))) return rc;
  c->launches++;
  return 0;
}

// Owner of the text chunk a rank scores at step k. Unidirectional: r, r+1, r+2, ... (the pairs of the reference's ring,
// rwightman_sigmoid_loss.py:108-122). Bidirectional: r, r+1, r-1, r+2, r-2, ... (the order of its bidir exchange,
// rwightman_sigmoid_loss.py:75-107). At every step each owner is read by exactly one rank either way.
inline int step_offset(const siglip_ctx* c, int k) {
  if (!c->bidir) return k;
  return (k & 1) ? (k + 1) / 2 : -(k / 2);
}
inline int step_owner(const siglip_ctx* c, int rank, int k) {
  return ((rank + step_offset(c, k)) % c->world + c->world) % c->world;
}

int signal_peers(siglip_ctx* c, int kind, unsigned int value, cudaStream_t st) {
  CKI(siglip::launch_signal_flags(c->signal_ptrs_dev + kind * c->world, c->world, value, st));
  c->launches++;
  return 0;
}

int wait_peers(siglip_ctx* c, int kind, unsigned int value, cudaStream_t st) {
  CKI(siglip::launch_wait_flags(c->flags + kind * kMaxWorld, c->world, value, c->dbg_dev, st));
  c->launches++;
  return 0;
}

int wait_one(siglip_ctx* c, int kind, int rank, unsigned int value, cudaStream_t st) {
  CKI(siglip::launch_wait_flags(c->flags + kind * kMaxWorld + rank, 1, value, c->dbg_dev, st));
  c->launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Forward: W loss kernels. Step k scores my images against the text chunk owned by rank (r + k) % W — the pairs the
// reference's ring covers (rwightman_sigmoid_loss.py:108-122) without the hop-by-hop forwarding: every chunk is pulled
// straight from its owner through the NVSwitch by the idle warps of the loss kernel of the previous step.
// ---------------------------------------------------------------------------------------------------------------
int forward_impl(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* bias, float* loss,
                 bool save, cudaStream_t st) {
  if (c == nullptr || img == nullptr || txt == nullptr || t_prime == nullptr || bias == nullptr || loss == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null argument");
  if (c->world > 1 && !c->peers_ready)
    return fail(SIGLIP_ERR_STATE, "world > 1 but peer handles were not imported (siglip_ctx_import_handles)");
  int rc;
  if ((rc = check_dbg(c, "siglip forward (previous launch)"))) return rc;
  CK(cudaSetDevice(c->device));
  const int W = c->world, r = c->rank;
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  const size_t chunk_bytes = chunk_elems * sizeof(__nv_bfloat16);
  const unsigned int s = ++c->n_fwd;
  // the saved state is being overwritten (a forward without save still replaces my gathered text slot, which the
  // backward of a multi-rank job reads for the positive-pair term); valid again once a saving forward is enqueued
  if (save || c->world > 1) c->gen = 0;

  const __nv_bfloat16* own_txt = reinterpret_cast<const __nv_bfloat16*>(txt);
  if (W > 1) {
    // peers must have finished pulling my text slot in their previous forward before I overwrite it
    if ((rc = wait_peers(c, 2, s - 1, st))) return rc;
    CK(cudaMemcpyAsync(c->txt_all + r * chunk_elems, txt, chunk_bytes, cudaMemcpyDeviceToDevice, st));
    if ((rc = signal_peers(c, 0, s, st))) return rc;
    own_txt = c->txt_all + r * chunk_elems;
  }
  // (every launch of one forward has the same grid: chunk 0 overwrites its slots, the last CTA sums exactly those)
  // loss, and (for backward) dt' / dbias for an upstream gradient of 1: written by the last CTA of the last chunk
  FinJob fin;
  fin.loss = loss;
  fin.dt_prime = save ? c->scalars + kSavedScalars : nullptr;
  fin.dbias = save ? c->scalars + kSavedScalars + 1 : nullptr;
  for (int k = 0; k < W; ++k) {
    const int cidx = step_owner(c, r, k);
    const __nv_bfloat16* txt_c = (k == 0) ? own_txt : c->txt_all + cidx * chunk_elems;
    PullJob pull;
    if (k + 1 < W) {
      const int nxt = step_owner(c, r, k + 1);
      pull.src = c->peer_txt[nxt] + nxt * chunk_elems;
      pull.dst = c->txt_all + nxt * chunk_elems;
      pull.bytes = chunk_bytes;
      pull.flag = c->flags + 0 * kMaxWorld + nxt;
      pull.value = s;
      if (!c->overlap_pull) {
        // un-overlapped variant (A/B measurements): wait + copy as separate stream operations
        if ((rc = wait_one(c, 0, nxt, s, st))) return rc;
        CK(cudaMemcpyAsync(pull.dst, pull.src, pull.bytes, cudaMemcpyDefault, st));
        pull = PullJob();
      }
    }
    if ((rc = run_loss_chunk(c, k, img, txt_c, t_prime, bias, save, pull, (k == W - 1) ? &fin : nullptr, st)))
      return rc;
  }
  if (W > 1) {
    if ((rc = signal_peers(c, 2, s, st))) return rc;
  }
  CK(cudaGetLastError());
  if (save) {
    c->saved_f16 = c->input_f16;
    static unsigned long long next_gen = 0;
    c->gen = ++next_gen;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Backward: W gradient kernels over the sigma operands the forward saved, the OWN chunk last. Gradient slot j = 1..W
// handles step k = j (j < W) or k = 0 (j == W). My contribution to owner (r + k) % W goes to a local fp32 slot and is
// published with flag value (n-1) W + j; one slot later the owner folds it into its accumulator from inside its own
// gradient kernel (P2P loads over NVSwitch), so every remote contribution has a whole gradient kernel of slack and
// only a local add remains at the end: this is all_gather's backward (reduce-scatter SUM, torch functional.py:343-354;
// reverse ring, distributed_utils.py:75-77, 94-98) without an exposed collective.
// ---------------------------------------------------------------------------------------------------------------
int backward_impl(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* grad_out,
                  void* dimg, void* dtxt, float* dt_prime, float* dbias, cudaStream_t st) {
  if (c == nullptr || img == nullptr || txt == nullptr || t_prime == nullptr || dimg == nullptr || dtxt == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null argument");
  if (c->gen == 0) return fail(SIGLIP_ERR_STATE, "no forward state saved for backward (call siglip_forward with save = 1)");
  int rc;
  if ((rc = check_dbg(c, "siglip backward (previous launch)"))) return rc;
  CK(cudaSetDevice(c->device));
  const int W = c->world, r = c->rank;
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  const unsigned int n = ++c->n_bwd;
  const unsigned int base = (n - 1) * static_cast<unsigned int>(W);
  if (W > 1) {
    // peers must have finished reading my contribution slots of the previous backward before I overwrite them
    if ((rc = wait_peers(c, 3, n - 1, st))) return rc;
  }
  // bf16 text of the own chunk (positive-pair term of dimg): my gathered slot, or the caller's tensor for one rank
  const __nv_bfloat16* own_txt =
      (W > 1) ? c->txt_all + r * chunk_elems : reinterpret_cast<const __nv_bfloat16*>(txt);
  for (int j = 1; j <= W; ++j) {
    const int k = (j < W) ? j : 0;
    const int cidx = step_owner(c, r, k);
    const bool last = (j == W);
    const __nv_bfloat16* txt_c = (k == 0) ? own_txt : nullptr;
    const float* dimg_add = (j > 1) ? c->dimg_acc : nullptr;
    void* dimg_out = last ? dimg : static_cast<void*>(c->dimg_acc);
    void* dtxt_out = (W == 1) ? dtxt : static_cast<void*>(c->slots + cidx * chunk_elems);
    FoldJob fold;
    if (W > 1 && c->overlap_reduce && j >= 2) {
      // the contribution for me that rank p = r - offset(j-1) produced in ITS gradient slot j-1
      const int pr = ((r - step_offset(c, j - 1)) % W + W) % W;
      fold.in = (j == 2) ? nullptr : c->dtxt_acc;
      fold.remote = c->peer_slots[pr] + r * chunk_elems;
      fold.flag = c->flags + 1 * kMaxWorld + pr;
      fold.value = base + static_cast<unsigned int>(j - 1);
    }
    // dt' / dbias = saved * grad_out is written by the last gradient launch (the in-kernel mean over ranks, when
    // enabled, is a separate one-warp kernel below)
    const bool scalars_here = last && !(W > 1 && c->sync_scalar_grads);
    if ((rc = run_grad_chunk(c, k, img, txt_c, t_prime, grad_out, dimg_add, dimg_out, last && c->grad_bf16, dtxt_out,
                             W == 1 && c->grad_bf16, fold, scalars_here ? dt_prime : nullptr,
                             scalars_here ? dbias : nullptr, st)))
      return rc;
    if (W > 1 && (!last || !c->overlap_reduce)) {
      if ((rc = signal_peers(c, 1, base + static_cast<unsigned int>(j), st))) return rc;
    }
  }
  if (W > 1) {
    if (c->overlap_reduce) {
      // dtxt = (sum of the W-1 remote contributions, already local) + my own contribution: a local add
      CKI(siglip::launch_reduce_slots(dtxt, c->grad_bf16, c->final_ptrs_dev, 2, chunk_elems, c->num_sms, st));
    } else {
      if ((rc = wait_peers(c, 1, base + static_cast<unsigned int>(W), st))) return rc;
      CKI(siglip::launch_reduce_slots(dtxt, c->grad_bf16, c->reduce_ptrs_dev, W, chunk_elems, c->num_sms, st));
    }
    c->launches++;
    if ((rc = signal_peers(c, 3, n, st))) return rc;
  }
  if (W > 1 && c->sync_scalar_grads) {
    // a collective: issued on every rank whether or not this caller wants the two values. My mailbox is free again:
    // every peer has signalled the text of a later forward, i.e. finished the backward that read it.
    CKI(siglip::launch_allreduce_scalars(c->scalars + kSavedScalars, grad_out,
                                         reinterpret_cast<float*>(c->flags + kMailboxOffset), c->mailbox_ptrs_dev,
                                         c->signal_ptrs_dev + 4 * W, c->flags + 4 * kMaxWorld, W, n, dt_prime, dbias,
                                         c->dbg_dev, st));
    c->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

const char* siglip_version(void) { return "siglip_b200 0.3.0 sm_100a"; }

const char* siglip_last_error(void) { return g_last_error.c_str(); }

int siglip_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

int siglip_ctx_create(siglip_ctx** out, int device, int rank, int world, int B, int D) {
  if (out == nullptr) return fail(SIGLIP_ERR_INVALID, "out is null");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return fail(SIGLIP_ERR_INVALID, "rank/world out of range (world <= 32)");
  if (B < 1 || D < 8 || (D % 8) != 0) return fail(SIGLIP_ERR_INVALID, "need B >= 1 and D a positive multiple of 8");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SIGLIP_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(SIGLIP_ERR_INVALID, "device ordinal out of range");
  int major = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(SIGLIP_ERR_NO_DEVICE, "device is not compute capability 10.x (B200, sm_100a required)");
  CK(cudaSetDevice(device));
  siglip_ctx* c = new siglip_ctx();
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->B = B;
  c->D = D;
  c->Bp = round_up(B, 256);
  c->dbg_no_gstore = getenv("SIGLIP_DEBUG_NO_GSTORE") != nullptr;
  c->dbg_no_cvt = getenv("SIGLIP_DEBUG_NO_CVT") != nullptr;
  c->dbg_loss_waitstats = getenv("SIGLIP_DEBUG_LOSS_WAITSTATS") != nullptr;
  CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  const size_t chunk_elems = static_cast<size_t>(B) * D;
  size_t total = 0;
  auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
    total += bytes;
    return cudaMalloc(p, bytes);
  };
  CK(alloc(reinterpret_cast<void**>(&c->g_diag), static_cast<size_t>(c->Bp) * sizeof(float)));
  CK(alloc(reinterpret_cast<void**>(&c->img16), chunk_elems * sizeof(__nv_bfloat16)));
  CK(alloc(reinterpret_cast<void**>(&c->txt16), chunk_elems * world * sizeof(__nv_bfloat16)));
  CK(alloc(reinterpret_cast<void**>(&c->partials), static_cast<size_t>(c->num_sms) * 4 * sizeof(double)));
  CK(alloc(reinterpret_cast<void**>(&c->fin_counter), sizeof(unsigned int)));
  CK(cudaMemset(c->fin_counter, 0, sizeof(unsigned int)));
  CK(alloc(reinterpret_cast<void**>(&c->flags), kFlagBytes));
  CK(alloc(reinterpret_cast<void**>(&c->scalars), 16 * sizeof(float)));
  CK(cudaMemset(c->flags, 0, kFlagBytes));
  CK(cudaMemset(c->partials, 0, static_cast<size_t>(c->num_sms) * 4 * sizeof(double)));
  CK(cudaMemset(c->g_diag, 0, static_cast<size_t>(c->Bp) * sizeof(float)));
  CK(cudaMemset(c->scalars, 0, 16 * sizeof(float)));
  if (world > 1) {
    CK(alloc(reinterpret_cast<void**>(&c->txt_all), chunk_elems * world * sizeof(__nv_bfloat16)));
    CK(alloc(reinterpret_cast<void**>(&c->slots), chunk_elems * world * sizeof(float)));
    CK(alloc(reinterpret_cast<void**>(&c->dimg_acc), chunk_elems * sizeof(float)));
    CK(alloc(reinterpret_cast<void**>(&c->dtxt_acc), chunk_elems * sizeof(float)));
    CK(alloc(reinterpret_cast<void**>(&c->reduce_ptrs_dev), world * sizeof(float*)));
    CK(alloc(reinterpret_cast<void**>(&c->final_ptrs_dev), 2 * sizeof(float*)));
    CK(alloc(reinterpret_cast<void**>(&c->signal_ptrs_dev), kFlagKinds * world * sizeof(unsigned int*)));
    CK(alloc(reinterpret_cast<void**>(&c->mailbox_ptrs_dev), world * sizeof(float*)));
    const float* fin[2] = {c->dtxt_acc, c->slots + rank * chunk_elems};
    CK(cudaMemcpy(c->final_ptrs_dev, fin, sizeof(fin), cudaMemcpyHostToDevice));
  }
  CK(cudaHostAlloc(reinterpret_cast<void**>(&c->dbg_host), sizeof(DebugRecord), cudaHostAllocMapped));
  memset(c->dbg_host, 0, sizeof(DebugRecord));
  CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->dbg_dev), c->dbg_host, 0));
  c->peer_txt[rank] = c->txt_all;
  c->peer_slots[rank] = c->slots;
  c->peer_flags[rank] = c->flags;
  c->workspace_bytes = total;
  int rc = ensure_g(c, 0);
  if (rc) return rc;
  CK(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

int siglip_ctx_set_option(siglip_ctx* c, int option, int value) {
  if (c == nullptr) return fail(SIGLIP_ERR_INVALID, "ctx is null");
  switch (option) {
    case SIGLIP_OPT_CTA_GROUP:
      if (value != 1 && value != 2) return fail(SIGLIP_ERR_INVALID, "cta_group must be 1 or 2");
      c->cta_group = value;
      return 0;
    case SIGLIP_OPT_OVERLAP_PULL:
      c->overlap_pull = value ? 1 : 0;
      return 0;
    case SIGLIP_OPT_OVERLAP_REDUCE:
      c->overlap_reduce = value ? 1 : 0;
      return 0;
    case SIGLIP_OPT_GRAD_BF16:
      c->grad_bf16 = value ? 1 : 0;
      return 0;
    case SIGLIP_OPT_MCAST:
      if (value != 1 && value != 2) return fail(SIGLIP_ERR_INVALID, "mcast must be 1 or 2");
      c->mcast = value;
      return 0;
    case SIGLIP_OPT_EPI_SLEEP_GRAD_NS:
      c->epi_sleep_grad_ns = value < 0 ? 0 : value;
      return 0;
    case SIGLIP_OPT_EPI_SLEEP_LOSS_NS:
      c->epi_sleep_loss_ns = value < 0 ? 0 : value;
      return 0;
    case SIGLIP_OPT_GRAD_TILE_N:
      if (value != 0 && value != 128 && value != 256) return fail(SIGLIP_ERR_INVALID, "grad_tile_n must be 0, 128 or 256");
      c->grad_tile_n = value;
      return 0;
    case SIGLIP_OPT_INPUT_F16:
      c->input_f16 = value ? 1 : 0;
      return 0;
    case SIGLIP_OPT_BIDIR:
      c->bidir = value ? 1 : 0;
      c->gen = 0;   // a saved forward was laid out in the other order
      return 0;
    case SIGLIP_OPT_SYNC_SCALAR_GRADS:
      c->sync_scalar_grads = value ? 1 : 0;
      return 0;
    case SIGLIP_OPT_STAGES_LOSS:
      c->stages_loss = value;
      return 0;
    case SIGLIP_OPT_STAGES_GRAD:
      c->stages_grad = value;
      return 0;
    case SIGLIP_OPT_KERNEL_TIMING:
      c->kernel_timing = value ? 1 : 0;
      c->ev_loss_used = c->ev_grad_used = 0;
      return 0;
    default:
      return fail(SIGLIP_ERR_INVALID, "unknown option");
  }
}

size_t siglip_ctx_workspace_bytes(const siglip_ctx* c) { return c ? c->workspace_bytes : 0; }

size_t siglip_ctx_handle_bytes(void) { return sizeof(IpcBlob); }

int siglip_ctx_export_handles(siglip_ctx* c, void* out_bytes, size_t capacity) {
  if (c == nullptr || out_bytes == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (capacity < sizeof(IpcBlob)) return fail(SIGLIP_ERR_INVALID, "handle buffer too small");
  if (c->world == 1) return fail(SIGLIP_ERR_STATE, "world == 1 has no peers to export to");
  CK(cudaSetDevice(c->device));
  IpcBlob b;
  memset(&b, 0, sizeof(b));
  CK(cudaIpcGetMemHandle(&b.txt, c->txt_all));
  CK(cudaIpcGetMemHandle(&b.slots, c->slots));
  CK(cudaIpcGetMemHandle(&b.flags, c->flags));
  b.rank = c->rank;
  b.device = c->device;
  b.B = c->B;
  b.D = c->D;
  memcpy(out_bytes, &b, sizeof(b));
  return 0;
}

static int publish_peer_tables(siglip_ctx* c) {
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  std::vector<const float*> red(c->world);
  for (int p = 0; p < c->world; ++p) red[p] = c->peer_slots[p] + c->rank * chunk_elems;
  CK(cudaMemcpy(c->reduce_ptrs_dev, red.data(), c->world * sizeof(float*), cudaMemcpyHostToDevice));
  std::vector<unsigned int*> sig(kFlagKinds * c->world);
  for (int k = 0; k < kFlagKinds; ++k)
    for (int p = 0; p < c->world; ++p)
      sig[k * c->world + p] = c->peer_flags[p] + k * kMaxWorld + (c->loopback ? p : c->rank);
  CK(cudaMemcpy(c->signal_ptrs_dev, sig.data(), sig.size() * sizeof(unsigned int*), cudaMemcpyHostToDevice));
  std::vector<const float*> mb(c->world);
  for (int p = 0; p < c->world; ++p) mb[p] = reinterpret_cast<const float*>(c->peer_flags[p] + kMailboxOffset);
  CK(cudaMemcpy(c->mailbox_ptrs_dev, mb.data(), c->world * sizeof(float*), cudaMemcpyHostToDevice));
  c->peers_ready = true;
  return 0;
}

int siglip_ctx_import_handles(siglip_ctx* c, const void* all_ranks_bytes, size_t bytes_per_rank) {
  if (c == nullptr || all_ranks_bytes == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (bytes_per_rank != sizeof(IpcBlob)) return fail(SIGLIP_ERR_INVALID, "bytes_per_rank != siglip_ctx_handle_bytes()");
  if (c->world == 1) return fail(SIGLIP_ERR_STATE, "world == 1 has no peers to import");
  CK(cudaSetDevice(c->device));
  const char* base = static_cast<const char*>(all_ranks_bytes);
  for (int p = 0; p < c->world; ++p) {
    IpcBlob b;
    memcpy(&b, base + static_cast<size_t>(p) * bytes_per_rank, sizeof(b));
    if (b.rank != p) return fail(SIGLIP_ERR_INVALID, "handle blobs are not ordered by rank");
    if (b.B != c->B || b.D != c->D)
      return fail(SIGLIP_ERR_INVALID, "peer context has a different (B, D): every rank must use the same batch");
    if (p == c->rank) continue;
    void *pt = nullptr, *ps = nullptr, *pf = nullptr;
    CK(cudaIpcOpenMemHandle(&pt, b.txt, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&ps, b.slots, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&pf, b.flags, cudaIpcMemLazyEnablePeerAccess));
    c->peer_txt[p] = static_cast<__nv_bfloat16*>(pt);
    c->peer_slots[p] = static_cast<float*>(ps);
    c->peer_flags[p] = static_cast<unsigned int*>(pf);
  }
  return publish_peer_tables(c);
}

int siglip_forward(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* bias,
                   float* loss, int save_for_backward, void* cuda_stream) {
  return forward_impl(c, img, txt, t_prime, bias, loss, save_for_backward != 0, static_cast<cudaStream_t>(cuda_stream));
}

unsigned long long siglip_ctx_saved_generation(const siglip_ctx* c) { return c ? c->gen : 0ull; }

int siglip_backward(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* grad_out,
                    void* dimg, void* dtxt, float* dt_prime, float* dbias, void* cuda_stream) {
  return backward_impl(c, img, txt, t_prime, grad_out, dimg, dtxt, dt_prime, dbias,
                       static_cast<cudaStream_t>(cuda_stream));
}

int siglip_fwd_bwd(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* bias,
                   float* loss, void* dimg, void* dtxt, float* dt_prime, float* dbias, void* cuda_stream) {
  if (dimg == nullptr || dtxt == nullptr || dt_prime == nullptr || dbias == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null gradient pointer");
  int rc = siglip_forward(c, img, txt, t_prime, bias, loss, 1, cuda_stream);
  if (rc) return rc;
  return siglip_backward(c, img, txt, t_prime, nullptr, dimg, dtxt, dt_prime, dbias, cuda_stream);
}

int siglip_fwd(siglip_ctx* c, const void* img, const void* txt, const float* t_prime, const float* bias, float* loss,
               void* cuda_stream) {
  return siglip_forward(c, img, txt, t_prime, bias, loss, 0, cuda_stream);
}

// device scalars of host-entry set s: t', bias, loss, dt', dbias
static inline float* host_set_scalars(siglip_ctx* c, int s) { return c->scalars + (s ? 10 : 0); }

static int host_entry_init(siglip_ctx* c) {
  if (c->h_img[0] != nullptr) return 0;
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  for (int s = 0; s < 2; ++s) {
    CK(cudaMalloc(reinterpret_cast<void**>(&c->h_img[s]), chunk_elems * sizeof(__nv_bfloat16)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->h_txt[s]), chunk_elems * sizeof(__nv_bfloat16)));
    CK(cudaEventCreateWithFlags(&c->ev_h2d[s], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_done[s], cudaEventDisableTiming));
  }
  CK(cudaMalloc(reinterpret_cast<void**>(&c->h_dimg), chunk_elems * sizeof(float)));
  CK(cudaMalloc(reinterpret_cast<void**>(&c->h_dtxt), chunk_elems * sizeof(float)));
  CK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_pinned), 16 * sizeof(float), cudaHostAllocDefault));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  c->workspace_bytes += chunk_elems * (4 * sizeof(__nv_bfloat16) + 2 * sizeof(float));
  return 0;
}

// Enqueue one end-to-end step: the host->device copies of ITS inputs go to an internal copy stream into staging set
// (ticket & 1), the step runs on the caller's stream once they have landed, its (loss, dt', dbias) are copied to pinned
// host memory behind it. With two staging sets the copies of step n+1 overlap the kernels of step n.
int siglip_host_submit(siglip_ctx* c, const void* img_host, const void* txt_host, float t_prime, float bias,
                       unsigned long long* ticket, void* cuda_stream) {
  if (c == nullptr || img_host == nullptr || txt_host == nullptr || ticket == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = host_entry_init(c))) return rc;
  const size_t chunk_bytes = static_cast<size_t>(c->B) * c->D * sizeof(__nv_bfloat16);
  const unsigned long long n = c->host_submitted;
  const int s = static_cast<int>(n & 1);
  if (n >= 2) {
    // set s was used by step n-2: its kernels must be done before the staging buffers are overwritten, and the
    // caller must have collected its results (siglip_host_wait) before the pinned slot is reused
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_done[s], 0));
    CK(cudaEventSynchronize(c->ev_done[s]));
  }
  float* pin = c->h_pinned + 8 * s;
  pin[0] = t_prime;
  pin[1] = bias;
  float* sc = host_set_scalars(c, s);
  CK(cudaMemcpyAsync(sc, pin, 2 * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaMemcpyAsync(c->h_img[s], img_host, chunk_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaMemcpyAsync(c->h_txt[s], txt_host, chunk_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaEventRecord(c->ev_h2d[s], c->copy_stream));
  CK(cudaStreamWaitEvent(st, c->ev_h2d[s], 0));
  const int saved_bf16 = c->grad_bf16, saved_fmt = c->input_f16;
  c->grad_bf16 = 0;  // the host entries produce fp32 gradients
  c->input_f16 = 0;  // ... from bf16 host buffers
  rc = siglip_fwd_bwd(c, c->h_img[s], c->h_txt[s], sc + 0, sc + 1, sc + 2, c->h_dimg, c->h_dtxt, sc + 3, sc + 4, st);
  c->grad_bf16 = saved_bf16;
  c->input_f16 = saved_fmt;
  if (rc) return rc;
  CK(cudaMemcpyAsync(pin + 4, sc + 2, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(c->ev_done[s], st));
  c->host_submitted = n + 1;
  *ticket = n;
  return 0;
}

int siglip_host_wait(siglip_ctx* c, unsigned long long ticket, float* loss_host, float* dt_prime_host,
                     float* dbias_host) {
  if (c == nullptr || loss_host == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (ticket >= c->host_submitted || ticket + 2 < c->host_submitted)
    return fail(SIGLIP_ERR_STATE, "ticket is not one of the last two submitted steps");
  CK(cudaSetDevice(c->device));
  const int s = static_cast<int>(ticket & 1);
  CK(cudaEventSynchronize(c->ev_done[s]));
  int rc;
  if ((rc = check_dbg(c, "siglip_host_wait"))) return rc;
  const float* pin = c->h_pinned + 8 * s;
  *loss_host = pin[4];
  if (dt_prime_host) *dt_prime_host = pin[5];
  if (dbias_host) *dbias_host = pin[6];
  return 0;
}

int siglip_fwd_bwd_host(siglip_ctx* c, const void* img_host, const void* txt_host, float t_prime, float bias,
                        float* loss_host, float* dimg_host, float* dtxt_host, float* dt_prime_host,
                        float* dbias_host, void* cuda_stream) {
  if (loss_host == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  unsigned long long ticket = 0;
  int rc = siglip_host_submit(c, img_host, txt_host, t_prime, bias, &ticket, cuda_stream);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  if (dimg_host) CK(cudaMemcpyAsync(dimg_host, c->h_dimg, chunk_elems * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (dtxt_host) CK(cudaMemcpyAsync(dtxt_host, c->h_dtxt, chunk_elems * sizeof(float), cudaMemcpyDeviceToHost, st));
  if ((rc = siglip_host_wait(c, ticket, loss_host, dt_prime_host, dbias_host))) return rc;
  if (dimg_host || dtxt_host) CK(cudaStreamSynchronize(st));
  return 0;
}

int siglip_ctx_kernel_times(siglip_ctx* c, double* loss_ms, int* loss_launches, double* grad_ms, int* grad_launches) {
  if (c == nullptr) return fail(SIGLIP_ERR_INVALID, "ctx is null");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  double tl = 0.0, tg = 0.0;
  for (size_t i = 0; i + 1 < c->ev_loss_used; i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev_loss[i], c->ev_loss[i + 1]));
    tl += ms;
  }
  for (size_t i = 0; i + 1 < c->ev_grad_used; i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev_grad[i], c->ev_grad[i + 1]));
    tg += ms;
  }
  if (getenv("SIGLIP_DEBUG_PRINT_TIMES")) {  // per-launch durations of the last W loss / gradient launches
    std::string s = "[kernel times rank " + std::to_string(c->rank) + "] loss:";
    char b[32];
    const size_t nl = c->ev_loss_used / 2, ng = c->ev_grad_used / 2;
    for (size_t i = (nl > (size_t)c->world ? nl - c->world : 0); i < nl; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->ev_loss[2 * i], c->ev_loss[2 * i + 1]);
      snprintf(b, sizeof(b), " %.3f", ms);
      s += b;
    }
    s += " | grad:";
    for (size_t i = (ng > (size_t)c->world ? ng - c->world : 0); i < ng; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->ev_grad[2 * i], c->ev_grad[2 * i + 1]);
      snprintf(b, sizeof(b), " %.3f", ms);
      s += b;
    }
    printf("%s\n", s.c_str());
    fflush(stdout);
  }
  if (loss_ms) *loss_ms = tl;
  if (grad_ms) *grad_ms = tg;
  if (loss_launches) *loss_launches = static_cast<int>(c->ev_loss_used / 2);
  if (grad_launches) *grad_launches = static_cast<int>(c->ev_grad_used / 2);
  c->ev_loss_used = c->ev_grad_used = 0;
  return 0;
}

int siglip_debug_loopback(siglip_ctx* c) {
  if (c == nullptr) return fail(SIGLIP_ERR_INVALID, "ctx is null");
  if (c->world == 1) return fail(SIGLIP_ERR_STATE, "loopback needs world > 1");
  CK(cudaSetDevice(c->device));
  for (int p = 0; p < c->world; ++p) {
    c->peer_txt[p] = c->txt_all;
    c->peer_slots[p] = c->slots;
    c->peer_flags[p] = c->flags;
  }
  c->loopback = true;   // every signal then raises the flag entry of EVERY rank in the local table
  return publish_peer_tables(c);
}

int siglip_debug_set_text_chunk(siglip_ctx* c, int chunk, const void* txt_dev, void* cuda_stream) {
  if (c == nullptr || txt_dev == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (c->world == 1 || chunk < 0 || chunk >= c->world) return fail(SIGLIP_ERR_INVALID, "chunk out of range");
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  CK(cudaMemcpyAsync(c->txt_all + chunk * chunk_elems, txt_dev, chunk_elems * sizeof(__nv_bfloat16),
                     cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

int siglip_debug_get_slot(siglip_ctx* c, int chunk, float* out_dev, void* cuda_stream) {
  if (c == nullptr || out_dev == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (c->world == 1 || chunk < 0 || chunk >= c->world) return fail(SIGLIP_ERR_INVALID, "chunk out of range");
  const size_t chunk_elems = static_cast<size_t>(c->B) * c->D;
  CK(cudaMemcpyAsync(out_dev, c->slots + chunk * chunk_elems, chunk_elems * sizeof(float), cudaMemcpyDeviceToDevice,
                     static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

int siglip_normalize_fwd(siglip_ctx* c, const void* x, int in_bf16, void* xhat_bf16, float* inv_norm,
                         void* cuda_stream) {
  if (c == nullptr || x == nullptr || xhat_bf16 == nullptr || inv_norm == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(xhat_bf16) & 15u))
    return fail(SIGLIP_ERR_INVALID, "buffers must be 16-byte aligned");
  CK(cudaSetDevice(c->device));
  CKI(siglip::launch_normalize_fwd(x, in_bf16, static_cast<__nv_bfloat16*>(xhat_bf16), inv_norm, c->B, c->D,
                                   c->input_f16 ? kXScale : 0.0f, c->num_sms, static_cast<cudaStream_t>(cuda_stream)));
  c->launches++;
  return 0;
}

int siglip_convert_f32(siglip_ctx* c, const float* x_f32, void* out_16bit, void* cuda_stream) {
  if (c == nullptr || x_f32 == nullptr || out_16bit == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(x_f32) & 15u) || (reinterpret_cast<uintptr_t>(out_16bit) & 15u))
    return fail(SIGLIP_ERR_INVALID, "buffers must be 16-byte aligned");
  CK(cudaSetDevice(c->device));
  CKI(siglip::launch_convert_f32(x_f32, out_16bit, static_cast<size_t>(c->B) * c->D, c->input_f16 ? kXScale : 0.0f,
                                 c->num_sms, static_cast<cudaStream_t>(cuda_stream)));
  c->launches++;
  return 0;
}

int siglip_normalize_bwd(siglip_ctx* c, const void* x, int in_bf16, const float* inv_norm, const void* dxhat,
                         int grad_bf16, void* dx, void* cuda_stream) {
  if (c == nullptr || x == nullptr || inv_norm == nullptr || dxhat == nullptr || dx == nullptr)
    return fail(SIGLIP_ERR_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(dxhat) & 15u) ||
      (reinterpret_cast<uintptr_t>(dx) & 15u))
    return fail(SIGLIP_ERR_INVALID, "buffers must be 16-byte aligned");
  CK(cudaSetDevice(c->device));
  CKI(siglip::launch_normalize_bwd(x, in_bf16, inv_norm, dxhat, grad_bf16, dx, c->B, c->D, c->num_sms,
                                   static_cast<cudaStream_t>(cuda_stream)));
  c->launches++;
  return 0;
}

int siglip_scale(siglip_ctx* c, const void* src, void* dst, size_t nbytes, int is_bf16, const float* g,
                 void* cuda_stream) {
  if (c == nullptr || src == nullptr || dst == nullptr || g == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if ((nbytes % 16) != 0 || (reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(dst) & 15u))
    return fail(SIGLIP_ERR_INVALID, "siglip_scale needs 16-byte aligned buffers and a multiple of 16 bytes");
  CK(cudaSetDevice(c->device));
  CKI(siglip::launch_scale(src, dst, is_bf16, g, nbytes, c->num_sms, static_cast<cudaStream_t>(cuda_stream)));
  c->launches++;
  return 0;
}

unsigned long long siglip_ctx_launch_count(const siglip_ctx* c) { return c ? c->launches : 0ull; }

int siglip_debug_gemm(int device, int cta_group, int M, int N, int K, const void* A, long long lda, int a_mn,
                      const void* Bm, long long ldb, int b_mn, float* C, long long ldc, void* cuda_stream) {
  return siglip_debug_gemm_timed(device, cta_group, M, N, K, A, lda, a_mn, Bm, ldb, b_mn, C, ldc, 1, nullptr,
                                 cuda_stream);
}

int siglip_debug_gemm_timed(int device, int cta_group, int M, int N, int K, const void* A, long long lda, int a_mn,
                            const void* Bm, long long ldb, int b_mn, float* C, long long ldc, int iters,
                            float* ms_per_iter, void* cuda_stream) {
  if (A == nullptr || Bm == nullptr || C == nullptr) return fail(SIGLIP_ERR_INVALID, "null argument");
  if (cta_group != 1 && cta_group != 2) return fail(SIGLIP_ERR_INVALID, "cta_group must be 1 or 2");
  if (M < 1 || N < 8 || K < 1 || (N % 8) != 0) return fail(SIGLIP_ERR_INVALID, "need N % 8 == 0");
  if (siglip_device_count() == 0) return fail(SIGLIP_ERR_NO_DEVICE, "no sm_100 device; no CPU fallback");
  CK(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int num_sms = 0;
  CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_operand(&tmA, A, M, K, lda, a_mn, 128))) return rc;
  const char* env_mc = getenv("SIGLIP_DEBUG_MCAST");
  const int mcast = env_mc ? atoi(env_mc) : 1;
  if ((rc = encode_operand(&tmB, Bm, N, K, ldb, b_mn, 256 / (cta_group * mcast)))) return rc;
  float* zero = nullptr;  // t' = 0 -> scale exp(0) * 1 = 1
  CK(cudaMalloc(reinterpret_cast<void**>(&zero), sizeof(float)));
  CK(cudaMemsetAsync(zero, 0, sizeof(float), st));
  DebugRecord* dbg_host = nullptr;
  DebugRecord* dbg_dev = nullptr;
  CK(cudaHostAlloc(reinterpret_cast<void**>(&dbg_host), sizeof(DebugRecord), cudaHostAllocMapped));
  memset(dbg_host, 0, sizeof(DebugRecord));
  CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dbg_dev), dbg_host, 0));
  KernelParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = 1;
  p.prob[0].M = M;
  p.prob[0].N = N;
  p.prob[0].K = K;
  p.prob[0].tiles_m = ceil_div(M, 128 * cta_group);
  p.prob[0].tiles_n = ceil_div(N, 256);
  p.prob[0].a_mn = a_mn ? 1 : 0;
  p.prob[0].b_mn = b_mn ? 1 : 0;
  p.prob[0].ab_f16 = getenv("SIGLIP_DEBUG_AB_F16") ? 1 : 0;
  p.prob[0].acc_scale = 1.0f;
  p.prob[0].out = C;
  p.prob[0].ldo = ldc;
  p.t_prime = zero;
  p.inv_b = 1.0f;
  p.dbg = dbg_dev;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int lrc = 0;
  unsigned long long* wstats = nullptr;
  if (getenv("SIGLIP_DEBUG_WAITSTATS")) {
    printf("[waitstats cg=%d] max co-resident clusters: %d (SMs %d)\n", cta_group,
           siglip::query_max_active_clusters(cta_group), num_sms);
    CK(cudaMalloc(reinterpret_cast<void**>(&wstats), 8 * 256 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(wstats, 0, 8 * 256 * sizeof(unsigned long long), st));
    p.wait_stats = wstats;
  }
  const char* env_sl = getenv("SIGLIP_DEBUG_EPI_SLEEP");
  p.epi_sleep_ns = env_sl ? static_cast<unsigned int>(atoi(env_sl)) : 0u;
  const char* env_st = getenv("SIGLIP_DEBUG_STAGES");
  const int stages = env_st ? atoi(env_st) : 0;
  if (iters > 1)  // warm-up
    lrc = siglip::launch_gemm(cta_group, siglip::kModeOut, stages, mcast, &tmA, &tmB, &tmA, &tmB, &tmA, p, num_sms, st);
  cudaEventRecord(e0, st);
  for (int it = 0; it < iters && lrc == 0; ++it)
    lrc = siglip::launch_gemm(cta_group, siglip::kModeOut, stages, mcast, &tmA, &tmB, &tmA, &tmB, &tmA, p, num_sms, st);
  cudaEventRecord(e1, st);
  cudaError_t se = cudaStreamSynchronize(st);
  if (ms_per_iter != nullptr && se == cudaSuccess && lrc == 0) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_iter = ms / static_cast<float>(iters > 0 ? iters : 1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (wstats != nullptr) {
    std::vector<unsigned long long> h(8 * 256);
    cudaMemcpy(h.data(), wstats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s[6] = {0, 0, 0, 0, 0, 0};
    int nlead = 0, nall = 0;
    for (int b = 0; b < 256; ++b) {
      if (h[8 * b + 0] || h[8 * b + 3]) nall++;
      s[0] += static_cast<double>(h[8 * b + 0]);
      if (h[8 * b + 3]) {
        nlead++;
        for (int j = 1; j < 6; ++j) s[j] += static_cast<double>(h[8 * b + j]);
      }
    }
    const double tot = s[3] > 0 ? s[3] : 1;
    printf("[waitstats cg=%d] producer empty-wait avg %.0f cyc (%d CTAs); MMA thread (%d issuers): loop %.0f cyc = "
           "full-wait %.1f%% + tmem-wait %.1f%% + mma-issue %.1f%% + commit %.1f%% + other %.1f%%\n",
           cta_group, s[0] / (nall ? nall : 1), nall, nlead, s[3] / (nlead ? nlead : 1), 100.0 * s[1] / tot,
           100.0 * s[2] / tot, 100.0 * s[4] / tot, 100.0 * s[5] / tot,
           100.0 * (s[3] - s[1] - s[2] - s[4] - s[5]) / tot);
    fflush(stdout);
    cudaFree(wstats);
  }
  int result = 0;
  if (dbg_host->code != 0) {
    char buf[256];
    snprintf(buf, sizeof(buf), "debug gemm: device wait timed out at site %u (block %u thread %u aux %u %u %u)",
             dbg_host->code, dbg_host->block, dbg_host->thread, dbg_host->aux0, dbg_host->aux1, dbg_host->aux2);
    result = fail(SIGLIP_ERR_CUDA, buf);
  } else if (lrc != 0) {
    result = fail(SIGLIP_ERR_CUDA, std::string("debug gemm launch failed: ") +
                                       cudaGetErrorString(static_cast<cudaError_t>(lrc)));
  } else if (se != cudaSuccess) {
    result = fail(SIGLIP_ERR_CUDA, std::string("debug gemm execution failed: ") + cudaGetErrorString(se));
  }
  cudaFreeHost(dbg_host);
  cudaFree(zero);
  return result;
}

void siglip_ctx_destroy(siglip_ctx* c) {
  if (c == nullptr) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (!c->loopback) {
    for (int p = 0; p < c->world; ++p) {
      if (p == c->rank) continue;
      if (c->peer_txt[p]) cudaIpcCloseMemHandle(c->peer_txt[p]);
      if (c->peer_slots[p]) cudaIpcCloseMemHandle(c->peer_slots[p]);
      if (c->peer_flags[p]) cudaIpcCloseMemHandle(c->peer_flags[p]);
    }
  }
  cudaFree(c->txt_all);
  for (int k = 0; k < kMaxWorld; ++k) cudaFree(c->G[k]);
  cudaFree(c->g_diag);
  cudaFree(c->img16);
  cudaFree(c->txt16);
  cudaFree(c->slots);
  cudaFree(c->dimg_acc);
  cudaFree(c->dtxt_acc);
  cudaFree(c->final_ptrs_dev);
  cudaFree(c->partials);
  cudaFree(c->fin_counter);
  cudaFree(c->flags);
  cudaFree(c->scalars);
  cudaFree(c->reduce_ptrs_dev);
  cudaFree(c->signal_ptrs_dev);
  cudaFree(c->mailbox_ptrs_dev);
  for (int s = 0; s < 2; ++s) {
    cudaFree(c->h_img[s]);
    cudaFree(c->h_txt[s]);
    if (c->ev_h2d[s]) cudaEventDestroy(c->ev_h2d[s]);
    if (c->ev_done[s]) cudaEventDestroy(c->ev_done[s]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  cudaFree(c->h_dimg);
  cudaFree(c->h_dtxt);
  if (c->dbg_host) cudaFreeHost(c->dbg_host);
  for (cudaEvent_t e : c->ev_loss) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_grad) cudaEventDestroy(e);
  cudaGetLastError();
  delete c;
}

}  // extern "C"
