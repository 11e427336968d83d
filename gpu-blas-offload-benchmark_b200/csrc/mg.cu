// mg.cu -- multi-GPU teams: ONE GEMM / GEMV partitioned over the GPUs of an NVSwitch box
// (north star: "GEMM shards into column slabs of B and C, A replicated over
// NVLink, C gathered back; GEMV shards by row blocks of A, x broadcast").
//
// The reference has no multi-GPU notion: its backend reads one device in
// gemm_gpu::initialise (cuBLAS/gemm.hh:45-63, cuBLAS/gemv.hh:45-60).  This file is what sits
// behind that slot when B200_NGPUS > 1 (B200/gemm.hh, B200/gemv.hh), and behind
// bench.py --gpus N (one process per GPU).
//
// Team = `world` members, one per GPU.  Two ways to own them:
//   local  one process drives every member (the harness: b200_mg_create);
//   rank   one process per member (b200_mg_create_rank); peers' buffers and flag pages are
//          mapped through CUDA IPC handles the caller exchanges (torch.distributed, MPI, ...).
// Both run the SAME protocol, below; only the way a peer address is obtained differs.
//
// Partitioning (column-major, nothing is repacked):
//   GEMM  member g owns a contiguous column slab of B and C; A is needed everywhere.  Member g
//         holds a "K-share" of A (a contiguous block of A's columns): it uploads only that
//         share over its own PCIe link and PUSHES it to every peer's replica over NVLink with
//         copy-engine copies (no SM is taken from the GEMM), sub-chunk by sub-chunk.  The
//         consumer computes  C_slab (+)= A(:, Kc) * B_slab(Kc, :)  chunk by chunk in arrival
//         order, so the all-gather of A hides behind the tensor-core work tile by tile.
//   GEMV  member g owns a compact row block of A and y; x is replicated the same way.
//
// Cross-GPU ordering uses flag words in each member's own device memory:
//   producer: data copy, then a 4-byte copy of the op's epoch into the consumer's flag -- same
//             stream, so the flag lands after the data (works across processes, needs no SM);
//   consumer: cuStreamWaitValue32(flag >= epoch) on its compute stream (a one-thread spin kernel
//             when stream memory operations are unavailable: B200_MG_WAIT=kernel).
// Write-after-read on a replica is closed by an "entered" flag: a member announces an op only
// after its previous kernels have finished reading the replica, and nobody pushes to it before.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace b200;

namespace {

constexpr int kMaxTeam = 16;
constexpr int kFlagWords = 1024;
constexpr int F_ENTER = 0;     // + member: that member has entered op <epoch>
constexpr int F_READY = 64;    // + chunk : that chunk of the replicated operand has landed here
constexpr int F_GATHER = 576;  // + member: that member's result slab has landed here
constexpr int kMaxChunks = F_GATHER - F_READY;
constexpr int kValTable = 1 << 16;
constexpr int kMaxSub = 8, kMaxPieces = 8;

typedef CUresult (*wait_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct MgMember {
  b200_ctx* ctx = nullptr;      // local members only
  uint32_t* flags = nullptr;    // this member's flag page, at an address valid in this process
  bool flags_imported = false;  // mapped from an IPC handle
  uint32_t* vals = nullptr;     // local: vals[i] = vals_base + i, the source of flag copies
  uint32_t vals_base = 0;
  cudaStream_t push = nullptr;  // local: NVLink pushes (copy engines)
  cudaEvent_t push_done = nullptr;
  bool push_dirty = false;
  uint32_t ops = 0;             // team operations issued so far = the epoch of the last one
};

// One host thread per member of a LOCAL team.  A member's step is issued asynchronously, but an
// issue can still block on its own device (a lazily loaded module, scratch growth, a full launch
// queue) while that device waits for a flag only a peer's step will raise -- so the members of a
// team must never be issued one after the other from a single thread.
struct MgWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = 0;
  char err[512] = {};
};

}  // namespace

struct b200_mg {
  int world = 1;
  int self = -1;  // rank mode: the member this process owns; local mode: -1
  MgMember mem[kMaxTeam];
  int wait_kernel = 0;  // 1: spin kernel instead of cuStreamWaitValue32
  wait_value32_fn wait_value32 = nullptr;
  int opt_sub = 0, opt_pieces = 0;  // tuning overrides (B200_MG_SUB, B200_MG_PIECES)
  size_t opt_chunk_bytes = (size_t)256 << 20;
  MgWorker* workers = nullptr;  // local teams with more than one member
};

struct b200_mg_problem {
  b200_mg* mg = nullptr;
  int kind = 0, elem = 8, mode = 0, m = 0, n = 0, k = 0, active = 1;
  void* h[3] = {};   // host-visible operands the harness fills / reads
  void* d0[3] = {};  // member 0's device twins (b200_alloc)
  size_t bytes[3] = {};
  // per member
  void* rep[kMaxTeam] = {};   // GEMM: replica of A;  GEMV: replica of x
  void* slab[kMaxTeam] = {};  // GEMM: column slab of B;  GEMV: compact row block of A
  void* out[kMaxTeam] = {};   // GEMM: column slab of C;  GEMV: row block of y
  long long start[kMaxTeam] = {}, count[kMaxTeam] = {};  // columns of B/C, or rows of A/y
  long long share0[kMaxTeam] = {};                        // first column of A / element of x in the member's share
};

namespace {

__global__ void mg_wait_kernel(const volatile uint32_t* flag, uint32_t value) {
  while ((int32_t)(*flag - value) < 0) __nanosleep(100);
  __threadfence_system();
}

// Contiguous split of `total` into `parts`; boundaries are multiples of `align` when every
// part can still get at least 2 * align (kernels like whole k-tiles / 128-column tiles).
void split(long long total, int parts, int idx, int align, long long* start, long long* count) {
  if (align < 1 || total < 2LL * align * parts) align = 1;
  auto bound = [&](int p) -> long long {
    if (p <= 0) return 0;
    if (p >= parts) return total;
    return total * p / parts / align * align;
  };
  *start = bound(idx);
  *count = bound(idx + 1) - *start;
}

inline size_t extent(size_t elem, long long rows, long long cols, long long ld) {
  return rows > 0 && cols > 0 ? elem * ((size_t)ld * (cols - 1) + rows) : 0;
}

int fill_vals(MgMember& me, uint32_t base) {
  std::vector<uint32_t> host(kValTable);
  for (int i = 0; i < kValTable; i++) host[i] = base + (uint32_t)i;
  B200_CUDA(cudaMemcpy(me.vals, host.data(), sizeof(uint32_t) * kValTable, cudaMemcpyHostToDevice));
  me.vals_base = base;
  return 0;
}

int setup_member(b200_mg* mg, int g, int device) {
  MgMember& me = mg->mem[g];
  if (b200_ctx_create(&me.ctx, device)) return 1;
  B200_CUDA(cudaSetDevice(device));  // the flag page and the push stream live on this member's GPU
  B200_CUDA(cudaMalloc((void**)&me.flags, sizeof(uint32_t) * kFlagWords));
  B200_CUDA(cudaMemset(me.flags, 0, sizeof(uint32_t) * kFlagWords));
  B200_CUDA(cudaMalloc((void**)&me.vals, sizeof(uint32_t) * kValTable));
  if (fill_vals(me, 1)) return 1;
  B200_CUDA(cudaStreamCreateWithFlags(&me.push, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&me.push_done, cudaEventDisableTiming));
  B200_CUDA(cudaDeviceSynchronize());
  return 0;
}

int common_init(b200_mg* mg) {
  if (const char* e = getenv("B200_MG_WAIT")) mg->wait_kernel = !strcmp(e, "kernel");
  if (const char* e = getenv("B200_MG_SUB")) mg->opt_sub = std::max(0, std::min(kMaxSub, atoi(e)));
  if (const char* e = getenv("B200_MG_PIECES")) mg->opt_pieces = std::max(0, std::min(kMaxPieces, atoi(e)));
  if (const char* e = getenv("B200_MG_CHUNK_MB")) mg->opt_chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;
  if (!mg->wait_kernel) {
    // the driver entry point is looked up at run time: the library links no libcuda
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess && fn)
      mg->wait_value32 = (wait_value32_fn)fn;
    else {
      cudaGetLastError();
      mg->wait_kernel = 1;
    }
  }
  return 0;
}

// flag `word` of member `to` <- epoch, ordered after everything already on `st`
int mg_signal(b200_mg* mg, int from, cudaStream_t st, int to, int word, uint32_t epoch) {
  MgMember& me = mg->mem[from];
  B200_REQUIRE(mg->mem[to].flags, "b200_mg: the flag page of member %d has not been imported", to);
  B200_CUDA(cudaMemcpyAsync(mg->mem[to].flags + word, me.vals + (epoch - me.vals_base), sizeof(uint32_t),
                            cudaMemcpyDefault, st));
  return 0;
}

// `st` (a stream of member g's device) waits until g's own flag `word` reaches epoch
int mg_wait(b200_mg* mg, int g, cudaStream_t st, int word, uint32_t epoch) {
  uint32_t* addr = mg->mem[g].flags + word;
  if (!mg->wait_kernel) {
    const CUresult r = mg->wait_value32((CUstream)st, (CUdeviceptr)(uintptr_t)addr, epoch, CU_STREAM_WAIT_VALUE_GEQ);
    B200_REQUIRE(r == CUDA_SUCCESS, "b200_mg: cuStreamWaitValue32 failed (%d); set B200_MG_WAIT=kernel", (int)r);
    return 0;
  }
  mg_wait_kernel<<<1, 1, 0, st>>>(addr, epoch);
  B200_CUDA(cudaGetLastError());
  return 0;
}

// Start a team op on local member g: device selected, epoch taken, flag-value table refreshed.
int begin_op(b200_mg* mg, int g, uint32_t* epoch) {
  B200_REQUIRE(mg && g >= 0 && g < mg->world && mg->mem[g].ctx, "b200_mg: member %d is not local to this process", g);
  MgMember& me = mg->mem[g];
  B200_CUDA(cudaSetDevice(me.ctx->device));
  *epoch = ++me.ops;
  if (*epoch - me.vals_base >= (uint32_t)kValTable) {
    B200_CUDA(cudaStreamSynchronize(me.push));
    if (fill_vals(me, *epoch)) return 1;
  }
  return 0;
}

enum { PH_UPLOAD = B200_MG_UPLOAD, PH_REPLICATE = B200_MG_REPLICATE, PH_COMPUTE = B200_MG_COMPUTE,
       PH_DOWNLOAD = B200_MG_DOWNLOAD, PH_GATHER = B200_MG_GATHER };

int call_gemm(b200_ctx* ctx, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
              double beta, double* C, int ldc) {
  return b200_dgemm(ctx, 0, 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
int call_gemm(b200_ctx* ctx, int m, int n, int k, double alpha, const float* A, int lda, const float* B, int ldb,
              double beta, float* C, int ldc) {
  return b200_sgemm(ctx, 0, 0, m, n, k, (float)alpha, A, lda, B, ldb, (float)beta, C, ldc);
}
int call_gemv(b200_ctx* ctx, int m, int n, double alpha, const double* A, int lda, const double* x, double beta,
              double* y) {
  return b200_dgemv(ctx, 0, m, n, alpha, A, lda, x, 1, beta, y, 1);
}
int call_gemv(b200_ctx* ctx, int m, int n, double alpha, const float* A, int lda, const float* x, double beta,
              float* y) {
  return b200_sgemv(ctx, 0, m, n, (float)alpha, A, lda, x, 1, (float)beta, y, 1);
}

// Sub-chunks per K-share: pieces of ~256 MiB (B200_MG_CHUNK_MB) so the first block starts after
// a few ms of upload; the same on every member (it indexes the flag words).
int share_sub(const b200_mg* mg, size_t elem, long long rows, long long k, int W) {
  if (mg->opt_sub) return mg->opt_sub;
  const size_t share_bytes = elem * (size_t)rows * (size_t)((k + W - 1) / W);
  const long long s = (long long)((share_bytes + mg->opt_chunk_bytes - 1) / mg->opt_chunk_bytes);
  return (int)std::max<long long>(1, std::min<long long>(4, s));
}

struct Chunk { int owner; long long k0, kk; };

template <typename T>
int gemm_step(b200_mg* mg, int g, const b200_mg_gemm_args& a) {
  ProfileRange pr("mg gemm step member %d: %dx%dx%d phases 0x%x", g, a.m, a.n_slab, a.k, a.phases);
  uint32_t epoch = 0;
  if (begin_op(mg, g, &epoch)) return 1;
  MgMember& me = mg->mem[g];
  b200_ctx* ctx = me.ctx;
  const int W = a.active;
  B200_REQUIRE(W >= 1 && W <= mg->world, "b200_mg_gemm: active = %d outside [1, %d]", W, mg->world);
  if (g >= W) return 0;  // this member sits the problem out (too small to spread further)
  const int ph = a.phases;
  const bool upl = (ph & PH_UPLOAD) != 0, rep = (ph & PH_REPLICATE) && W > 1, comp = (ph & PH_COMPUTE) != 0;
  const bool down = (ph & PH_DOWNLOAD) && a.hC, gather = (ph & PH_GATHER) && W > 1;
  const size_t s = sizeof(T);
  const int m = a.m, n = a.n_slab, k = a.k, lda = a.lda, ldb = a.ldb, ldc = a.ldc;
  B200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "b200_mg_gemm: negative dimension");
  B200_REQUIRE(a.A_rep && a.A_rep[g] && a.dB && a.dC, "b200_mg_gemm: null device buffer");
  T* A = (T*)a.A_rep[g];
  T* B = (T*)a.dB;
  T* C = (T*)a.dC;
  const T* hA = (const T*)a.hA_share;
  const T* hB = (const T*)a.hB;
  T* hC = (T*)a.hC;
  cudaStream_t up = ctx->lane[0], dn = ctx->lane[2], push = me.push, cs = ctx->compute;
  cudaEvent_t ev;

  if (engine_begin(ctx)) return 1;
  if (me.push_dirty) {  // earlier pushes still read buffers this op may overwrite
    B200_CUDA(cudaStreamWaitEvent(cs, me.push_done, 0));
    B200_CUDA(cudaStreamWaitEvent(up, me.push_done, 0));
    me.push_dirty = false;
  }

  // ---- chunk table: every member's K-share of A, cut into `sub` sub-chunks
  const int sub = share_sub(mg, s, lda, k, W);
  B200_REQUIRE(W * sub <= kMaxChunks, "b200_mg_gemm: too many chunks");
  std::vector<Chunk> chunk((size_t)W * sub);
  long long my_k0 = 0;
  for (int p = 0; p < W; p++) {
    long long s0, sk;
    split(k, W, p, 32, &s0, &sk);
    if (p == g) my_k0 = s0;
    for (int i = 0; i < sub; i++) {
      long long c0, ck;
      split(sk, sub, i, 32, &c0, &ck);
      chunk[(size_t)p * sub + i] = {p, s0 + c0, ck};
    }
  }
  // ---- column pieces of my slab (uploads, downloads and gathers overlap compute piece by piece)
  int ns = 1;
  if (upl || down || gather) ns = mg->opt_pieces ? mg->opt_pieces : (int)std::max(1, std::min(kMaxPieces, n / 512));
  // pre-split SGEMM (sgemm_fused = off) splits its operands into TF32 hi/lo first, a pass over all
  // of A per piece: few pieces then, or that pass costs more than the overlap returns.  The fused
  // kernels (the default) have no such pass and take the same pieces as DGEMM.
  if (sizeof(T) == 4 && !mg->opt_pieces && ctx->opt_sgemm_fused == 2) ns = (upl || down) ? std::min(ns, 2) : 1;
  long long pj0[kMaxPieces], pnc[kMaxPieces];
  for (int j = 0; j < ns; j++) split(n, ns, j, 128, &pj0[j], &pnc[j]);

  // ---- announce the op: my earlier kernels have stopped reading my replica of A
  if (rep) {
    if (record_on(ctx, cs, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(push, ev, 0));
    for (int d = 1; d < W; d++)
      if (mg_signal(mg, g, push, (g + d) % W, F_ENTER + g, epoch)) return 1;
  }
  const bool upload_c = upl && hC && (a.beta != 0.0 || ctx->opt_upload_c);
  if (upload_c && m > 0 && n > 0) {
    B200_CUDA(cudaMemcpyAsync(C, hC, extent(s, m, n, ldc), cudaMemcpyHostToDevice, up));
    if (record_on(ctx, up, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(cs, ev, 0));
  }

  // ---- uploads (own K-share in sub-chunks, interleaved with the pieces of B) and pushes
  cudaEvent_t evA[kMaxSub] = {}, evB[kMaxPieces] = {};
  bool haveA[kMaxSub] = {}, haveB[kMaxPieces] = {};
  bool entered = false, pushed = false;
  for (int i = 0; i < std::max(sub, ns); i++) {
    if (i < sub) {
      const Chunk& c = chunk[(size_t)g * sub + i];
      if (c.kk > 0 && m > 0) {
        if (upl && hA) {
          B200_CUDA(cudaMemcpyAsync(A + c.k0 * lda, hA + (c.k0 - my_k0) * lda, extent(s, m, c.kk, lda),
                                    cudaMemcpyHostToDevice, up));
          if (record_on(ctx, up, &evA[i])) return 1;
          haveA[i] = true;
        }
        if (rep) {
          if (haveA[i]) B200_CUDA(cudaStreamWaitEvent(push, evA[i], 0));
          if (!entered) {  // nobody's replica is written before its owner has entered this op
            for (int d = 1; d < W; d++)
              if (mg_wait(mg, g, push, F_ENTER + (g + d) % W, epoch)) return 1;
            entered = true;
          }
          for (int d = 1; d < W; d++) {
            const int q = (g + d) % W;
            B200_REQUIRE(a.A_rep[q], "b200_mg_gemm: replica of member %d is not mapped", q);
            B200_CUDA(cudaMemcpyAsync((T*)a.A_rep[q] + c.k0 * lda, A + c.k0 * lda, extent(s, m, c.kk, lda),
                                      cudaMemcpyDefault, push));
            if (mg_signal(mg, g, push, q, F_READY + g * sub + i, epoch)) return 1;
          }
          pushed = true;
        }
      }
    }
    if (i < ns && upl && hB && pnc[i] > 0 && k > 0) {
      B200_CUDA(cudaMemcpyAsync(B + pj0[i] * ldb, hB + pj0[i] * ldb, extent(s, k, pnc[i], ldb),
                                cudaMemcpyHostToDevice, up));
      if (record_on(ctx, up, &evB[i])) return 1;
      haveB[i] = true;
    }
  }

  // Arrival order of the chunks at member g: sub-chunk i of the member d places behind it
  // (own first); producers push in the matching order, so at any time a member receives
  // from one peer.
  auto wait_chunk = [&](int p, int i) -> int {
    if (p == g) {
      if (haveA[i]) B200_CUDA(cudaStreamWaitEvent(cs, evA[i], 0));
      return 0;
    }
    return rep ? mg_wait(mg, g, cs, F_READY + p * sub + i, epoch) : 0;
  };

  if (comp && m > 0 && n > 0) {
    for (int j = 0; j < ns; j++) {
      const long long j0 = pj0[j], nc = pnc[j];
      if (nc <= 0) continue;
      if (haveB[j]) B200_CUDA(cudaStreamWaitEvent(cs, evB[j], 0));
      bool first = true;
      if (j == 0 && (rep || upl) && k > 0) {
        for (int i = 0; i < sub; i++)
          for (int d = 0; d < W; d++) {
            const int p = (g - d + W) % W;
            const Chunk& c = chunk[(size_t)p * sub + i];
            if (c.kk <= 0) continue;
            if (wait_chunk(p, i)) return 1;
            if (call_gemm(ctx, m, (int)nc, (int)c.kk, a.alpha, A + c.k0 * lda, lda, B + c.k0 + j0 * ldb, ldb,
                          first ? a.beta : 1.0, C + j0 * ldc, ldc))
              return 1;
            first = false;
          }
      }
      if (first)  // every chunk of A is here (or k == 0): one GEMM over the whole K range
        if (call_gemm(ctx, m, (int)nc, k, a.alpha, A, lda, B + j0 * ldb, ldb, a.beta, C + j0 * ldc, ldc)) return 1;
      if (down || (gather && a.C_gather && a.C_gather != (void*)C)) {
        if (record_on(ctx, cs, &ev)) return 1;
        if (down) {
          B200_CUDA(cudaStreamWaitEvent(dn, ev, 0));
          if (ldc == m || nc == 1)
            B200_CUDA(cudaMemcpyAsync(hC + j0 * ldc, C + j0 * ldc, extent(s, m, nc, ldc), cudaMemcpyDeviceToHost, dn));
          else
            B200_CUDA(cudaMemcpy2DAsync(hC + j0 * ldc, s * ldc, C + j0 * ldc, s * ldc, s * m, nc,
                                        cudaMemcpyDeviceToHost, dn));
        }
        if (gather && a.C_gather && a.C_gather != (void*)C) {
          B200_CUDA(cudaStreamWaitEvent(push, ev, 0));
          B200_CUDA(cudaMemcpyAsync((T*)a.C_gather + j0 * ldc, C + j0 * ldc, extent(s, m, nc, ldc), cudaMemcpyDefault,
                                    push));
          pushed = true;
        }
      }
    }
  } else if (!comp) {
    // no kernel in this op (Once mode's preLoop / postLoop): whatever was uploaded or
    // replicated must be in place before the kernels of later ops
    for (int i = 0; i < sub; i++)
      for (int d = 0; d < W; d++) {
        const int p = (g - d + W) % W;
        if (chunk[(size_t)p * sub + i].kk > 0 && m > 0 && wait_chunk(p, i)) return 1;
      }
    for (int j = 0; j < ns; j++)
      if (haveB[j]) B200_CUDA(cudaStreamWaitEvent(cs, evB[j], 0));
    if (down && m > 0 && n > 0) {  // engine_begin ordered the lanes after the kernels
      if (ldc == m || n == 1)
        B200_CUDA(cudaMemcpyAsync(hC, C, extent(s, m, n, ldc), cudaMemcpyDeviceToHost, dn));
      else
        B200_CUDA(cudaMemcpy2DAsync(hC, s * ldc, C, s * ldc, s * m, n, cudaMemcpyDeviceToHost, dn));
    }
  }
  if (gather) {
    if (g != 0) {
      if (mg_signal(mg, g, push, 0, F_GATHER + g, epoch)) return 1;
      pushed = true;
    } else {
      for (int p = 1; p < W; p++)
        if (mg_wait(mg, 0, cs, F_GATHER + p, epoch)) return 1;
    }
  }
  if (pushed || rep) {
    B200_CUDA(cudaEventRecord(me.push_done, push));
    me.push_dirty = true;
  }
  ctx->lane_dirty[0] = ctx->lane_dirty[2] = true;
  ctx->compute_dirty = true;
  return 0;
}

template <typename T>
int gemv_step(b200_mg* mg, int g, const b200_mg_gemv_args& a) {
  ProfileRange pr("mg gemv step member %d: %dx%d phases 0x%x", g, a.m_blk, a.n, a.phases);
  uint32_t epoch = 0;
  if (begin_op(mg, g, &epoch)) return 1;
  MgMember& me = mg->mem[g];
  b200_ctx* ctx = me.ctx;
  const int W = a.active;
  B200_REQUIRE(W >= 1 && W <= mg->world, "b200_mg_gemv: active = %d outside [1, %d]", W, mg->world);
  if (g >= W) return 0;
  const int ph = a.phases;
  const bool upl = (ph & PH_UPLOAD) != 0, rep = (ph & PH_REPLICATE) && W > 1, comp = (ph & PH_COMPUTE) != 0;
  const bool down = (ph & PH_DOWNLOAD) && a.hy, gather = (ph & PH_GATHER) && W > 1;
  const size_t s = sizeof(T);
  const int m = a.m_blk, n = a.n, lda = a.lda;
  B200_REQUIRE(m >= 0 && n >= 0 && lda >= std::max(1, m), "b200_mg_gemv: bad dimensions");
  B200_REQUIRE(a.x_rep && a.x_rep[g] && a.dA && a.dy, "b200_mg_gemv: null device buffer");
  T* A = (T*)a.dA;
  T* x = (T*)a.x_rep[g];
  T* y = (T*)a.dy;
  const T* hA = (const T*)a.hA;
  const T* hx = (const T*)a.hx_share;
  T* hy = (T*)a.hy;
  cudaStream_t side = ctx->lane[2], push = me.push, cs = ctx->compute;
  cudaEvent_t ev;

  if (engine_begin(ctx)) return 1;
  if (me.push_dirty) {
    B200_CUDA(cudaStreamWaitEvent(cs, me.push_done, 0));
    B200_CUDA(cudaStreamWaitEvent(side, me.push_done, 0));
    me.push_dirty = false;
  }
  long long x0[kMaxTeam], xn[kMaxTeam];
  for (int p = 0; p < W; p++) split(n, W, p, 1, &x0[p], &xn[p]);
  bool pushed = false;

  if (rep) {
    if (record_on(ctx, cs, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(push, ev, 0));
    for (int d = 1; d < W; d++)
      if (mg_signal(mg, g, push, (g + d) % W, F_ENTER + g, epoch)) return 1;
  }
  // ---- x: my share up over PCIe, then to every peer over NVLink
  bool have_x = false;
  cudaEvent_t ev_x = nullptr;
  if (upl && hx && xn[g] > 0) {
    B200_CUDA(cudaMemcpyAsync(x + x0[g], hx, s * (size_t)xn[g], cudaMemcpyHostToDevice, side));
    if (record_on(ctx, side, &ev_x)) return 1;
    have_x = true;
  }
  if (upl && hy && m > 0 && (a.beta != 0.0 || ctx->opt_upload_c)) {
    B200_CUDA(cudaMemcpyAsync(y, hy, s * (size_t)m, cudaMemcpyHostToDevice, side));
    if (record_on(ctx, side, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(cs, ev, 0));
  }
  if (rep && xn[g] > 0) {
    if (have_x) B200_CUDA(cudaStreamWaitEvent(push, ev_x, 0));
    for (int d = 1; d < W; d++)
      if (mg_wait(mg, g, push, F_ENTER + (g + d) % W, epoch)) return 1;
    for (int d = 1; d < W; d++) {
      const int q = (g + d) % W;
      B200_REQUIRE(a.x_rep[q], "b200_mg_gemv: replica of member %d is not mapped", q);
      B200_CUDA(cudaMemcpyAsync((T*)a.x_rep[q] + x0[g], x + x0[g], s * (size_t)xn[g], cudaMemcpyDefault, push));
      if (mg_signal(mg, g, push, q, F_READY + g, epoch)) return 1;
    }
    pushed = true;
  }
  // the compute stream needs all of x
  if (have_x) B200_CUDA(cudaStreamWaitEvent(cs, ev_x, 0));
  if (rep)
    for (int d = 1; d < W; d++) {
      const int p = (g - d + W) % W;
      if (xn[p] > 0 && mg_wait(mg, g, cs, F_READY + p, epoch)) return 1;
    }

  // ---- A: my row block, streamed up in column slabs (kernel j hides behind upload j + 1)
  auto upload_cols = [&](long long j0, long long nc, cudaStream_t lane) -> int {
    if (a.h_lda == (long long)lda || nc == 1 || m == 0)
      B200_CUDA(cudaMemcpyAsync(A + j0 * lda, hA + j0 * a.h_lda, extent(s, m, nc, lda), cudaMemcpyHostToDevice, lane));
    else  // a row block of a wider column-major matrix: rows of m elements, pitch h_lda
      B200_CUDA(cudaMemcpy2DAsync(A + j0 * lda, s * lda, hA + j0 * a.h_lda, s * (size_t)a.h_lda, s * m, nc,
                                  cudaMemcpyHostToDevice, lane));
    return 0;
  };
  const size_t bytesA = extent(s, m, n, lda);
  if (upl && hA && m > 0 && n > 0) {
    const int slabs = bytesA < ((size_t)8 << 20) ? 1
                      : (int)std::max<size_t>(2, std::min<size_t>(16, bytesA / ((size_t)16 << 20)));
    const long long cols = (n + slabs - 1) / slabs;
    int i = 0;
    for (long long j0 = 0; j0 < n; j0 += cols, i++) {
      const long long nc = std::min<long long>(cols, n - j0);
      cudaStream_t lane = ctx->lane[i & 1];
      if (upload_cols(j0, nc, lane)) return 1;
      if (record_on(ctx, lane, &ev)) return 1;
      B200_CUDA(cudaStreamWaitEvent(cs, ev, 0));
      if (comp && call_gemv(ctx, m, (int)nc, a.alpha, A + j0 * lda, lda, x + j0, i == 0 ? a.beta : 1.0, y)) return 1;
    }
    ctx->lane_dirty[0] = ctx->lane_dirty[1] = true;
  } else if (comp && m > 0) {
    if (call_gemv(ctx, m, n, a.alpha, A, lda, x, a.beta, y)) return 1;
  }
  if ((down || gather) && m > 0) {
    if (comp) {
      if (record_on(ctx, cs, &ev)) return 1;
      B200_CUDA(cudaStreamWaitEvent(side, ev, 0));
      B200_CUDA(cudaStreamWaitEvent(push, ev, 0));
    }
    if (down) B200_CUDA(cudaMemcpyAsync(hy, y, s * (size_t)m, cudaMemcpyDeviceToHost, side));
    if (gather && a.y_gather && a.y_gather != (void*)y) {
      B200_CUDA(cudaMemcpyAsync(a.y_gather, y, s * (size_t)m, cudaMemcpyDefault, push));
      pushed = true;
    }
  }
  if (gather) {
    if (g != 0) {
      if (mg_signal(mg, g, push, 0, F_GATHER + g, epoch)) return 1;
      pushed = true;
    } else {
      for (int p = 1; p < W; p++)
        if (mg_wait(mg, 0, cs, F_GATHER + p, epoch)) return 1;
    }
  }
  if (pushed || rep) {
    B200_CUDA(cudaEventRecord(me.push_done, push));
    me.push_dirty = true;
  }
  ctx->lane_dirty[2] = true;
  ctx->compute_dirty = true;
  return 0;
}

void worker_main(MgWorker* w, int device) {
  cudaSetDevice(device);  // this thread only ever drives one member
  std::unique_lock<std::mutex> lk(w->mu);
  for (;;) {
    w->cv.wait(lk, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    w->rc = w->job();
    if (w->rc) snprintf(w->err, sizeof(w->err), "%s", b200_last_error());
    w->has_job = false;
    w->done = true;
    w->cv.notify_all();
  }
}

void start_workers(b200_mg* mg) {
  if (mg->self >= 0 || mg->world < 2) return;
  mg->workers = new MgWorker[mg->world];
  for (int g = 0; g < mg->world; g++)
    mg->workers[g].th = std::thread(worker_main, &mg->workers[g], mg->mem[g].ctx ? mg->mem[g].ctx->device : 0);
}

void stop_workers(b200_mg* mg) {
  if (!mg->workers) return;
  for (int g = 0; g < mg->world; g++) {
    {
      std::lock_guard<std::mutex> lk(mg->workers[g].mu);
      mg->workers[g].quit = true;
    }
    mg->workers[g].cv.notify_all();
    if (mg->workers[g].th.joinable()) mg->workers[g].th.join();
  }
  delete[] mg->workers;
  mg->workers = nullptr;
}

// fn(g) for every member, each from its own thread; returns the first failure
int run_members(b200_mg* mg, int active, const std::function<int(int)>& fn) {
  if (!mg->workers || active <= 1) {  // nobody waits for a peer: plain calls, no thread hand-off
    for (int g = 0; g < mg->world; g++)
      if (mg->mem[g].ctx)
        if (int rc = fn(g)) return rc;
    return 0;
  }
  for (int g = 0; g < mg->world; g++) {
    MgWorker& w = mg->workers[g];
    {
      std::lock_guard<std::mutex> lk(w.mu);
      w.job = [&fn, g] { return fn(g); };
      w.done = false;
      w.has_job = true;
    }
    w.cv.notify_all();
  }
  int rc = 0;
  for (int g = 0; g < mg->world; g++) {
    MgWorker& w = mg->workers[g];
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) {
      rc = w.rc;
      set_error("%s", w.err);
    }
  }
  return rc;
}

int sync_member(MgMember& me) {
  if (!me.ctx) return 0;
  B200_CUDA(cudaStreamSynchronize(me.push));
  me.push_dirty = false;
  return b200_sync(me.ctx);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ partitioning (pure host)

int b200_mg_partition(long long total, int parts, int index, int align, long long* start, long long* count) {
  B200_REQUIRE(total >= 0 && parts >= 1 && index >= 0 && index < parts && start && count,
               "b200_mg_partition: bad argument (total=%lld parts=%d index=%d)", total, parts, index);
  split(total, parts, index, align, start, count);
  return 0;
}

int b200_mg_active_gemm(int world, int elem, int m, int n, int k) {
  // spread over as many GPUs as still leaves each one >= 256 columns and >= ~4 GFLOP (about
  // 100 us of FP64 tensor work): below that the handshakes and extra launches cost more than
  // the parallelism returns, and the single-GPU paths (zero copy, one shot, ...) are better
  (void)elem;
  if (world <= 1 || m <= 0 || n <= 0 || k <= 0) return 1;
  const double flops = 2.0 * m * n * (double)k;
  long long a = std::min<long long>(world, std::min<long long>(n / 256, (long long)(flops / 4e9)));
  return (int)std::max<long long>(1, a);
}

int b200_mg_active_gemv(int world, int elem, int m, int n) {
  // >= 1024 rows and >= 64 MiB of A per GPU (~10 us of HBM streaming)
  if (world <= 1 || m <= 0 || n <= 0) return 1;
  const double bytes = (double)elem * m * n;
  long long a = std::min<long long>(world, std::min<long long>(m / 1024, (long long)(bytes / (double)((size_t)64 << 20))));
  return (int)std::max<long long>(1, a);
}

// ------------------------------------------------------------------ teams

int b200_mg_create(b200_mg** out, int ngpus) {
  DeviceRestore dr;
  B200_REQUIRE(out, "b200_mg_create: out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("b200_mg_create: no CUDA device (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return 1;
  }
  B200_REQUIRE(ngpus >= 1 && ngpus <= kMaxTeam, "b200_mg_create: ngpus = %d outside [1, %d]", ngpus, kMaxTeam);
  B200_REQUIRE(ngpus <= count, "b200_mg_create: %d GPUs requested, %d visible", ngpus, count);
  b200_mg* mg = new b200_mg();
  mg->world = ngpus;
  if (common_init(mg)) { b200_mg_destroy(mg); return 1; }
  for (int g = 0; g < ngpus; g++)
    if (setup_member(mg, g, g)) { b200_mg_destroy(mg); return 1; }
  for (int g = 0; g < ngpus; g++) {
    cudaSetDevice(g);
    for (int p = 0; p < ngpus; p++) {
      if (p == g) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, g, p);
      if (!can) {
        set_error("b200_mg_create: GPU %d cannot access GPU %d (no NVLink / P2P path)", g, p);
        b200_mg_destroy(mg);
        return 1;
      }
      e = cudaDeviceEnablePeerAccess(p, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("b200_mg_create: cudaDeviceEnablePeerAccess(%d -> %d): %s", g, p, cudaGetErrorString(e));
        b200_mg_destroy(mg);
        return 1;
      }
      cudaGetLastError();
    }
  }
  start_workers(mg);
  *out = mg;
  return 0;
}

int b200_mg_create_rank(b200_mg** out, int device, int rank, int world) {
  DeviceRestore dr;
  B200_REQUIRE(out, "b200_mg_create_rank: out is null");
  *out = nullptr;
  B200_REQUIRE(world >= 1 && world <= kMaxTeam && rank >= 0 && rank < world,
               "b200_mg_create_rank: rank %d / world %d", rank, world);
  b200_mg* mg = new b200_mg();
  mg->world = world;
  mg->self = rank;
  if (common_init(mg) || setup_member(mg, rank, device)) { b200_mg_destroy(mg); return 1; }
  *out = mg;
  return 0;
}

int b200_mg_destroy(b200_mg* mg) {
  DeviceRestore dr;
  if (!mg) return 0;
  stop_workers(mg);
  for (int g = 0; g < mg->world; g++) {
    MgMember& me = mg->mem[g];
    if (me.ctx) {
      cudaSetDevice(me.ctx->device);
      if (me.push) { cudaStreamSynchronize(me.push); cudaStreamDestroy(me.push); }
      if (me.push_done) cudaEventDestroy(me.push_done);
      if (me.flags) cudaFree(me.flags);
      if (me.vals) cudaFree(me.vals);
      b200_ctx_destroy(me.ctx);
    } else if (me.flags_imported && me.flags) {
      cudaIpcCloseMemHandle(me.flags);
    }
  }
  cudaGetLastError();
  delete mg;
  return 0;
}

int b200_mg_size(const b200_mg* mg) { return mg ? mg->world : 0; }

int b200_mg_ctx(b200_mg* mg, int member, b200_ctx** ctx) {
  B200_REQUIRE(mg && ctx && member >= 0 && member < mg->world, "b200_mg_ctx: bad argument");
  B200_REQUIRE(mg->mem[member].ctx, "b200_mg_ctx: member %d belongs to another process", member);
  *ctx = mg->mem[member].ctx;
  return 0;
}

int b200_mg_sync(b200_mg* mg) {
  DeviceRestore dr;
  B200_REQUIRE(mg, "b200_mg_sync: null team");
  for (int g = 0; g < mg->world; g++)
    if (sync_member(mg->mem[g])) return 1;
  return 0;
}

// ---- rank mode: CUDA IPC plumbing (the caller moves the 64-byte handles between processes)

int b200_mg_export(b200_mg* mg, const void* dev, void* handle64) {
  DeviceRestore dr;
  B200_REQUIRE(mg && handle64 && mg->self >= 0, "b200_mg_export: needs a rank-mode team");
  B200_CUDA(cudaSetDevice(mg->mem[mg->self].ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == B200_MG_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  B200_CUDA(cudaIpcGetMemHandle(&h, dev ? const_cast<void*>(dev) : (void*)mg->mem[mg->self].flags));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

int b200_mg_import(b200_mg* mg, int member, const void* handle64, void** dev) {
  DeviceRestore dr;
  B200_REQUIRE(mg && handle64 && mg->self >= 0 && member >= 0 && member < mg->world && member != mg->self,
               "b200_mg_import: bad argument (member %d)", member);
  B200_CUDA(cudaSetDevice(mg->mem[mg->self].ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  B200_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  if (dev) {
    *dev = p;  // a data buffer of that member
  } else {      // its flag page
    mg->mem[member].flags = (uint32_t*)p;
    mg->mem[member].flags_imported = true;
  }
  return 0;
}

int b200_mg_unmap(b200_mg* mg, void* dev) {
  DeviceRestore dr;
  B200_REQUIRE(mg && mg->self >= 0, "b200_mg_unmap: needs a rank-mode team");
  if (dev) B200_CUDA(cudaIpcCloseMemHandle(dev));
  return 0;
}

// ------------------------------------------------------------------ one member's step

int b200_mg_gemm(b200_mg* mg, int member, const b200_mg_gemm_args* args) {
  DeviceRestore dr;
  B200_REQUIRE(mg && args, "b200_mg_gemm: null argument");
  B200_REQUIRE(args->elem == 4 || args->elem == 8, "b200_mg_gemm: elem must be 4 or 8");
  return args->elem == 8 ? gemm_step<double>(mg, member, *args) : gemm_step<float>(mg, member, *args);
}

int b200_mg_gemv(b200_mg* mg, int member, const b200_mg_gemv_args* args) {
  DeviceRestore dr;
  B200_REQUIRE(mg && args, "b200_mg_gemv: null argument");
  B200_REQUIRE(args->elem == 4 || args->elem == 8, "b200_mg_gemv: elem must be 4 or 8");
  return args->elem == 8 ? gemv_step<double>(mg, member, *args) : gemv_step<float>(mg, member, *args);
}

int b200_mg_gemm_team(b200_mg* mg, const b200_mg_gemm_args* args) {
  DeviceRestore dr;
  B200_REQUIRE(mg && args, "b200_mg_gemm_team: null argument");
  return run_members(mg, args[0].active, [mg, args](int g) { return b200_mg_gemm(mg, g, &args[g]); });
}

int b200_mg_gemv_team(b200_mg* mg, const b200_mg_gemv_args* args) {
  DeviceRestore dr;
  B200_REQUIRE(mg && args, "b200_mg_gemv_team: null argument");
  return run_members(mg, args[0].active, [mg, args](int g) { return b200_mg_gemv(mg, g, &args[g]); });
}

// ------------------------------------------------------------------ whole problems (local teams)
//
// The five virtuals of the reference backend class (initialise / preLoopRequirements / callGemm
// / postLoopRequirements / postCallKernelCleanup, cuBLAS/gemm.hh:45-237), for a team.

int b200_mg_problem_create(b200_mg* mg, int kind, int elem, int mode, int m, int n, int k, void** h0, void** h1,
                           void** h2, b200_mg_problem** out) {
  DeviceRestore dr;
  B200_REQUIRE(mg && out && h0 && h1 && h2, "b200_mg_problem_create: null argument");
  B200_REQUIRE(mg->self < 0, "b200_mg_problem_create: whole-problem calls need a local team (b200_mg_create)");
  B200_REQUIRE((kind == B200_MG_GEMM || kind == B200_MG_GEMV) && (elem == 4 || elem == 8) && m >= 0 && n >= 0 && k >= 0,
               "b200_mg_problem_create: bad argument");
  B200_REQUIRE(mode == B200_OFFLOAD_ALWAYS || mode == B200_OFFLOAD_ONCE || mode == B200_OFFLOAD_UNIFIED,
               "b200_mg_problem_create: unknown offload mode %d", mode);
  b200_mg_problem* p = new b200_mg_problem();
  p->mg = mg;
  p->kind = kind; p->elem = elem; p->mode = mode; p->m = m; p->n = n; p->k = kind == B200_MG_GEMM ? k : 0;
  const size_t s = (size_t)elem;
  b200_ctx* c0 = mg->mem[0].ctx;
  cudaSetDevice(c0->device);
  if (kind == B200_MG_GEMM) {
    p->active = b200_mg_active_gemm(mg->world, elem, m, n, k);
    p->bytes[0] = s * (size_t)m * k; p->bytes[1] = s * (size_t)k * n; p->bytes[2] = s * (size_t)m * n;
  } else {
    p->active = b200_mg_active_gemv(mg->world, elem, m, n);
    p->bytes[0] = s * (size_t)m * n; p->bytes[1] = s * (size_t)n; p->bytes[2] = s * (size_t)m;
  }
  // Unified mode stays on ONE GPU unless B200_MG_UNIFIED=spread: managed pages migrate 2 MiB at a
  // time under the UVM driver, and operands shared (A, x) or strided (row blocks) across GPUs
  // ping-pong -- measured with the harness, 2 GPUs: DGEMM 4096^3 26.2 -> 2.7 TFLOP/s, DGEMV 8192^2
  // 106 -> 1.2 GFLOP/s (profiles/r2_multigpu.md).
  static const bool spread_unified = getenv("B200_MG_UNIFIED") && !strcmp(getenv("B200_MG_UNIFIED"), "spread");
  if (mode == B200_OFFLOAD_UNIFIED && !spread_unified) p->active = 1;
  for (int i = 0; i < 3; i++)
    if (b200_alloc(c0, mode, p->bytes[i], &p->h[i], &p->d0[i])) { b200_mg_problem_destroy(p); return 1; }
  const int W = p->active;
  const bool unified = mode == B200_OFFLOAD_UNIFIED;
  long long unused = 0;
  for (int g = 0; g < W; g++) {
    b200_ctx* cg = mg->mem[g].ctx;
    if (kind == B200_MG_GEMM) {
      split(n, W, g, 128, &p->start[g], &p->count[g]);
      split(k, W, g, 32, &p->share0[g], &unused);
      const int ldb = std::max(1, k), ldc = std::max(1, m);
      if (unified || g == 0) {
        p->rep[g] = p->d0[0];
        p->slab[g] = (char*)p->d0[1] + s * (size_t)p->start[g] * ldb;
        p->out[g] = (char*)p->d0[2] + s * (size_t)p->start[g] * ldc;
      } else {
        cudaSetDevice(cg->device);
        if (alloc_device(cg, p->bytes[0], &p->rep[g]) || alloc_device(cg, s * (size_t)ldb * p->count[g], &p->slab[g]) ||
            alloc_device(cg, s * (size_t)ldc * p->count[g], &p->out[g])) { b200_mg_problem_destroy(p); return 1; }
      }
    } else {
      split(m, W, g, 32, &p->start[g], &p->count[g]);
      split(n, W, g, 1, &p->share0[g], &unused);
      if (unified) {
        p->rep[g] = p->d0[1];
        p->slab[g] = (char*)p->d0[0] + s * (size_t)p->start[g];  // row block inside the managed matrix (ld = m)
        p->out[g] = (char*)p->d0[2] + s * (size_t)p->start[g];
      } else if (W == 1) {
        p->rep[0] = p->d0[1]; p->slab[0] = p->d0[0]; p->out[0] = p->d0[2];
      } else {
        cudaSetDevice(cg->device);
        if (g == 0) { p->rep[0] = p->d0[1]; p->out[0] = p->d0[2]; p->slab[0] = p->d0[0]; }  // compact block fits in the full-size twin
        else if (alloc_device(cg, p->bytes[1], &p->rep[g]) ||
                 alloc_device(cg, s * (size_t)std::max<long long>(1, p->count[g]) * n, &p->slab[g]) ||
                 alloc_device(cg, s * (size_t)std::max<long long>(1, p->count[g]), &p->out[g])) { b200_mg_problem_destroy(p); return 1; }
      }
    }
  }
  if (unified && W > 1) {
    // every GPU reads all of A (GEMM) / x (GEMV): read-mostly lets each keep its own copy
    const int ri = kind == B200_MG_GEMM ? 0 : 1;
    if (mem_advise(p->h[ri], p->bytes[ri], cudaMemAdviseSetReadMostly, c0->device) != cudaSuccess) cudaGetLastError();
  } else if (unified) {
    b200_advise(c0, p->h[0], p->bytes[0], 1);
    b200_advise(c0, p->h[1], p->bytes[1], 1);
    b200_advise(c0, p->h[2], p->bytes[2], 0);
  }
  if (kind == B200_MG_GEMM && elem == 4) {
    for (int g = 0; g < W; g++) {
      cudaSetDevice(mg->mem[g].ctx->device);
      if (sgemm_tc_prepare(mg->mem[g].ctx, m, (int)p->count[g], k, std::max(1, m), std::max(1, k))) { b200_mg_problem_destroy(p); return 1; }
    }
  }
  cudaSetDevice(c0->device);
  *h0 = p->h[0]; *h1 = p->h[1]; *h2 = p->h[2];
  *out = p;
  return 0;
}

int b200_mg_problem_active(const b200_mg_problem* p) { return p ? p->active : 0; }

static int problem_step(b200_mg_problem* p, int phases, double alpha, double beta) {
  b200_mg* mg = p->mg;
  const int W = p->active;
  const size_t s = (size_t)p->elem;
  if (p->kind == B200_MG_GEMM) {
    b200_mg_gemm_args args[kMaxTeam] = {};
    for (int g = 0; g < mg->world; g++) {
      b200_mg_gemm_args& a = args[g];
      a.elem = p->elem; a.active = W; a.m = p->m; a.n_slab = g < W ? (int)p->count[g] : 0; a.k = p->k;
      a.alpha = alpha; a.beta = beta;
      a.A_rep = p->rep; a.lda = std::max(1, p->m); a.ldb = std::max(1, p->k); a.ldc = std::max(1, p->m);
      if (g < W) {
        a.hA_share = (char*)p->h[0] + s * (size_t)p->share0[g] * a.lda;
        a.hB = (char*)p->h[1] + s * (size_t)p->start[g] * a.ldb;
        a.hC = (char*)p->h[2] + s * (size_t)p->start[g] * a.ldc;
        a.dB = p->slab[g]; a.dC = p->out[g];
      }
      a.phases = phases;
    }
    return b200_mg_gemm_team(mg, args);
  }
  b200_mg_gemv_args args[kMaxTeam] = {};
  for (int g = 0; g < mg->world; g++) {
    b200_mg_gemv_args& a = args[g];
    a.elem = p->elem; a.active = W; a.m_blk = g < W ? (int)p->count[g] : 0; a.n = p->n;
    a.alpha = alpha; a.beta = beta;
    a.x_rep = p->rep;
    if (g < W) {
      a.dA = p->slab[g]; a.lda = (int)std::max<long long>(1, W == 1 ? p->m : p->count[g]);
      a.hA = (char*)p->h[0] + s * (size_t)p->start[g]; a.h_lda = std::max(1, p->m);
      a.hx_share = (char*)p->h[1] + s * (size_t)p->share0[g];
      a.hy = (char*)p->h[2] + s * (size_t)p->start[g]; a.dy = p->out[g];
    }
    a.phases = phases;
  }
  return b200_mg_gemv_team(mg, args);
}

// B200_MG_TRACE=1: host wall time of each problem-level call, with a sync after it (diagnostic)
static bool mg_trace() {
  static const bool t = getenv("B200_MG_TRACE") != nullptr;
  return t;
}
struct TraceScope {
  b200_mg_problem* p;
  const char* what;
  std::chrono::steady_clock::time_point t0;
  TraceScope(b200_mg_problem* p_, const char* w) : p(p_), what(w), t0(std::chrono::steady_clock::now()) {}
  ~TraceScope() {
    if (!mg_trace()) return;
    const double issue = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    b200_mg_sync(p->mg);
    const double done = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[b200 mg] %-8s kind=%d elem=%d mode=%d m=%d n=%d k=%d active=%d: issued %.3f ms, complete %.3f ms\n", what,
            p->kind, p->elem, p->mode, p->m, p->n, p->k, p->active, issue, done);
  }
};

int b200_mg_problem_preloop(b200_mg_problem* p, double beta) {
  DeviceRestore dr;
  B200_REQUIRE(p, "b200_mg_problem_preloop: null problem");
  TraceScope ts(p, "preloop");
  b200_mg* mg = p->mg;
  if (p->mode == B200_OFFLOAD_ONCE) return problem_step(p, B200_MG_UPLOAD | B200_MG_REPLICATE, 1.0, beta);
  if (p->mode == B200_OFFLOAD_UNIFIED) {
    const size_t s = (size_t)p->elem;
    for (int g = 0; g < p->active; g++) {
      b200_ctx* cg = mg->mem[g].ctx;
      B200_CUDA(cudaSetDevice(cg->device));
      if (p->kind == B200_MG_GEMM) {
        const size_t ldb = std::max(1, p->k), ldc = std::max(1, p->m);
        if (b200_prefetch_async(cg, p->h[0], p->bytes[0], 1, 0)) return 1;
        if (b200_prefetch_async(cg, p->slab[g], s * ldb * p->count[g], 1, 1)) return 1;
        if (b200_prefetch_async(cg, p->out[g], s * ldc * p->count[g], 1, 2)) return 1;
      } else {
        // a row block of the managed matrix is strided: the pages of all its columns move with the
        // owner of each page range; member g prefetches the g-th share of the columns, the rest
        // migrates on demand over NVLink
        long long c0, cn;
        split(p->n, p->active, g, 1, &c0, &cn);
        if (b200_prefetch_async(cg, (char*)p->h[0] + s * (size_t)c0 * std::max(1, p->m), s * (size_t)cn * std::max(1, p->m), 1, 0)) return 1;
        if (b200_prefetch_async(cg, p->h[1], p->bytes[1], 1, 1)) return 1;
        if (b200_prefetch_async(cg, p->out[g], s * (size_t)p->count[g], 1, 2)) return 1;
      }
    }
  }
  return 0;
}

int b200_mg_problem_call(b200_mg_problem* p, double alpha, double beta) {
  DeviceRestore dr;
  B200_REQUIRE(p, "b200_mg_problem_call: null problem");
  TraceScope ts(p, "call");
  b200_mg* mg = p->mg;
  if (p->mode == B200_OFFLOAD_ALWAYS) {
    if (p->active == 1) {  // small problems: the single-GPU transfer engine (zero copy / one shot / block pipeline)
      b200_ctx* c0 = mg->mem[0].ctx;
      B200_CUDA(cudaSetDevice(c0->device));
      const int lda = std::max(1, p->m), ldb = std::max(1, p->k), ldc = std::max(1, p->m);
      if (p->kind == B200_MG_GEMM)
        return p->elem == 8 ? b200_dgemm_offload(c0, p->m, p->n, p->k, alpha, (const double*)p->h[0], (double*)p->d0[0], lda, (const double*)p->h[1], (double*)p->d0[1], ldb, beta, (double*)p->h[2], (double*)p->d0[2], ldc)
                            : b200_sgemm_offload(c0, p->m, p->n, p->k, (float)alpha, (const float*)p->h[0], (float*)p->d0[0], lda, (const float*)p->h[1], (float*)p->d0[1], ldb, (float)beta, (float*)p->h[2], (float*)p->d0[2], ldc);
      return p->elem == 8 ? b200_dgemv_offload(c0, p->m, p->n, alpha, (const double*)p->h[0], (double*)p->d0[0], lda, (const double*)p->h[1], (double*)p->d0[1], beta, (double*)p->h[2], (double*)p->d0[2])
                          : b200_sgemv_offload(c0, p->m, p->n, (float)alpha, (const float*)p->h[0], (float*)p->d0[0], lda, (const float*)p->h[1], (float*)p->d0[1], (float)beta, (float*)p->h[2], (float*)p->d0[2]);
    }
    if (problem_step(p, B200_MG_UPLOAD | B200_MG_REPLICATE | B200_MG_COMPUTE | B200_MG_DOWNLOAD, alpha, beta)) return 1;
    return b200_mg_sync(mg);  // the result is valid on the host on return, like cuBLAS/gemm.hh:152
  }
  if (p->mode == B200_OFFLOAD_ONCE) return problem_step(p, B200_MG_COMPUTE, alpha, beta);
  // unified: kernels on the managed pointers
  for (int g = 0; g < p->active; g++) {
    b200_ctx* cg = mg->mem[g].ctx;
    B200_CUDA(cudaSetDevice(cg->device));
    if (p->count[g] <= 0) continue;
    if (p->kind == B200_MG_GEMM) {
      const int lda = std::max(1, p->m), ldb = std::max(1, p->k), ldc = std::max(1, p->m);
      const int rc = p->elem == 8 ? b200_dgemm(cg, 0, 0, p->m, (int)p->count[g], p->k, alpha, (const double*)p->rep[g], lda, (const double*)p->slab[g], ldb, beta, (double*)p->out[g], ldc)
                                  : b200_sgemm(cg, 0, 0, p->m, (int)p->count[g], p->k, (float)alpha, (const float*)p->rep[g], lda, (const float*)p->slab[g], ldb, (float)beta, (float*)p->out[g], ldc);
      if (rc) return rc;
    } else {
      const int lda = std::max(1, p->m);
      const int rc = p->elem == 8 ? b200_dgemv(cg, 0, (int)p->count[g], p->n, alpha, (const double*)p->slab[g], lda, (const double*)p->rep[g], 1, beta, (double*)p->out[g], 1)
                                  : b200_sgemv(cg, 0, (int)p->count[g], p->n, (float)alpha, (const float*)p->slab[g], lda, (const float*)p->rep[g], 1, (float)beta, (float*)p->out[g], 1);
      if (rc) return rc;
    }
  }
  return 0;
}

int b200_mg_problem_postloop(b200_mg_problem* p) {
  DeviceRestore dr;
  B200_REQUIRE(p, "b200_mg_problem_postloop: null problem");
  TraceScope ts(p, "postloop");
  b200_mg* mg = p->mg;
  if (p->mode == B200_OFFLOAD_ONCE) {
    if (problem_step(p, B200_MG_DOWNLOAD, 1.0, 0.0)) return 1;
    return b200_mg_sync(mg);
  }
  if (p->mode == B200_OFFLOAD_UNIFIED) {
    const size_t s = (size_t)p->elem;
    for (int g = 0; g < p->active; g++) {
      b200_ctx* cg = mg->mem[g].ctx;
      B200_CUDA(cudaSetDevice(cg->device));
      const size_t bytes = p->kind == B200_MG_GEMM ? s * (size_t)std::max(1, p->m) * p->count[g] : s * (size_t)p->count[g];
      if (b200_prefetch_async(cg, p->out[g], bytes, 0, 2)) return 1;
    }
    return b200_mg_sync(mg);
  }
  return 0;
}

int b200_mg_problem_destroy(b200_mg_problem* p) {
  DeviceRestore dr;
  if (!p) return 0;
  b200_mg* mg = p->mg;
  int rc = b200_mg_sync(mg);
  const bool unified = p->mode == B200_OFFLOAD_UNIFIED;
  if (unified && p->active > 1 && p->h[p->kind == B200_MG_GEMM ? 0 : 1]) {
    const int ri = p->kind == B200_MG_GEMM ? 0 : 1;  // pooled block: leave no advice behind
    if (mem_advise(p->h[ri], p->bytes[ri], cudaMemAdviseUnsetReadMostly, mg->mem[0].ctx->device) != cudaSuccess) cudaGetLastError();
  }
  for (int g = 1; g < p->active && !unified; g++) {
    b200_ctx* cg = mg->mem[g].ctx;
    cudaSetDevice(cg->device);
    rc |= free_device(cg, p->rep[g]);
    rc |= free_device(cg, p->slab[g]);
    rc |= free_device(cg, p->out[g]);
  }
  b200_ctx* c0 = mg->mem[0].ctx;
  cudaSetDevice(c0->device);
  for (int i = 0; i < 3; i++)
    if (p->h[i]) rc |= b200_free(c0, p->mode, p->h[i], p->d0[i]);
  delete p;
  return rc;
}

}  // extern "C"
