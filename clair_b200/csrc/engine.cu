// C-ABI + host orchestration of the clair_b200 forward path (see include/clair_b200.h).
//
// One engine drives one B200.  Weights arrive by TF variable name, are re-laid-out once in
// clairb_finalize_weights and stay resident in HBM.  A predict call is cut into chunks of whole
// predict-batches; each chunk is: async H2D -> forward kernels -> async D2H, with the copies of
// neighbouring chunks overlapping the kernels on separate streams (double-buffered input/output).
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clair_b200.h"
#include "common.cuh"
// The product library carries ONE engine: the tensor-core path (tc_kernels.cuh).  -DCLAIRB_CROSSCHECK adds the fp32
// CUDA-core kernels and the two-kernel layer 2 as on-device cross-checks (libclair_b200_xcheck.so, loaded by tests only).
#ifdef CLAIRB_CROSSCHECK
#include "simt_kernels.cuh"
#endif
#include "tc_kernels.cuh"
#include "decide_kernels.cuh"
#include "decode_host.cuh"
#include "create_tensor_kernels.cuh"
#include "encode_host.cuh"
#include "blosc_host.cuh"
#include "format_host.cuh"

using namespace clairb;

namespace {

std::mutex g_err_mu;
std::string g_create_error;

const char* kHeadNames[4] = {"Y_base_change_logits", "Y_genotype_logits", "Y_indel_length_logits_1",
                             "Y_indel_length_logits_2"};
const int kHeadSize[4] = {21, 3, 33, 33};
const int kHeadOffH[5] = {0, 21, 24, 57, 90};

std::string lstm_name(int layer, int dir, const char* var) {
  char buf[256];
  snprintf(buf, sizeof buf, "LSTM%d/stack_bidirectional_rnn/cell_0/bidirectional_rnn/%s/cudnn_compatible_lstm_cell/%s",
           layer, dir ? "bw" : "fw", var);
  return buf;
}

struct HostWeight {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

enum EngineKind { ENGINE_SIMT = 0, ENGINE_TC = 1 };
constexpr int TC_PAIR_SITES = 256;

}  // namespace

// One clairb_predict_async call: the sites of a request are cut into segments, a segment rides in one chunk ("group") of
// the pipeline next to segments of other requests, and the request is done when all its sites were delivered.
struct AsyncRequest {
  int64_t ticket = 0;
  const char* x = nullptr;
  int dtype = 0;
  int64_t n = 0;
  float* outs[4] = {nullptr, nullptr, nullptr, nullptr};   // four head arrays, or outs[0] = packed [n][90] when !split
  bool split = false;
  float* out_dev = nullptr;          // instead of outs: packed [n][90] rows left in device memory (clairb_predict_async_to_device)
  const uint8_t* ref = nullptr;      // with `dec`: first-choice decision behind the heads
  int32_t* dec = nullptr;
  int64_t issued = 0, delivered = 0;
  int rc = 0;
  std::string err;
  bool done = false;
};

struct clairb_engine {
  int device = 0;
  int64_t max_sites = 0;
  int batch = 1000;
  int bp = 1024;
  int64_t chunk_sites = 0;     // real sites per chunk
  int64_t chunk_np = 0;        // padded sites per chunk
  EngineKind kind = ENGINE_SIMT;
  bool finalized = false;
  bool fuse_tail = false;      // TC engine: slice-dense + L4 on tensor cores (l3l4_fused)
  bool l2_stream = true;       // TC engine: layer-2 input projection streamed through the recurrent kernel (lstm_seq_x2)
  bool ramp = false;           // first chunk of a multi-chunk host call is half a chunk (one wave)
  std::string err;
  std::atomic<int64_t> launches{0};

  // ---- asynchronous submission (clairb_predict_async / clairb_predict_wait) ----
  // run_mu owns the chunk pipeline (device buffers, staging, streams): a synchronous call holds it for its duration, the
  // worker thread holds it while requests are in flight.  q_mu guards the queue and the tickets.
  std::mutex run_mu, q_mu;
  std::condition_variable q_cv, done_cv;
  std::deque<AsyncRequest*> pending;
  std::map<int64_t, AsyncRequest*> tickets;
  int64_t next_ticket = 1;
  std::thread worker;
  bool worker_started = false, stop = false;
  cudaEvent_t ev_dev = nullptr;     // end of the last clairb_predict_device enqueue (the workspace is shared with it)
  bool dev_pending = false;

  std::map<std::string, HostWeight> hw;

  // device weights (SIMT layouts, fp32)
  float *d_Wp[2] = {nullptr, nullptr}, *d_bp[2] = {nullptr, nullptr};   // per layer: [2][K][512], [2][512]
  float *d_w3p = nullptr, *d_b3p = nullptr, *d_W4 = nullptr, *d_b4 = nullptr;
  float *d_W5 = nullptr, *d_b5 = nullptr, *d_Whd = nullptr, *d_bhd = nullptr;
  tc::Weights tcw;             // tensor-core operand layouts (fp16 hi/lo, tile-blocked)

  // workspace (sized for one chunk)
  float *d_xT = nullptr, *d_h1 = nullptr, *d_h2 = nullptr, *d_l3T = nullptr, *d_l4T = nullptr, *d_logits = nullptr;
  tc::Workspace tcws;
  void* d_x[2] = {nullptr, nullptr};
  float* d_out[2] = {nullptr, nullptr};
  float* h_out[2] = {nullptr, nullptr};     // pinned staging for results (one per output buffer)
  // first-choice decision stage (clairb_predict_decide / clairb_decide): reference bases in, 6-word records out
  uint8_t* d_ref[2] = {nullptr, nullptr};
  int32_t* d_dec[2] = {nullptr, nullptr};
  int32_t* h_dec[2] = {nullptr, nullptr};   // pinned staging
  uint8_t* h_ref[2] = {nullptr, nullptr};   // pinned staging (a pageable source would make the copy synchronous and stall the chunk pipeline)

  cudaStream_t s_h2d = nullptr, s_h2d2 = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d2[2] = {nullptr, nullptr};
  int h2d_split = 1;           // CLAIRB_H2D_SPLIT=2: the input copy of a chunk goes out as two halves on two streams
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;

  SiteMap last_map{0, 1000, 1024, 0};
  bool last_single_chunk = false;

  // CreateTensor on the device (clairb_create_tensors / clairb_predict_created): growable device buffers.  ct_x holds the
  // tensor block of the most recent clairb_create_tensors call ([ct_rows][1056] int16), the rest is upload scratch.
  struct DevBuf { void* p = nullptr; size_t cap = 0; };
  DevBuf ct_in[12], ct_x, ct_meta, ct_gather, ct_rows_idx, ct_out;
  int* ct_overflow = nullptr;
  int64_t ct_rows = 0;
  bool ct_subtracted = false;   // the resident block was written with CLAIRB_CT_SUBTRACT (what the forward takes)

  // optional per-kernel CUDA-event timing (bench.py's roofline leg)
  bool profiling = false;
  struct ProfSpan { cudaEvent_t a, b; int id; };
  std::vector<ProfSpan> prof_open;
  std::vector<cudaEvent_t> prof_free;
  std::map<int, std::pair<double, int64_t>> prof_acc;     // kernel id -> (ms, launches)
};

namespace {

int fail(clairb_engine* e, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  else {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_create_error = buf;
  }
  return code;
}

#define CU_TRY(e, call)                                                                          \
  do {                                                                                           \
    cudaError_t _st = (call);                                                                    \
    if (_st != cudaSuccess)                                                                      \
      return fail((e), _st == cudaErrorMemoryAllocation ? CLAIRB_ENOMEM : CLAIRB_ECUDA,          \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__);  \
  } while (0)

template <typename Tp>
int dev_alloc(clairb_engine* e, Tp** p, size_t count) {
  CU_TRY(e, cudaMalloc((void**)p, count * sizeof(Tp)));
  return CLAIRB_OK;
}

// device buffer that only ever grows (with 25 % slack), for inputs whose size changes from call to call
int grow(clairb_engine* e, clairb_engine::DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return CLAIRB_OK;
  if (b.p) CU_TRY(e, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 4 + 256;
  CU_TRY(e, cudaMalloc(&b.p, want));
  b.cap = want;
  return CLAIRB_OK;
}

template <typename Tp>
int upload(clairb_engine* e, Tp** dptr, const std::vector<Tp>& h) {
  int rc = dev_alloc(e, dptr, h.size());
  if (rc) return rc;
  CU_TRY(e, cudaMemcpy(*dptr, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice));
  return CLAIRB_OK;
}

const HostWeight* find_weight(clairb_engine* e, const std::string& name, std::vector<int64_t> shape) {
  auto it = e->hw.find(name);
  if (it == e->hw.end()) {
    fail(e, CLAIRB_EWEIGHTS, "variable %s was never set", name.c_str());
    return nullptr;
  }
  if (it->second.shape != shape) {
    fail(e, CLAIRB_EWEIGHTS, "variable %s has the wrong shape", name.c_str());
    return nullptr;
  }
  return &it->second;
}

SiteMap make_map(const clairb_engine* e, int64_t n) {
  SiteMap m;
  m.n = n;
  if (e->kind == ENGINE_TC) {
    // sites are independent, so the tensor-core path tiles straight through predict-batch boundaries:
    // padded row == real row, padded up to a whole CTA pair (256 sites)
    m.batch = m.bp = TC_PAIR_SITES;
    m.np = (n + TC_PAIR_SITES - 1) / TC_PAIR_SITES * TC_PAIR_SITES;
    return m;
  }
  m.batch = e->batch;
  m.bp = e->bp;
  int64_t nb = (n + e->batch - 1) / e->batch;
  m.np = nb * e->bp;
  return m;
}

// ---- per-kernel event timing --------------------------------------------------------------------
const char* kKernelNames[] = {"prep_input", "lstm_layer1", "lstm_layer2", "l3_slice_dense", "l4_dense",
                              "tail_heads", "prep_tiles", "lstm_seq1", "xproj2", "lstm_seq2", "l3l4_fused", "heads_tc",
                              "lstm_seq_x2", "decide_sites", "create_tensors", "gather_rows"};
constexpr int kNumKernelNames = 16;

void prof_fold(clairb_engine* e) {
  for (auto& sp : e->prof_open) {
    cudaEventSynchronize(sp.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      auto& acc = e->prof_acc[sp.id];
      acc.first += ms;
      acc.second += 1;
    }
    e->prof_free.push_back(sp.a);
    e->prof_free.push_back(sp.b);
  }
  e->prof_open.clear();
}

cudaEvent_t prof_event(clairb_engine* e) {
  if (!e->prof_free.empty()) {
    cudaEvent_t ev = e->prof_free.back();
    e->prof_free.pop_back();
    return ev;
  }
  cudaEvent_t ev = nullptr;
  cudaEventCreate(&ev);
  return ev;
}

struct ProfScope {
  clairb_engine* e;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  ProfScope(clairb_engine* e_, int id, cudaStream_t st_) : e(e_), st(st_) {
    if (!e->profiling) return;
    if (e->prof_open.size() >= 4096) prof_fold(e);
    cudaEvent_t a = prof_event(e);
    b = prof_event(e);
    cudaEventRecord(a, st);
    e->prof_open.push_back({a, b, id});
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, st);
  }
};

// ---- forward on one chunk, all launches on `st` ------------------------------------------------
#ifdef CLAIRB_CROSSCHECK
int forward_tail(clairb_engine* e, SiteMap sm, float* out_dev, cudaStream_t st);
int forward_heads(clairb_engine* e, SiteMap sm, float* out_dev, cudaStream_t st);

int forward_simt(clairb_engine* e, const void* x_dev, int dtype, SiteMap sm, float* out_dev, cudaStream_t st) {
  const int64_t np = sm.np;
  dim3 gprep((unsigned)(np / 32), 4);
  {
    ProfScope ps(e, 0, st);
    if (dtype == CLAIRB_DTYPE_F32)
      simt::prep_input<float><<<gprep, 256, 0, st>>>((const float*)x_dev, e->d_xT, sm);
    else
      simt::prep_input<int16_t><<<gprep, 256, 0, st>>>((const int16_t*)x_dev, e->d_xT, sm);
  }
  dim3 glstm((unsigned)(np / simt::LSTM_TM), 2);
  {
    ProfScope ps(e, 1, st);
    simt::lstm_layer<F_IN><<<glstm, simt::LSTM_THREADS, simt::lstm_smem_bytes<F_IN>(), st>>>(
        e->d_xT, e->d_Wp[0], e->d_bp[0], e->d_h1, np);
  }
  {
    ProfScope ps(e, 2, st);
    simt::lstm_layer<2 * H><<<glstm, simt::LSTM_THREADS, simt::lstm_smem_bytes<2 * H>(), st>>>(
        e->d_h1, e->d_Wp[1], e->d_bp[1], e->d_h2, np);
  }
  return forward_tail(e, sm, out_dev, st);
}

// slice-dense L3 -> L4 -> L5/heads/softmax from the h2 planes (shared by both engines for now)
int forward_tail(clairb_engine* e, SiteMap sm, float* out_dev, cudaStream_t st) {
  const int64_t np = sm.np;
  dim3 gl3((unsigned)(np / 128), 2 * H / simt::L3_CPB);
  {
    ProfScope ps(e, 3, st);
    simt::l3_slice_dense<<<gl3, 128, 0, st>>>(e->d_h2, e->d_w3p, e->d_b3p, e->d_l3T, np);
  }
  {
    ProfScope ps(e, 4, st);
    simt::l4_dense<<<(unsigned)(np / simt::L4_TM), 256, simt::l4_smem_bytes(), st>>>(e->d_l3T, e->d_W4, e->d_b4,
                                                                                    e->d_l4T, np);
  }
  e->launches += 2;
  return forward_heads(e, sm, out_dev, st);
}

// L5_1..4 -> heads -> softmax from the l4T planes
int forward_heads(clairb_engine* e, SiteMap sm, float* out_dev, cudaStream_t st) {
  const int64_t np = sm.np;
  {
    ProfScope ps(e, 5, st);
    simt::tail_heads<<<(unsigned)(np / simt::TL_TM), 256, simt::tail_smem_bytes(), st>>>(
        e->d_l4T, e->d_W5, e->d_b5, e->d_Whd, e->d_bhd, out_dev, e->d_logits, sm);
  }
  e->launches += 1;
  CU_TRY(e, cudaGetLastError());
  return CLAIRB_OK;
}

#endif  // CLAIRB_CROSSCHECK

int forward_chunk(clairb_engine* e, const void* x_dev, int dtype, SiteMap sm, float* out_dev, cudaStream_t st,
                  int64_t split_rows = 0) {
  if (e->kind == ENGINE_TC) {
    int nl = 0;
    ProfScope* open = nullptr;
    auto hook = [&](int id, bool begin) {
      if (begin) open = new ProfScope(e, 6 + id, st);
      else { delete open; open = nullptr; }
    };
    cudaError_t cst = tc::forward_lstm(e->tcw, e->tcws, x_dev, dtype == CLAIRB_DTYPE_I16, sm.n, sm.np, e->d_h2, e->d_l4T,
                                       out_dev, e->d_logits, e->fuse_tail, e->l2_stream, st, &nl, hook, split_rows);
    e->launches += nl;
    if (cst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "tensor-core forward failed: %s", cudaGetErrorString(cst));
    if (e->fuse_tail) return CLAIRB_OK;
#ifdef CLAIRB_CROSSCHECK
    return forward_tail(e, sm, out_dev, st);
#endif
  }
#ifdef CLAIRB_CROSSCHECK
  return forward_simt(e, x_dev, dtype, sm, out_dev, st);
#else
  return fail(e, CLAIRB_EINVAL, "this build carries the tensor-core engine only");
#endif
}

size_t elem_bytes(int dtype) { return dtype == CLAIRB_DTYPE_I16 ? 2 : 4; }

void free_all(clairb_engine* e) {
  cudaSetDevice(e->device);
  prof_fold(e);
  for (cudaEvent_t ev : e->prof_free) cudaEventDestroy(ev);
  e->prof_free.clear();
  for (int l = 0; l < 2; ++l) {
    cudaFree(e->d_Wp[l]);
    cudaFree(e->d_bp[l]);
    cudaFree(e->d_x[l]);
    cudaFree(e->d_out[l]);
    if (e->h_out[l]) cudaFreeHost(e->h_out[l]);
    cudaFree(e->d_ref[l]);
    cudaFree(e->d_dec[l]);
    if (e->h_dec[l]) cudaFreeHost(e->h_dec[l]);
    if (e->h_ref[l]) cudaFreeHost(e->h_ref[l]);
    if (e->ev_h2d[l]) cudaEventDestroy(e->ev_h2d[l]);
    if (e->ev_comp[l]) cudaEventDestroy(e->ev_comp[l]);
    if (e->ev_d2h[l]) cudaEventDestroy(e->ev_d2h[l]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_dev) cudaEventDestroy(e->ev_dev);
  cudaFree(e->d_w3p); cudaFree(e->d_b3p); cudaFree(e->d_W4); cudaFree(e->d_b4);
  cudaFree(e->d_W5); cudaFree(e->d_b5); cudaFree(e->d_Whd); cudaFree(e->d_bhd);
  cudaFree(e->d_xT); cudaFree(e->d_h1); cudaFree(e->d_h2); cudaFree(e->d_l3T); cudaFree(e->d_l4T);
  cudaFree(e->d_logits);
  tc::free_weights(e->tcw);
  tc::free_workspace(e->tcws);
  for (auto& b : e->ct_in) cudaFree(b.p);
  cudaFree(e->ct_x.p); cudaFree(e->ct_meta.p); cudaFree(e->ct_gather.p); cudaFree(e->ct_rows_idx.p); cudaFree(e->ct_out.p);
  cudaFree(e->ct_overflow);
  if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
  if (e->s_h2d2) cudaStreamDestroy(e->s_h2d2);
  for (int l = 0; l < 2; ++l) if (e->ev_h2d2[l]) cudaEventDestroy(e->ev_h2d2[l]);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
}

}  // namespace

extern "C" {

const char* clairb_version(void) {
#ifdef CLAIRB_CROSSCHECK
  return "clair_b200 0.6 sm_100a cross-check build (tcgen05 engine + fp32 CUDA-core engines)";
#else
  return "clair_b200 0.6 sm_100a (tcgen05 BiLSTM, streamed layer-2 projection, async batch queue, decision stage, create_tensors, training step)";
#endif
}

const char* clairb_last_error(const clairb_engine* e) {
  if (e) return e->err.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  return g_create_error.c_str();
}

int64_t clairb_kernel_launches(const clairb_engine* e) { return e ? e->launches.load() : 0; }

int clairb_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return CLAIRB_EINVAL;
  cudaError_t st = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable);
  return st == cudaSuccess ? CLAIRB_OK : CLAIRB_ENOMEM;
}

int clairb_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? CLAIRB_OK : CLAIRB_ECUDA; }

int clairb_create(int device, int64_t max_sites, int batch_sites, clairb_engine** out) {
  if (!out || max_sites <= 0 || batch_sites <= 0) return fail(nullptr, CLAIRB_EINVAL, "clairb_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(nullptr, CLAIRB_ENODEVICE, "clairb_create: CUDA device %d not available (%d visible)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return fail(nullptr, CLAIRB_ENODEVICE,
                "clairb_create: device %d is not compute capability 10.x (sm_100a only, no fallback)", device);
  clairb_engine* e = new clairb_engine();
  e->device = device;
  e->max_sites = max_sites;
  e->batch = batch_sites;
  e->bp = (batch_sites + TILE - 1) / TILE * TILE;
  e->kind = ENGINE_TC;
  e->fuse_tail = true;
  e->l2_stream = true;
  {
    const char* kind = getenv("CLAIRB_ENGINE");
    const char* ft = getenv("CLAIRB_FUSED_TAIL");
    const char* ls = getenv("CLAIRB_L2_STREAM");
    const bool want_simt = kind && !strcmp(kind, "simt"), no_fuse = ft && !strcmp(ft, "0"), no_stream = ls && !strcmp(ls, "0");
#ifdef CLAIRB_CROSSCHECK
    if (want_simt) e->kind = ENGINE_SIMT;
    e->fuse_tail = !no_fuse;
    e->l2_stream = !no_stream;
#else
    if (want_simt || no_fuse || no_stream) {
      delete e;
      return fail(nullptr, CLAIRB_EINVAL, "clairb_create: CLAIRB_ENGINE / CLAIRB_FUSED_TAIL / CLAIRB_L2_STREAM select the cross-check "
                  "engines, which this library was built without (they live in libclair_b200_xcheck.so, -DCLAIRB_CROSSCHECK)");
    }
#endif
    const char* rp = getenv("CLAIRB_RAMP");
    e->ramp = !(rp && !strcmp(rp, "0"));
  }
  if (e->kind == ENGINE_TC) {
    // Sites are independent, so a chunk need not be whole predict-batches.  Default: 2 waves of CTA pairs per chunk
    // (sm_count/2 pairs x 256 sites x 2 waves / 2 directions = 18,944 sites on 148 SMs = exactly one wave of
    // 148 tiles for the per-tile kernels); small enough that the H2D copy of the next chunk hides behind this one.
    int64_t cs = (int64_t)(prop.multiProcessorCount / 2) * TC_PAIR_SITES;
    if (const char* c = getenv("CLAIRB_CHUNK_SITES")) cs = atoll(c) > 0 ? atoll(c) : cs;
    if (cs > max_sites) cs = max_sites;
    e->chunk_sites = cs;
    e->chunk_np = (cs + TC_PAIR_SITES - 1) / TC_PAIR_SITES * TC_PAIR_SITES;
  } else {
    int64_t chunk_batches = 32;
    if (const char* cb = getenv("CLAIRB_CHUNK_BATCHES")) chunk_batches = atoll(cb) > 0 ? atoll(cb) : chunk_batches;
    int64_t nb_max = (max_sites + batch_sites - 1) / batch_sites;
    if (chunk_batches > nb_max) chunk_batches = nb_max;
    e->chunk_sites = chunk_batches * batch_sites;
    e->chunk_np = chunk_batches * e->bp;
  }
  auto bail = [&](int rc) {
    g_create_error = e->err;
    free_all(e);
    delete e;
    return rc;
  };
#define CR_TRY(call)                                                                  \
  do {                                                                                \
    cudaError_t _st = (call);                                                         \
    if (_st != cudaSuccess) {                                                         \
      fail(e, CLAIRB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_st));         \
      return bail(_st == cudaErrorMemoryAllocation ? CLAIRB_ENOMEM : CLAIRB_ECUDA);   \
    }                                                                                 \
  } while (0)
  CR_TRY(cudaSetDevice(device));
  CR_TRY(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
  CR_TRY(cudaStreamCreateWithFlags(&e->s_h2d2, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) CR_TRY(cudaEventCreateWithFlags(&e->ev_h2d2[b], cudaEventDisableTiming));
  if (const char* hs = getenv("CLAIRB_H2D_SPLIT")) e->h2d_split = atoi(hs) == 2 ? 2 : 1;
  CR_TRY(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
  CR_TRY(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CR_TRY(cudaEventCreateWithFlags(&e->ev_h2d[b], cudaEventDisableTiming));
    CR_TRY(cudaEventCreateWithFlags(&e->ev_comp[b], cudaEventDisableTiming));
    CR_TRY(cudaEventCreateWithFlags(&e->ev_d2h[b], cudaEventDisableTiming));
    CR_TRY(cudaMalloc(&e->d_x[b], (size_t)e->chunk_sites * SITE_ELEMS * sizeof(float)));
    CR_TRY(cudaMalloc((void**)&e->d_out[b], (size_t)e->chunk_sites * N_OUT * sizeof(float)));
    CR_TRY(cudaHostAlloc((void**)&e->h_out[b], (size_t)e->chunk_sites * N_OUT * sizeof(float), cudaHostAllocDefault));
    CR_TRY(cudaMalloc((void**)&e->d_ref[b], (size_t)e->chunk_sites));
    CR_TRY(cudaMalloc((void**)&e->d_dec[b], (size_t)e->chunk_sites * decide::REC_WORDS * sizeof(int32_t)));
    CR_TRY(cudaHostAlloc((void**)&e->h_dec[b], (size_t)e->chunk_sites * decide::REC_WORDS * sizeof(int32_t), cudaHostAllocDefault));
    CR_TRY(cudaHostAlloc((void**)&e->h_ref[b], (size_t)e->chunk_sites, cudaHostAllocDefault));
  }
  CR_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CR_TRY(cudaEventCreateWithFlags(&e->ev_dev, cudaEventDisableTiming));
  const size_t np = (size_t)e->chunk_np;
  CR_TRY(cudaMalloc((void**)&e->d_logits, np * N_OUT * sizeof(float)));
  CR_TRY(cudaMemset(e->d_logits, 0, np * N_OUT * sizeof(float)));
  // fp32 planes of the LSTM2 / L3 activations: only the CUDA-core tail reads them (the fused tensor-core path keeps both
  // on chip; the L3 parity hook allocates its sink on first use)
  if (e->kind == ENGINE_SIMT || !e->fuse_tail) {
    CR_TRY(cudaMalloc((void**)&e->d_h2, np * T_STEPS * 2 * H * sizeof(float)));
    CR_TRY(cudaMalloc((void**)&e->d_l3T, np * L3_K * sizeof(float)));
  }
  CR_TRY(cudaMalloc((void**)&e->d_l4T, np * L4_UNITS * sizeof(float)));
#ifdef CLAIRB_CROSSCHECK
  CR_TRY(cudaFuncSetAttribute(simt::l4_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)simt::l4_smem_bytes()));
  CR_TRY(cudaFuncSetAttribute(simt::tail_heads, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)simt::tail_smem_bytes()));
  if (e->kind == ENGINE_SIMT) {
    CR_TRY(cudaMalloc((void**)&e->d_xT, np * SITE_ELEMS * sizeof(float)));
    CR_TRY(cudaMalloc((void**)&e->d_h1, np * T_STEPS * 2 * H * sizeof(float)));
    CR_TRY(cudaFuncSetAttribute(simt::lstm_layer<F_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)simt::lstm_smem_bytes<F_IN>()));
    CR_TRY(cudaFuncSetAttribute(simt::lstm_layer<2 * H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)simt::lstm_smem_bytes<2 * H>()));
  } else
#endif
  {
    CR_TRY(tc::alloc_workspace(e->tcws, e->chunk_np, device, !e->l2_stream));
  }
#undef CR_TRY
  *out = e;
  return CLAIRB_OK;
}

int clairb_set_weight(clairb_engine* e, const char* tf_name, const float* data, const int64_t* shape, int rank) {
  if (!e) return CLAIRB_EINVAL;
  if (!tf_name || !data || !shape || rank < 1 || rank > 2)
    return fail(e, CLAIRB_EINVAL, "clairb_set_weight: bad arguments");
  HostWeight w;
  size_t count = 1;
  for (int i = 0; i < rank; ++i) {
    if (shape[i] <= 0) return fail(e, CLAIRB_EINVAL, "clairb_set_weight: non-positive dimension");
    w.shape.push_back(shape[i]);
    count *= (size_t)shape[i];
  }
  w.data.assign(data, data + count);
  e->hw[tf_name] = std::move(w);
  e->finalized = false;
  return CLAIRB_OK;
}

int clairb_finalize_weights(clairb_engine* e) {
  if (!e) return CLAIRB_EINVAL;
  CU_TRY(e, cudaSetDevice(e->device));
  // ---- gather + check every variable of the forward graph ----
  const int fin[2] = {F_IN, 2 * H};
  std::vector<float> Wp[2], bp[2];
  for (int l = 0; l < 2; ++l) {
    const int K = fin[l] + H;
    Wp[l].assign((size_t)2 * K * G4, 0.f);
    bp[l].assign((size_t)2 * G4, 0.f);
    for (int d = 0; d < 2; ++d) {
      const HostWeight* k = find_weight(e, lstm_name(l + 1, d, "kernel"), {K, G4});
      const HostWeight* b = find_weight(e, lstm_name(l + 1, d, "bias"), {G4});
      if (!k || !b) return CLAIRB_EWEIGHTS;
      for (int g = 0; g < 4; ++g)
        for (int u = 0; u < H; ++u) {
          const int colp = (u / 64) * 256 + (u % 64) * 4 + g, col = g * H + u;
          bp[l][(size_t)d * G4 + colp] = b->data[col];
          for (int r = 0; r < K; ++r) Wp[l][((size_t)d * K + r) * G4 + colp] = k->data[(size_t)r * G4 + col];
        }
    }
  }
  std::vector<float> w3p((size_t)2 * H * T_STEPS * 32, 0.f), b3p((size_t)2 * H * 32, 0.f);
  std::vector<float> w3((size_t)2 * H * T_STEPS * L3_UNITS), b3((size_t)2 * H * L3_UNITS);
  for (int c = 0; c < 2 * H; ++c) {
    char nm[64];
    snprintf(nm, sizeof nm, "L3/Unit_%d/kernel", c);
    const HostWeight* k = find_weight(e, nm, {T_STEPS, L3_UNITS});
    snprintf(nm, sizeof nm, "L3/Unit_%d/bias", c);
    const HostWeight* b = find_weight(e, nm, {L3_UNITS});
    if (!k || !b) return CLAIRB_EWEIGHTS;
    for (int t = 0; t < T_STEPS; ++t)
      for (int o = 0; o < L3_UNITS; ++o) {
        w3p[((size_t)c * T_STEPS + t) * 32 + o] = k->data[t * L3_UNITS + o];
        w3[((size_t)c * T_STEPS + t) * L3_UNITS + o] = k->data[t * L3_UNITS + o];
      }
    for (int o = 0; o < L3_UNITS; ++o) {
      b3p[(size_t)c * 32 + o] = b->data[o];
      b3[(size_t)c * L3_UNITS + o] = b->data[o];
    }
  }
  const HostWeight* W4 = find_weight(e, "L4/kernel", {L3_K, L4_UNITS});
  const HostWeight* b4 = find_weight(e, "L4/bias", {L4_UNITS});
  if (!W4 || !b4) return CLAIRB_EWEIGHTS;
  std::vector<float> W5((size_t)L4_UNITS * L5_ALL), b5(L5_ALL), Whd((size_t)L5_UNITS * N_OUT), bhd(N_OUT);
  for (int k = 0; k < 4; ++k) {
    char nm[96];
    snprintf(nm, sizeof nm, "L5_%d/kernel", k + 1);
    const HostWeight* kk = find_weight(e, nm, {L4_UNITS, L5_UNITS});
    snprintf(nm, sizeof nm, "L5_%d/bias", k + 1);
    const HostWeight* bb = find_weight(e, nm, {L5_UNITS});
    snprintf(nm, sizeof nm, "Prediction/%s/kernel", kHeadNames[k]);
    const HostWeight* hk = find_weight(e, nm, {L5_UNITS, kHeadSize[k]});
    snprintf(nm, sizeof nm, "Prediction/%s/bias", kHeadNames[k]);
    const HostWeight* hb = find_weight(e, nm, {kHeadSize[k]});
    if (!kk || !bb || !hk || !hb) return CLAIRB_EWEIGHTS;
    for (int r = 0; r < L4_UNITS; ++r)
      for (int j = 0; j < L5_UNITS; ++j) W5[(size_t)r * L5_ALL + k * L5_UNITS + j] = kk->data[r * L5_UNITS + j];
    for (int j = 0; j < L5_UNITS; ++j) b5[k * L5_UNITS + j] = bb->data[j];
    for (int j = 0; j < L5_UNITS; ++j)
      for (int o = 0; o < kHeadSize[k]; ++o) Whd[(size_t)j * N_OUT + kHeadOffH[k] + o] = hk->data[j * kHeadSize[k] + o];
    for (int o = 0; o < kHeadSize[k]; ++o) bhd[kHeadOffH[k] + o] = hb->data[o];
  }
  // ---- upload ----
  auto drop = [](float*& p) { cudaFree(p); p = nullptr; };
  for (int l = 0; l < 2; ++l) { drop(e->d_Wp[l]); drop(e->d_bp[l]); }
  drop(e->d_w3p); drop(e->d_b3p); drop(e->d_W4); drop(e->d_b4);
  drop(e->d_W5); drop(e->d_b5); drop(e->d_Whd); drop(e->d_bhd);
  int rc = 0;
  if ((rc = upload(e, &e->d_b4, b4->data))) return rc;                // read by l3l4_fused
#ifdef CLAIRB_CROSSCHECK
  // fp32 layouts of the CUDA-core tail (slice-dense, L4, L5 / heads)
  if ((rc = upload(e, &e->d_W5, W5)) || (rc = upload(e, &e->d_b5, b5)) || (rc = upload(e, &e->d_Whd, Whd)) ||
      (rc = upload(e, &e->d_bhd, bhd)))
    return rc;
  if ((rc = upload(e, &e->d_w3p, w3p)) || (rc = upload(e, &e->d_b3p, b3p)) || (rc = upload(e, &e->d_W4, W4->data)))
    return rc;
#endif
  if (e->kind == ENGINE_SIMT) {
    for (int l = 0; l < 2; ++l)
      if ((rc = upload(e, &e->d_Wp[l], Wp[l])) || (rc = upload(e, &e->d_bp[l], bp[l]))) return rc;
  } else {
    tc::HostModel hm;
    for (int l = 0; l < 2; ++l)
      for (int d = 0; d < 2; ++d) {
        hm.lstm_kernel[l][d] = e->hw[lstm_name(l + 1, d, "kernel")].data.data();
        hm.lstm_bias[l][d] = e->hw[lstm_name(l + 1, d, "bias")].data.data();
      }
    hm.w3 = w3.data(); hm.b3 = b3.data(); hm.W4 = W4->data.data();
    for (int k = 0; k < 4; ++k) {
      char nm[96];
      snprintf(nm, sizeof nm, "L5_%d/kernel", k + 1);
      hm.W5[k] = e->hw[nm].data.data();
      snprintf(nm, sizeof nm, "L5_%d/bias", k + 1);
      hm.b5[k] = e->hw[nm].data.data();
      snprintf(nm, sizeof nm, "Prediction/%s/kernel", kHeadNames[k]);
      hm.Whd[k] = e->hw[nm].data.data();
      snprintf(nm, sizeof nm, "Prediction/%s/bias", kHeadNames[k]);
      hm.bhd[k] = e->hw[nm].data.data();
    }
    tc::free_weights(e->tcw);
    e->tcw.b4 = e->d_b4;
    cudaError_t cst = tc::build_weights(e->tcw, hm);
    if (cst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "tensor-core weight upload failed: %s", cudaGetErrorString(cst));
  }
  CU_TRY(e, cudaDeviceSynchronize());
  e->finalized = true;
  return CLAIRB_OK;
}

// body of clairb_predict_device; the caller holds run_mu
static int predict_device_locked(clairb_engine* e, const void* x_dev, int dtype, int64_t n, float* out_dev, void* stream) {
  if (!e->finalized) return fail(e, CLAIRB_EINVAL, "predict before clairb_finalize_weights");
  if (!x_dev || !out_dev || n <= 0 || n > e->max_sites) return fail(e, CLAIRB_EINVAL, "predict_device: bad n or buffers");
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(e, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  CU_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t eb = elem_bytes(dtype);
  int64_t done = 0;
  int nchunks = 0;
  while (done < n) {
    int64_t cn = n - done < e->chunk_sites ? n - done : e->chunk_sites;
    SiteMap sm = make_map(e, cn);
    int rc = forward_chunk(e, (const char*)x_dev + (size_t)done * SITE_ELEMS * eb, dtype, sm,
                           out_dev + (size_t)done * N_OUT, st);
    if (rc) return rc;
    e->last_map = sm;
    done += cn;
    ++nchunks;
  }
  e->last_single_chunk = nchunks == 1;
  // the workspace is shared with the host-buffer pipeline, whose own stream must not start before this work is done
  CU_TRY(e, cudaEventRecord(e->ev_dev, st));
  e->dev_pending = true;
  return CLAIRB_OK;
}

int clairb_predict_device(clairb_engine* e, const void* x_dev, int dtype, int64_t n, float* out_dev, void* stream) {
  if (!e) return CLAIRB_EINVAL;
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  return predict_device_locked(e, x_dev, dtype, n, out_dev, stream);
}

// Shared body of clairb_predict / clairb_predict_decide: chunked, copy-overlapped forward; when `ref_host` / `dec_host`
// are given the decision kernel runs on each chunk right behind the heads, on the probabilities and the input tensor
// that are still resident, and its 24-byte records travel back with the probabilities.
// `outs` (optional, instead of out_host): four head arrays [n][21], [n][3], [n][33], [n][33] as Clair.predict returns
// them; the heads kernel then writes head-major chunk buffers and the host only copies contiguous blocks.
// `out_dev_final` (instead of out_host / outs): the packed [n][90] rows stay on the device (clairb_predict_to_device).
static int predict_impl(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_host, const uint8_t* ref_host,
                        int32_t* dec_host, float* const* outs = nullptr, float* out_dev_final = nullptr) {
  if (!e) return CLAIRB_EINVAL;
  if (!e->finalized) return fail(e, CLAIRB_EINVAL, "predict before clairb_finalize_weights");
  if (outs && (!outs[0] || !outs[1] || !outs[2] || !outs[3])) return fail(e, CLAIRB_EINVAL, "predict_split: bad buffers");
  if (outs) out_host = outs[0];
  if (out_dev_final) out_host = out_dev_final;
  if (!x_host || !out_host || n <= 0 || n > e->max_sites) return fail(e, CLAIRB_EINVAL, "predict: bad n or buffers");
  // head-major chunk buffers need the kernel that can write them (tensor-core heads); the cross-check engines produce
  // packed rows and the host scatters them
  const bool dev_split = outs && e->kind == ENGINE_TC && e->fuse_tail;
  auto deliver = [&](const float* staged, int64_t at, int64_t cnt) {      // staged chunk -> the caller's array(s)
    if (!outs) {
      memcpy(out_host + (size_t)at * N_OUT, staged, (size_t)cnt * N_OUT * sizeof(float));
    } else if (dev_split) {
      for (int k = 0; k < 4; ++k)
        memcpy(outs[k] + (size_t)at * kHeadSize[k], staged + (size_t)cnt * kHeadOffH[k], (size_t)cnt * kHeadSize[k] * sizeof(float));
    } else {
      for (int64_t r = 0; r < cnt; ++r)
        for (int k = 0; k < 4; ++k)
          memcpy(outs[k] + (size_t)(at + r) * kHeadSize[k], staged + (size_t)r * N_OUT + kHeadOffH[k], kHeadSize[k] * sizeof(float));
    }
  };
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(e, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  CU_TRY(e, cudaSetDevice(e->device));
  if (e->dev_pending) {
    CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_dev, 0));
    e->dev_pending = false;
  }
  const size_t eb = elem_bytes(dtype);
  // Results go device -> pinned staging (truly asynchronous) -> the caller's array.  A pageable destination would make
  // every D2H copy synchronous and serialise the chunk pipeline, and the reference contract hands back a fresh numpy
  // array per call (clair/model.py:963), i.e. pageable memory.  A destination that is itself pinned is written directly.
  bool out_pinned = out_dev_final != nullptr;      // nothing to stage either way
  if (!outs && !out_dev_final) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out_host) == cudaSuccess) out_pinned = at.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  int64_t done = 0, prev_done = 0, prev_cn = 0;
  int c = 0;
  while (done < n) {
    const int b = c & 1;
    int64_t cn = n - done < e->chunk_sites ? n - done : e->chunk_sites;
    // ramp-up: nothing can overlap the very first host->device copy, so the first chunk of a multi-chunk call is half a
    // chunk = exactly ONE wave of CTA pairs (a quarter chunk copies faster but leaves half the SMs idle for a whole
    // wave time: measured 9 wave-times instead of 8 on a 75,000-site call)
    if (c == 0 && n > e->chunk_sites && e->ramp) {
      cn = (e->chunk_sites / 2 + TC_PAIR_SITES - 1) / TC_PAIR_SITES * TC_PAIR_SITES;
      if (cn > e->chunk_sites) cn = e->chunk_sites;
    }
    SiteMap sm = make_map(e, cn);
    // input buffer b is free once the forward of chunk c-2 has consumed it
    CU_TRY(e, cudaStreamWaitEvent(e->s_h2d, e->ev_comp[b], 0));
    {
      const size_t bytes = (size_t)cn * SITE_ELEMS * eb;
      const char* src = (const char*)x_host + (size_t)done * SITE_ELEMS * eb;
      size_t first = bytes;
      if (e->h2d_split == 2 && bytes >= (8u << 20)) {
        first = (bytes / 2) & ~(size_t)4095;
        CU_TRY(e, cudaStreamWaitEvent(e->s_h2d2, e->ev_comp[b], 0));
        CU_TRY(e, cudaMemcpyAsync((char*)e->d_x[b] + first, src + first, bytes - first, cudaMemcpyHostToDevice, e->s_h2d2));
        CU_TRY(e, cudaEventRecord(e->ev_h2d2[b], e->s_h2d2));
        CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_h2d2[b], 0));
      }
      CU_TRY(e, cudaMemcpyAsync(e->d_x[b], src, first, cudaMemcpyHostToDevice, e->s_h2d));
    }
    if (dec_host) {
      // the staging buffer was last read by the copy of chunk c-2, which ev_h2d[b] covers
      if (c >= 2) CU_TRY(e, cudaEventSynchronize(e->ev_h2d[b]));
      memcpy(e->h_ref[b], ref_host + done, (size_t)cn);
      CU_TRY(e, cudaMemcpyAsync(e->d_ref[b], e->h_ref[b], (size_t)cn, cudaMemcpyHostToDevice, e->s_h2d));
    }
    CU_TRY(e, cudaEventRecord(e->ev_h2d[b], e->s_h2d));
    CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_h2d[b], 0));
    CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_d2h[b], 0));   // output buffer b drained
    float* fwd_out = out_dev_final ? out_dev_final + (size_t)done * N_OUT : e->d_out[b];
    int rc = forward_chunk(e, e->d_x[b], dtype, sm, fwd_out, e->s_comp, dev_split ? cn : 0);
    if (rc) return rc;
    if (dec_host) {
      ProfScope ps(e, 13, e->s_comp);
      cudaError_t dst_ = dtype == CLAIRB_DTYPE_I16
                             ? decide::launch<int16_t>(e->d_out[b], e->d_ref[b], (const int16_t*)e->d_x[b], e->d_dec[b], cn, e->s_comp, dev_split ? cn : 0)
                             : decide::launch<float>(e->d_out[b], e->d_ref[b], (const float*)e->d_x[b], e->d_dec[b], cn, e->s_comp, dev_split ? cn : 0);
      if (dst_ != cudaSuccess) return fail(e, CLAIRB_ECUDA, "decide_sites launch failed: %s", cudaGetErrorString(dst_));
      e->launches += 1;
    }
    CU_TRY(e, cudaEventRecord(e->ev_comp[b], e->s_comp));
    CU_TRY(e, cudaStreamWaitEvent(e->s_d2h, e->ev_comp[b], 0));
    float* dst = out_pinned ? out_host + (size_t)done * N_OUT : e->h_out[b];
    if (!out_dev_final)
      CU_TRY(e, cudaMemcpyAsync(dst, e->d_out[b], (size_t)cn * N_OUT * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
    if (dec_host)
      CU_TRY(e, cudaMemcpyAsync(e->h_dec[b], e->d_dec[b], (size_t)cn * decide::REC_WORDS * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, e->s_d2h));
    CU_TRY(e, cudaEventRecord(e->ev_d2h[b], e->s_d2h));
    if (!out_pinned && c >= 1) {
      // while this chunk runs, hand the previous chunk's results to the caller
      CU_TRY(e, cudaEventSynchronize(e->ev_d2h[b ^ 1]));
      deliver(e->h_out[b ^ 1], prev_done, prev_cn);
    }
    if (dec_host && c >= 1) {
      CU_TRY(e, cudaEventSynchronize(e->ev_d2h[b ^ 1]));
      memcpy(dec_host + (size_t)prev_done * decide::REC_WORDS, e->h_dec[b ^ 1], (size_t)prev_cn * decide::REC_WORDS * sizeof(int32_t));
    }
    e->last_map = sm;
    prev_done = done;
    prev_cn = cn;
    done += cn;
    ++c;
  }
  if (!out_pinned) {
    CU_TRY(e, cudaEventSynchronize(e->ev_d2h[(c - 1) & 1]));
    deliver(e->h_out[(c - 1) & 1], prev_done, prev_cn);
  }
  if (dec_host) {
    CU_TRY(e, cudaEventSynchronize(e->ev_d2h[(c - 1) & 1]));
    memcpy(dec_host + (size_t)prev_done * decide::REC_WORDS, e->h_dec[(c - 1) & 1], (size_t)prev_cn * decide::REC_WORDS * sizeof(int32_t));
  }
  e->last_single_chunk = c == 1;
  // the caller reads out_host as soon as we return (call_var.py:1334-1338)
  CU_TRY(e, cudaStreamSynchronize(e->s_d2h));
  CU_TRY(e, cudaStreamSynchronize(e->s_comp));
  return CLAIRB_OK;
}

int clairb_predict(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_host) {
  return predict_impl(e, x_host, dtype, n, out_host, nullptr, nullptr);
}

int clairb_predict_to_device(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_dev) {
  if (!e) return CLAIRB_EINVAL;
  if (!out_dev) return fail(e, CLAIRB_EINVAL, "predict_to_device: out_dev is required");
  return predict_impl(e, x_host, dtype, n, nullptr, nullptr, nullptr, nullptr, out_dev);
}

int clairb_predict_split(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_gt21, float* out_genotype,
                         float* out_indel_1, float* out_indel_2) {
  float* outs[4] = {out_gt21, out_genotype, out_indel_1, out_indel_2};
  return predict_impl(e, x_host, dtype, n, nullptr, nullptr, nullptr, outs);
}

int clairb_predict_split_decide(clairb_engine* e, const void* x_host, int dtype, int64_t n, const uint8_t* ref_base,
                                float* out_gt21, float* out_genotype, float* out_indel_1, float* out_indel_2, int32_t* decision) {
  if (!e) return CLAIRB_EINVAL;
  if (!ref_base || !decision) return fail(e, CLAIRB_EINVAL, "predict_split_decide: ref_base and decision are required");
  float* outs[4] = {out_gt21, out_genotype, out_indel_1, out_indel_2};
  return predict_impl(e, x_host, dtype, n, nullptr, ref_base, decision, outs);
}

int clairb_predict_decide(clairb_engine* e, const void* x_host, int dtype, int64_t n, const uint8_t* ref_base, float* out_host,
                          int32_t* decision) {
  if (!e) return CLAIRB_EINVAL;
  if (!ref_base || !decision) return fail(e, CLAIRB_EINVAL, "predict_decide: ref_base and decision are required");
  return predict_impl(e, x_host, dtype, n, out_host, ref_base, decision);
}


// ---- asynchronous submission: many small predict calls in flight, coalesced into full chunks --------------------------
// The reference hands predict() one 1000-site batch per call (clair/call_var.py:1340-1344, shared/param.py:16); 1000 sites
// are 8 CTA pairs on a 148-SM device.  Requests queue here and a worker thread packs the sites of consecutive requests
// into chunks ("groups") of up to chunk_sites - tiles run straight through request boundaries, sites being independent -
// and drives the same double-buffered H2D -> forward -> D2H pipeline as the synchronous call.  While a group runs, new
// requests accumulate, so the group size follows the submission rate by itself.
namespace {

struct Segment { AsyncRequest* r; int64_t at, cnt, pos; };   // rows [at, at+cnt) of r sit at rows [pos, pos+cnt) of the chunk
struct Group {
  std::vector<Segment> segs;
  int64_t cn = 0;
  int b = 0, dtype = 0;
  bool split = false, decide = false, to_device = false;
};

int issue_group(clairb_engine* e, const Group& g) {
  const int b = g.b;
  const size_t eb = elem_bytes(g.dtype);
  const SiteMap sm = make_map(e, g.cn);
  const bool dev_split = g.split && e->kind == ENGINE_TC && e->fuse_tail;
  CU_TRY(e, cudaStreamWaitEvent(e->s_h2d, e->ev_comp[b], 0));      // the forward two groups back has consumed d_x[b]
  for (const Segment& s : g.segs)
    CU_TRY(e, cudaMemcpyAsync((char*)e->d_x[b] + (size_t)s.pos * SITE_ELEMS * eb, s.r->x + (size_t)s.at * SITE_ELEMS * eb,
                              (size_t)s.cnt * SITE_ELEMS * eb, cudaMemcpyHostToDevice, e->s_h2d));
  if (g.decide) {
    CU_TRY(e, cudaEventSynchronize(e->ev_h2d[b]));                 // staging last read by the copy two groups back
    for (const Segment& s : g.segs) memcpy(e->h_ref[b] + s.pos, s.r->ref + s.at, (size_t)s.cnt);
    CU_TRY(e, cudaMemcpyAsync(e->d_ref[b], e->h_ref[b], (size_t)g.cn, cudaMemcpyHostToDevice, e->s_h2d));
  }
  CU_TRY(e, cudaEventRecord(e->ev_h2d[b], e->s_h2d));
  CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_h2d[b], 0));
  CU_TRY(e, cudaStreamWaitEvent(e->s_comp, e->ev_d2h[b], 0));
  if (int rc = forward_chunk(e, e->d_x[b], g.dtype, sm, e->d_out[b], e->s_comp, dev_split ? g.cn : 0)) return rc;
  if (g.decide) {
    ProfScope ps(e, 13, e->s_comp);
    cudaError_t st = g.dtype == CLAIRB_DTYPE_I16
                         ? decide::launch<int16_t>(e->d_out[b], e->d_ref[b], (const int16_t*)e->d_x[b], e->d_dec[b], g.cn, e->s_comp, dev_split ? g.cn : 0)
                         : decide::launch<float>(e->d_out[b], e->d_ref[b], (const float*)e->d_x[b], e->d_dec[b], g.cn, e->s_comp, dev_split ? g.cn : 0);
    if (st != cudaSuccess) return fail(e, CLAIRB_ECUDA, "decide_sites launch failed: %s", cudaGetErrorString(st));
    e->launches += 1;
  }
  CU_TRY(e, cudaEventRecord(e->ev_comp[b], e->s_comp));
  CU_TRY(e, cudaStreamWaitEvent(e->s_d2h, e->ev_comp[b], 0));
  if (g.to_device) {
    // the rows stay on the device: every segment goes straight to its request's buffer
    for (const Segment& s : g.segs)
      CU_TRY(e, cudaMemcpyAsync(s.r->out_dev + (size_t)s.at * N_OUT, e->d_out[b] + (size_t)s.pos * N_OUT, (size_t)s.cnt * N_OUT * sizeof(float),
                                cudaMemcpyDeviceToDevice, e->s_d2h));
  } else {
    CU_TRY(e, cudaMemcpyAsync(e->h_out[b], e->d_out[b], (size_t)g.cn * N_OUT * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
  }
  if (g.decide)
    CU_TRY(e, cudaMemcpyAsync(e->h_dec[b], e->d_dec[b], (size_t)g.cn * decide::REC_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_d2h));
  CU_TRY(e, cudaEventRecord(e->ev_d2h[b], e->s_d2h));
  return CLAIRB_OK;
}

// staged results of a group -> the arrays of its requests; marks requests whose last sites these were
int deliver_group(clairb_engine* e, const Group& g, int rc_issue) {
  int rc = rc_issue;
  if (!rc) {
    cudaError_t st = cudaEventSynchronize(e->ev_d2h[g.b]);
    if (st != cudaSuccess) rc = fail(e, CLAIRB_ECUDA, "asynchronous predict failed: %s", cudaGetErrorString(st));
  }
  const bool dev_split = g.split && e->kind == ENGINE_TC && e->fuse_tail;
  const float* staged = e->h_out[g.b];
  if (!rc && !g.to_device)
    for (const Segment& s : g.segs) {
      AsyncRequest* r = s.r;
      if (!g.split) {
        memcpy(r->outs[0] + (size_t)s.at * N_OUT, staged + (size_t)s.pos * N_OUT, (size_t)s.cnt * N_OUT * sizeof(float));
      } else if (dev_split) {
        for (int k = 0; k < 4; ++k)
          memcpy(r->outs[k] + (size_t)s.at * kHeadSize[k], staged + (size_t)g.cn * kHeadOffH[k] + (size_t)s.pos * kHeadSize[k],
                 (size_t)s.cnt * kHeadSize[k] * sizeof(float));
      } else {
        for (int64_t i = 0; i < s.cnt; ++i)
          for (int k = 0; k < 4; ++k)
            memcpy(r->outs[k] + (size_t)(s.at + i) * kHeadSize[k], staged + (size_t)(s.pos + i) * N_OUT + kHeadOffH[k], kHeadSize[k] * sizeof(float));
      }
      if (g.decide)
        memcpy(r->dec + (size_t)s.at * decide::REC_WORDS, e->h_dec[g.b] + (size_t)s.pos * decide::REC_WORDS,
               (size_t)s.cnt * decide::REC_WORDS * sizeof(int32_t));
    }
  {
    std::lock_guard<std::mutex> lk(e->q_mu);
    for (const Segment& s : g.segs) {
      AsyncRequest* r = s.r;
      if (rc && !r->rc) { r->rc = rc; r->err = e->err; }
      r->delivered += s.cnt;
      if (r->delivered == r->n) r->done = true;
    }
  }
  e->done_cv.notify_all();
  return rc;
}

void worker_main(clairb_engine* e) {
  cudaSetDevice(e->device);
  Group prev;
  bool have_prev = false;
  int prev_rc = 0;
  int c = 0;
  std::unique_lock<std::mutex> run_lk(e->run_mu, std::defer_lock);
  for (;;) {
    Group g;
    {
      std::unique_lock<std::mutex> lk(e->q_mu);
      e->q_cv.wait(lk, [&] { return e->stop || !e->pending.empty() || have_prev; });
      if (e->pending.empty() && !have_prev) return;                // stop, nothing outstanding
      // pack segments of consecutive requests that agree on dtype / layout / decision into one chunk
      while (!e->pending.empty() && g.cn < e->chunk_sites) {
        AsyncRequest* r = e->pending.front();
        const bool dec = r->dec != nullptr;
        const bool dev = r->out_dev != nullptr;
        if (g.segs.empty()) { g.dtype = r->dtype; g.split = r->split; g.decide = dec; g.to_device = dev; }
        else if (g.dtype != r->dtype || g.split != r->split || g.decide != dec || g.to_device != dev) break;
        const int64_t left = r->n - r->issued, room = e->chunk_sites - g.cn;
        const int64_t cnt = left < room ? left : room;
        g.segs.push_back({r, r->issued, cnt, g.cn});
        g.cn += cnt;
        r->issued += cnt;
        if (r->issued == r->n) e->pending.pop_front();
      }
    }
    if (g.segs.empty()) {                                          // queue ran dry: drain the group in flight
      deliver_group(e, prev, prev_rc);
      have_prev = false;
      cudaStreamSynchronize(e->s_comp);
      if (run_lk.owns_lock()) run_lk.unlock();
      continue;
    }
    if (!run_lk.owns_lock()) {
      run_lk.lock();
      if (e->dev_pending) {
        cudaStreamWaitEvent(e->s_comp, e->ev_dev, 0);
        e->dev_pending = false;
      }
    }
    g.b = c & 1;
    ++c;
    const int rc = issue_group(e, g);
    if (have_prev) deliver_group(e, prev, prev_rc);                // while this group runs, hand the previous one over
    prev = std::move(g);
    prev_rc = rc;
    have_prev = true;
  }
}

}  // namespace

int clairb_predict_async(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_gt21, float* out_genotype,
                         float* out_indel_1, float* out_indel_2, const uint8_t* ref_base, int32_t* decision, int64_t* ticket) {
  if (!e) return CLAIRB_EINVAL;
  // e->err belongs to the pipeline owner while requests are in flight: argument errors are reported under q_mu
  std::unique_lock<std::mutex> lk(e->q_mu);
  if (!e->finalized) return fail(e, CLAIRB_EINVAL, "predict before clairb_finalize_weights");
  if (!ticket || !x_host || !out_gt21 || n <= 0) return fail(e, CLAIRB_EINVAL, "predict_async: bad n or buffers");
  const bool split = out_genotype || out_indel_1 || out_indel_2;
  if (split && (!out_genotype || !out_indel_1 || !out_indel_2))
    return fail(e, CLAIRB_EINVAL, "predict_async: pass all four head arrays, or only the first one for packed [n,90] rows");
  if ((ref_base == nullptr) != (decision == nullptr)) return fail(e, CLAIRB_EINVAL, "predict_async: ref_base and decision go together");
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(e, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  if (e->stop) return fail(e, CLAIRB_EINVAL, "predict_async: the engine is being destroyed");
  AsyncRequest* r = new AsyncRequest();
  r->ticket = e->next_ticket++;
  r->x = (const char*)x_host;
  r->dtype = dtype;
  r->n = n;
  r->outs[0] = out_gt21; r->outs[1] = out_genotype; r->outs[2] = out_indel_1; r->outs[3] = out_indel_2;
  r->split = split;
  r->ref = ref_base;
  r->dec = decision;
  e->tickets[r->ticket] = r;
  e->pending.push_back(r);
  if (!e->worker_started) {
    e->worker = std::thread(worker_main, e);
    e->worker_started = true;
  }
  *ticket = r->ticket;
  lk.unlock();
  e->q_cv.notify_one();
  return CLAIRB_OK;
}

int clairb_predict_async_to_device(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_dev, int64_t* ticket) {
  if (!e) return CLAIRB_EINVAL;
  std::unique_lock<std::mutex> lk(e->q_mu);
  if (!e->finalized) return fail(e, CLAIRB_EINVAL, "predict before clairb_finalize_weights");
  if (!ticket || !x_host || !out_dev || n <= 0) return fail(e, CLAIRB_EINVAL, "predict_async_to_device: bad n or buffers");
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(e, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  if (e->stop) return fail(e, CLAIRB_EINVAL, "predict_async_to_device: the engine is being destroyed");
  AsyncRequest* r = new AsyncRequest();
  r->ticket = e->next_ticket++;
  r->x = (const char*)x_host;
  r->dtype = dtype;
  r->n = n;
  r->out_dev = out_dev;
  e->tickets[r->ticket] = r;
  e->pending.push_back(r);
  if (!e->worker_started) {
    e->worker = std::thread(worker_main, e);
    e->worker_started = true;
  }
  *ticket = r->ticket;
  lk.unlock();
  e->q_cv.notify_one();
  return CLAIRB_OK;
}

int clairb_predict_wait(clairb_engine* e, int64_t ticket) {
  if (!e) return CLAIRB_EINVAL;
  std::unique_lock<std::mutex> lk(e->q_mu);
  auto it = e->tickets.find(ticket);
  if (it == e->tickets.end()) return fail(e, CLAIRB_EINVAL, "predict_wait: unknown ticket %lld (already waited for?)", (long long)ticket);
  AsyncRequest* r = it->second;
  e->done_cv.wait(lk, [&] { return r->done; });
  e->tickets.erase(it);
  const int rc = r->rc;
  if (rc) e->err = r->err;
  delete r;
  return rc;
}

int clairb_decide(clairb_engine* e, const float* probs_host, const uint8_t* ref_base, const void* x_host, int dtype, int64_t n,
                  int32_t* decision) {
  if (!e) return CLAIRB_EINVAL;
  if (!probs_host || !ref_base || !decision || n <= 0) return fail(e, CLAIRB_EINVAL, "decide: bad n or buffers");
  if (x_host && dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(e, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  CU_TRY(e, cudaSetDevice(e->device));
  const size_t eb = elem_bytes(dtype);
  for (int64_t done = 0; done < n; done += e->chunk_sites) {
    const int64_t cn = n - done < e->chunk_sites ? n - done : e->chunk_sites;
    CU_TRY(e, cudaMemcpyAsync(e->d_out[0], probs_host + (size_t)done * N_OUT, (size_t)cn * N_OUT * sizeof(float),
                              cudaMemcpyHostToDevice, e->s_comp));
    CU_TRY(e, cudaMemcpyAsync(e->d_ref[0], ref_base + done, (size_t)cn, cudaMemcpyHostToDevice, e->s_comp));
    if (x_host)
      CU_TRY(e, cudaMemcpyAsync(e->d_x[0], (const char*)x_host + (size_t)done * SITE_ELEMS * eb, (size_t)cn * SITE_ELEMS * eb,
                                cudaMemcpyHostToDevice, e->s_comp));
    cudaError_t st;
    {
      ProfScope ps(e, 13, e->s_comp);
      st = (x_host && dtype == CLAIRB_DTYPE_I16)
               ? decide::launch<int16_t>(e->d_out[0], e->d_ref[0], (const int16_t*)e->d_x[0], e->d_dec[0], cn, e->s_comp)
               : decide::launch<float>(e->d_out[0], e->d_ref[0], x_host ? (const float*)e->d_x[0] : nullptr, e->d_dec[0], cn, e->s_comp);
    }
    if (st != cudaSuccess) return fail(e, CLAIRB_ECUDA, "decide_sites launch failed: %s", cudaGetErrorString(st));
    e->launches += 1;
    CU_TRY(e, cudaMemcpyAsync(decision + (size_t)done * decide::REC_WORDS, e->d_dec[0],
                              (size_t)cn * decide::REC_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_comp));
    CU_TRY(e, cudaStreamSynchronize(e->s_comp));
  }
  return CLAIRB_OK;
}

// ---- CreateTensor on the device -------------------------------------------------------------
int clairb_create_tensors(clairb_engine* e, const clairb_alignments* a, const int32_t* centers, int64_t n_centers, int flags,
                          int16_t* x_host, int32_t* meta_host) {
  if (!e) return CLAIRB_EINVAL;
  if (!a || !centers || n_centers <= 0 || n_centers > 0x7fffffff || !meta_host)
    return fail(e, CLAIRB_EINVAL, "create_tensors: bad arguments");
  if (a->n_reads < 0 || a->n_ops < 0 || a->seq_len < 0 || a->ref_len < 0 || a->n_reads > 0x7ffffff0 || a->n_ops > 0x7ffffff0 ||
      a->seq_len > 0x7ffffff0 || a->ref_len > 0x7ffffff0)
    return fail(e, CLAIRB_EINVAL, "create_tensors: a block holds at most 2^31 reads / ops / bases");
  if (a->n_reads && (!a->read_pos || !a->read_end || !a->read_op0 || !a->read_strand))
    return fail(e, CLAIRB_EINVAL, "create_tensors: read arrays missing");
  if (a->n_ops && (!a->op_ref || !a->op_qry || !a->op_len || !a->seq)) return fail(e, CLAIRB_EINVAL, "create_tensors: op arrays missing");
  if (!a->ref || a->ref_len == 0) return fail(e, CLAIRB_EINVAL, "create_tensors: empty reference sequence");
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  const int64_t R = a->n_reads, O = a->n_ops;
  // the kernel finds the reads of a window by binary search: POS ascending (a sorted BAM), centres ascending
  std::vector<int32_t> maxend((size_t)R);
  for (int64_t r = 0; r < R; ++r) {
    if (r && a->read_pos[r] < a->read_pos[r - 1]) return fail(e, CLAIRB_EINVAL, "create_tensors: reads must be sorted by POS (read %lld)", (long long)r);
    if (a->read_op0[r] < 0 || a->read_op0[r + 1] < a->read_op0[r] || a->read_op0[r + 1] > O)
      return fail(e, CLAIRB_EINVAL, "create_tensors: read_op0 is not a prefix array (read %lld)", (long long)r);
    maxend[r] = r && maxend[r - 1] > a->read_end[r] ? maxend[r - 1] : a->read_end[r];
  }
  for (int64_t i = 1; i < n_centers; ++i)
    if (centers[i] <= centers[i - 1]) return fail(e, CLAIRB_EINVAL, "create_tensors: candidate positions must be strictly ascending");
  CU_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = e->s_comp;
  const void* src[9] = {a->read_pos, a->read_end, maxend.data(), a->read_op0, a->read_strand, a->op_ref, a->op_qry, a->op_len, a->seq};
  const size_t bytes[9] = {(size_t)R * 4, (size_t)R * 4, (size_t)R * 4, (size_t)(R + 1) * 4, (size_t)R, (size_t)O * 4, (size_t)O * 4, (size_t)O * 4,
                           (size_t)a->seq_len};
  for (int i = 0; i < 9; ++i) {
    if (int rc = grow(e, e->ct_in[i], (bytes[i] ? bytes[i] : 4) + (i == 8 ? 64 : 0))) return rc;      // seq: slack for the staged word loads
    if (bytes[i] && src[i]) CU_TRY(e, cudaMemcpyAsync(e->ct_in[i].p, src[i], bytes[i], cudaMemcpyHostToDevice, st));
  }
  if (int rc = grow(e, e->ct_in[9], (size_t)a->ref_len)) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->ct_in[9].p, a->ref, (size_t)a->ref_len, cudaMemcpyHostToDevice, st));
  if (int rc = grow(e, e->ct_in[10], (size_t)n_centers * 4)) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->ct_in[10].p, centers, (size_t)n_centers * 4, cudaMemcpyHostToDevice, st));
  if (int rc = grow(e, e->ct_x, (size_t)n_centers * ct::ELEMS * sizeof(int16_t))) return rc;
  if (int rc = grow(e, e->ct_meta, (size_t)n_centers * 2 * sizeof(int32_t))) return rc;
  if (!e->ct_overflow) CU_TRY(e, cudaMalloc((void**)&e->ct_overflow, 3 * sizeof(int)));      // [overflow, bad code op, bad SEQ op]
  CU_TRY(e, cudaMemsetAsync(e->ct_overflow, 0, sizeof(int), st));
  CU_TRY(e, cudaMemsetAsync(e->ct_overflow + 1, 0xff, 2 * sizeof(int), st));
  e->ct_rows = 0;
  ct::Alignments da;
  da.read_pos = (const int32_t*)e->ct_in[0].p;  da.read_end = (const int32_t*)e->ct_in[1].p;
  da.read_maxend = (const int32_t*)e->ct_in[2].p;  da.read_op0 = (const int32_t*)e->ct_in[3].p;
  da.read_strand = (const uint8_t*)e->ct_in[4].p;  da.op_ref = (const int32_t*)e->ct_in[5].p;
  da.op_qry = (const int32_t*)e->ct_in[6].p;  da.op_len = (const int32_t*)e->ct_in[7].p;
  da.seq = (const uint8_t*)e->ct_in[8].p;  da.ref = (const uint8_t*)e->ct_in[9].p;
  da.ref_start0 = a->ref_start0;  da.ref_len = (int32_t)a->ref_len;  da.n_reads = (int32_t)R;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
  // persistent grid: 13 blocks of 4 warps (4 sites, 16.9 KB of counters) fit an SM's shared memory
  const int64_t resident = (int64_t)sms * 13, blocks = (n_centers + ct::SITES_PER_BLOCK - 1) / ct::SITES_PER_BLOCK;
  const unsigned grid = (unsigned)(blocks < resident ? blocks : resident);
  if (O) ct::validate_ops<<<(unsigned)(sms * 8), 256, 0, st>>>(da.op_qry, da.op_len, (int)O, a->seq_len, e->ct_overflow + 1);
  {
    ProfScope ps(e, 14, st);
    ct::create_tensors<<<grid, ct::THREADS, 0, st>>>(da, (const int32_t*)e->ct_in[10].p, (int)n_centers, flags,
                                                     (int16_t*)e->ct_x.p, (int32_t*)e->ct_meta.p, e->ct_overflow, e->ct_overflow + 1);
  }
  cudaError_t lst = cudaGetLastError();
  if (lst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "create_tensors launch failed: %s", cudaGetErrorString(lst));
  e->launches += O ? 2 : 1;
  int status[3] = {0, -1, -1};
  int& overflow = status[0];
  CU_TRY(e, cudaMemcpyAsync(status, e->ct_overflow, sizeof status, cudaMemcpyDeviceToHost, st));
  CU_TRY(e, cudaMemcpyAsync(meta_host, e->ct_meta.p, (size_t)n_centers * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (x_host)
    CU_TRY(e, cudaMemcpyAsync(x_host, e->ct_x.p, (size_t)n_centers * ct::ELEMS * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(e, cudaStreamSynchronize(st));
  if (status[1] >= 0) return fail(e, CLAIRB_EINVAL, "create_tensors: op %d has an unknown code", status[1]);
  if (status[2] >= 0)
    return fail(e, CLAIRB_EINVAL, "create_tensors: op %d reads past the end of its SEQ (the reference raises IndexError)", status[2]);
  if (overflow) return fail(e, CLAIRB_EINVAL, "create_tensors: a count does not fit int16 (depth above 32767)");
  e->ct_rows = n_centers;
  e->ct_subtracted = (flags & ct::F_SUBTRACT) != 0;
  return CLAIRB_OK;
}

int clairb_predict_created(clairb_engine* e, const int64_t* rows, int64_t n, float* out_host) {
  if (!e) return CLAIRB_EINVAL;
  if (!e->finalized) return fail(e, CLAIRB_EINVAL, "predict before clairb_finalize_weights");
  if (!rows || !out_host || n <= 0 || n > e->max_sites) return fail(e, CLAIRB_EINVAL, "predict_created: bad n or buffers");
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  if (e->ct_rows <= 0) return fail(e, CLAIRB_EINVAL, "predict_created: no tensor block resident (call clairb_create_tensors first)");
  if (!e->ct_subtracted)
    return fail(e, CLAIRB_EINVAL, "predict_created: the resident block holds raw counts; create it with CLAIRB_CT_SUBTRACT (clair/utils.py:96-98)");
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= e->ct_rows) return fail(e, CLAIRB_EINVAL, "predict_created: row %lld is outside the resident block", (long long)rows[i]);
  CU_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = e->s_comp;
  if (int rc = grow(e, e->ct_rows_idx, (size_t)n * sizeof(int64_t))) return rc;
  if (int rc = grow(e, e->ct_gather, (size_t)n * ct::ELEMS * sizeof(int16_t))) return rc;
  if (int rc = grow(e, e->ct_out, (size_t)n * N_OUT * sizeof(float))) return rc;
  CU_TRY(e, cudaMemcpyAsync(e->ct_rows_idx.p, rows, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(e, 15, st);
    ct::gather_rows<<<(unsigned)n, 128, 0, st>>>((const int16_t*)e->ct_x.p, (const int64_t*)e->ct_rows_idx.p, n, (int16_t*)e->ct_gather.p);
  }
  cudaError_t lst = cudaGetLastError();
  if (lst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "gather_rows launch failed: %s", cudaGetErrorString(lst));
  e->launches += 1;
  if (int rc = predict_device_locked(e, e->ct_gather.p, CLAIRB_DTYPE_I16, n, (float*)e->ct_out.p, st)) return rc;
  CU_TRY(e, cudaMemcpyAsync(out_host, e->ct_out.p, (size_t)n * N_OUT * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_TRY(e, cudaStreamSynchronize(st));
  return CLAIRB_OK;
}

int clairb_encode_sam(const char* text, int64_t text_len, int min_mq, int dcov, int32_t* state, int64_t cap_reads, int64_t cap_ops,
                      int64_t cap_bases, int32_t* read_pos, int32_t* read_end, int32_t* read_op0, uint8_t* read_strand, int32_t* op_ref,
                      int32_t* op_qry, int32_t* op_len, uint8_t* seq, int64_t* n_reads, int64_t* n_ops, int64_t* n_bases) {
  if ((!text && text_len) || text_len < 0 || !state || !n_reads || !n_ops || !n_bases)
    return fail(nullptr, CLAIRB_EINVAL, "encode_sam: bad arguments");
  const bool fill = read_pos != nullptr;
  if (fill && (!read_end || !read_op0 || !read_strand || (cap_ops && (!op_ref || !op_qry || !op_len)) || (cap_bases && !seq)))
    return fail(nullptr, CLAIRB_EINVAL, "encode_sam: output arrays missing");
  static const int threads = getenv("CLAIRB_DECODE_THREADS") ? atoi(getenv("CLAIRB_DECODE_THREADS")) : 4;
  sam::Out out{read_pos, read_end, read_op0, read_strand, op_ref, op_qry, op_len, seq};
  int64_t bad = -1;
  const int rc = sam::encode(text, text_len, min_mq, dcov, state, fill ? &out : nullptr, cap_reads, cap_ops, cap_bases, n_reads, n_ops,
                             n_bases, &bad, threads);
  switch (rc) {
    case sam::OK: return CLAIRB_OK;
    case sam::MALFORMED: return fail(nullptr, CLAIRB_EINVAL, "encode_sam: row %lld is not a SAM alignment row (needs 10 columns, integer FLAG / POS / MAPQ)", (long long)bad);
    case sam::UNSORTED: return fail(nullptr, CLAIRB_EINVAL, "encode_sam: row %lld: alignments are not coordinate-sorted", (long long)bad);
    case sam::CIGAR_BEYOND_SEQ: return fail(nullptr, CLAIRB_EINVAL, "encode_sam: read %lld: CIGAR consumes more bases than SEQ holds", (long long)bad);
    case sam::CAPACITY: return fail(nullptr, CLAIRB_EINVAL, "encode_sam: output arrays too small (%lld reads, %lld ops, %lld bases)", (long long)*n_reads, (long long)*n_ops, (long long)*n_bases);
    default: return fail(nullptr, CLAIRB_EINVAL, "encode_sam: block too large for int32 offsets: split the region");
  }
}

int clairb_blosc_decompress(const void* src, int64_t src_len, void* dst, int64_t dst_cap, int64_t* nbytes) {
  if (!src || src_len < 0 || !nbytes || dst_cap < 0) return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: bad arguments");
  blosc::Info h;
  if (blosc::info((const uint8_t*)src, src_len, &h)) return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: shorter than a Blosc header");
  *nbytes = h.nbytes;
  if (!dst) return CLAIRB_OK;
  switch (blosc::decompress((const uint8_t*)src, src_len, (uint8_t*)dst, dst_cap, nbytes)) {
    case blosc::OK: return CLAIRB_OK;
    case blosc::TRUNCATED: return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: frame is truncated (header says %lld bytes)", (long long)h.cbytes);
    case blosc::UNSUPPORTED: return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: only LZ4 / stored frames without bit shuffle are read (flags 0x%02x)", h.flags);
    case blosc::CAPACITY: return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: destination holds %lld bytes, frame needs %lld", (long long)dst_cap, (long long)h.nbytes);
    default: return fail(nullptr, CLAIRB_EINVAL, "blosc_decompress: corrupt frame");
  }
}

int clairb_format_tensor_rows(const char* ctg_name, const int64_t* positions, const char* reference, int64_t reference_len,
                              const int64_t* window_start, const int16_t* x, int64_t n, char* out, int64_t out_cap, int64_t* out_len) {
  if (!ctg_name || n < 0 || !out_len || (n && (!positions || !reference || !window_start || !x)) || reference_len < 0 || out_cap < 0)
    return fail(nullptr, CLAIRB_EINVAL, "format_tensor_rows: bad arguments");
  static const int threads = getenv("CLAIRB_DECODE_THREADS") ? atoi(getenv("CLAIRB_DECODE_THREADS")) : 4;
  if (fmt::rows(ctg_name, positions, reference, reference_len, window_start, x, n, out, out_cap, out_len, threads))
    return fail(nullptr, CLAIRB_EINVAL, "format_tensor_rows: the rows need %lld bytes, the buffer holds %lld", (long long)*out_len, (long long)out_cap);
  return CLAIRB_OK;
}

int clairb_format_vcf_rows(int64_t n, const char* ctg_blob, const int32_t* ctg_off, const int64_t* pos, const uint8_t* ref,
                           const uint8_t* alt, const int32_t* quality, const uint8_t* filter_code, const uint8_t* gt_code,
                           const int32_t* depth, const double* af, char* out, int64_t out_cap, int64_t* out_len, int64_t* row_end) {
  if (n < 0 || !out_len || out_cap < 0 || (n && (!ctg_blob || !ctg_off || !pos || !ref || !alt || !quality || !filter_code || !gt_code || !depth || !af)))
    return fail(nullptr, CLAIRB_EINVAL, "format_vcf_rows: bad arguments");
  if (out && n && !row_end) return fail(nullptr, CLAIRB_EINVAL, "format_vcf_rows: row_end is required with out");
  for (int64_t i = 0; i < n; ++i) {
    if (filter_code[i] > 2 || gt_code[i] > 5 || ctg_off[i + 1] < ctg_off[i] || !memchr(alt + 4 * i, 0, 4))
      return fail(nullptr, CLAIRB_EINVAL, "format_vcf_rows: row %lld has an unknown filter / genotype code, a negative contig length or an unterminated ALT", (long long)i);
  }
  static const int max_threads = getenv("CLAIRB_DECODE_THREADS") ? atoi(getenv("CLAIRB_DECODE_THREADS")) : 4;
  const int threads = n < 16384 ? 1 : max_threads;      // a predict-batch of 1000 rows is ~60 us of work: no thread is worth spawning
  if (fmt::vcf_rows(n, ctg_blob, ctg_off, pos, ref, alt, quality, filter_code, gt_code, depth, af, out, out_cap, out_len, row_end, threads))
    return fail(nullptr, CLAIRB_EINVAL, "format_vcf_rows: the rows need %lld bytes, the buffer holds %lld", (long long)*out_len + 1, (long long)out_cap);
  return CLAIRB_OK;
}

int clairb_decode_rows(const char* text, int64_t text_len, int64_t max_rows, int dtype, void* x_out, int32_t* info_off,
                       int64_t* rows_read, int64_t* rows_kept, int64_t* consumed) {
  if (!text || text_len < 0 || max_rows <= 0 || !x_out || !info_off || !rows_read || !rows_kept || !consumed)
    return fail(nullptr, CLAIRB_EINVAL, "decode_rows: bad arguments");
  if (text_len >= ((int64_t)1 << 31)) return fail(nullptr, CLAIRB_EINVAL, "decode_rows: text block must be < 2 GiB");
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return fail(nullptr, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  int64_t bad = -1;
  static const int threads = getenv("CLAIRB_DECODE_THREADS") ? atoi(getenv("CLAIRB_DECODE_THREADS")) : 4;
  const int rc = dtype == CLAIRB_DTYPE_I16
                     ? decode::rows<int16_t>(text, text_len, max_rows, (int16_t*)x_out, info_off, rows_read, rows_kept, consumed, &bad, threads)
                     : decode::rows<float>(text, text_len, max_rows, (float*)x_out, info_off, rows_read, rows_kept, consumed, &bad, threads);
  if (rc == 1) return fail(nullptr, CLAIRB_EINVAL, "decode_rows: row %lld is malformed (needs ctg pos seq + 1056 values, seq of >= 17 bases)", (long long)bad);
  if (rc == 2) return fail(nullptr, CLAIRB_EINVAL, "decode_rows: row %lld holds a value that is not an int16 integer", (long long)bad);
  return CLAIRB_OK;
}

int clairb_get_layer(clairb_engine* e, int layer, float* out_host, int64_t n) {
  if (!e || !out_host) return CLAIRB_EINVAL;
  if (!e->last_single_chunk || n != e->last_map.n)
    return fail(e, CLAIRB_EINVAL, "get_layer: needs a preceding single-chunk predict of exactly n sites");
  CU_TRY(e, cudaSetDevice(e->device));
  CU_TRY(e, cudaDeviceSynchronize());
  const SiteMap sm = e->last_map;
  const size_t np = (size_t)sm.np;
  if (layer == CLAIRB_LAYER_LOGITS) {
    std::vector<float> buf(np * N_OUT);
    CU_TRY(e, cudaMemcpy(buf.data(), e->d_logits, buf.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < n; ++r)
      memcpy(out_host + r * N_OUT, buf.data() + sm.padded_row(r) * N_OUT, N_OUT * sizeof(float));
    return CLAIRB_OK;
  }
  if (e->kind == ENGINE_TC && e->fuse_tail && layer == CLAIRB_LAYER_L3) {
    if (!e->d_l3T) CU_TRY(e, cudaMalloc((void**)&e->d_l3T, (size_t)e->chunk_np * L3_K * sizeof(float)));
    // the fused path keeps L3 on chip: replay the fused kernel on the retained LSTM2 tiles with the debug sink set
    cudaError_t cst = tc::dump_l3(e->tcw, e->tcws, sm.np, e->d_l4T, e->d_l3T);
    if (cst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "get_layer(L3) replay failed: %s", cudaGetErrorString(cst));
  }
  if (e->kind == ENGINE_TC && e->fuse_tail && layer == CLAIRB_LAYER_LSTM2) {
    cudaError_t cst = tc::get_lstm2(e->tcws, n, sm.np, out_host);
    if (cst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "get_layer failed: %s", cudaGetErrorString(cst));
    return CLAIRB_OK;
  }
  if (e->kind == ENGINE_TC && layer == CLAIRB_LAYER_LSTM1) {
    cudaError_t cst = tc::get_lstm1(e->tcws, n, sm.np, out_host);
    if (cst != cudaSuccess) return fail(e, CLAIRB_ECUDA, "get_layer failed: %s", cudaGetErrorString(cst));
    return CLAIRB_OK;
  }
  const float* src = nullptr;
  size_t planes = 0;
  switch (layer) {
    case CLAIRB_LAYER_LSTM1: src = e->d_h1; planes = (size_t)T_STEPS * 2 * H; break;
    case CLAIRB_LAYER_LSTM2: src = e->d_h2; planes = (size_t)T_STEPS * 2 * H; break;
    case CLAIRB_LAYER_L3: src = e->d_l3T; planes = L3_K; break;
    case CLAIRB_LAYER_L4: src = e->d_l4T; planes = L4_UNITS; break;
    default: return fail(e, CLAIRB_EINVAL, "get_layer: unknown layer %d", layer);
  }
  std::vector<float> buf(planes * np);
  CU_TRY(e, cudaMemcpy(buf.data(), src, buf.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int64_t r = 0; r < n; ++r) {
    const size_t p = (size_t)sm.padded_row(r);
    if (layer == CLAIRB_LAYER_LSTM1 || layer == CLAIRB_LAYER_LSTM2) {
      // plane index t*256+f -> out[t][r][f]
      for (int t = 0; t < T_STEPS; ++t)
        for (int f = 0; f < 2 * H; ++f)
          out_host[((size_t)t * n + r) * 2 * H + f] = buf[((size_t)t * 2 * H + f) * np + p];
    } else {
      // plane index k -> out[r][k]   (L3: k = o*256+c, i.e. [n,30,256] row-major)
      for (size_t k = 0; k < planes; ++k) out_host[(size_t)r * planes + k] = buf[k * np + p];
    }
  }
  return CLAIRB_OK;
}

int clairb_set_profiling(clairb_engine* e, int enabled) {
  if (!e) return CLAIRB_EINVAL;
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  cudaSetDevice(e->device);
  prof_fold(e);
  e->prof_acc.clear();
  e->profiling = enabled != 0;
  return CLAIRB_OK;
}

int clairb_read_profile(clairb_engine* e, char* json, int64_t json_len) {
  if (!e || !json || json_len < 64) return CLAIRB_EINVAL;
  std::lock_guard<std::mutex> run_lk(e->run_mu);
  cudaSetDevice(e->device);
  prof_fold(e);
  std::string s = "[";
  for (auto& kv : e->prof_acc) {
    char buf[256];
    const char* nm = kv.first < kNumKernelNames ? kKernelNames[kv.first] : "kernel";
    snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f}", s.size() > 1 ? ", " : "", nm,
             (long long)kv.second.second, kv.second.first);
    s += buf;
  }
  s += "]";
  if ((int64_t)s.size() + 1 > json_len) return fail(e, CLAIRB_EINVAL, "read_profile: buffer too small");
  memcpy(json, s.c_str(), s.size() + 1);
  e->prof_acc.clear();
  return CLAIRB_OK;
}

int clairb_destroy(clairb_engine* e) {
  if (!e) return CLAIRB_EINVAL;
  {
    // requests still queued are completed first (their callers may be blocked in clairb_predict_wait)
    std::unique_lock<std::mutex> lk(e->q_mu);
    e->stop = true;
  }
  e->q_cv.notify_all();
  if (e->worker_started) e->worker.join();
  for (auto& kv : e->tickets) delete kv.second;
  e->tickets.clear();
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  free_all(e);
  delete e;
  return CLAIRB_OK;
}

}  // extern "C"

// ---- training step: its own handle type and entry points ----
#include "train_engine.cuh"
