// Host orchestration + C-ABI of the training step (kernels in train_kernels.cuh).  Included by engine.cu.
//
// One clairb_trainer drives one GPU.  Parameters, gradients and the two Adam moments are flat fp32 device buffers in the
// order of clair_b200/weights.py:weight_shapes() (LSTM1 fw / bw, LSTM2 fw / bw: kernel, bias; L3/Unit_0..255: kernel, bias;
// L4; L5_1..4; the four heads), so the gradient of a data-parallel step is ONE contiguous buffer for the all-reduce, with the
// dense layers - whose gradients are complete first - at its tail.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/clair_b200.h"
#include "train_kernels.cuh"

struct clairb_trainer {
  int device = 0;
  int64_t max_batch = 0, np_max = 0;
  std::string err;
  int64_t launches = 0;
  cudaStream_t st = nullptr, st2 = nullptr;          // the two directions of a BiLSTM layer run side by side (st: fw, st2: bw)
  cudaStream_t st_head[4] = {};                      // the four heads (L5_k -> head k and back) side by side: [0] = st, [1] = st2
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_head[4] = {};
  struct Param { std::string name; int64_t off, rows, cols; };       // bias: rows = 1
  std::vector<Param> params;
  std::map<std::string, int> index;
  int64_t n_params = 0, dense_off = 0;                                 // dense_off: first parameter behind the LSTMs
  float *P = nullptr, *G = nullptr, *G_own = nullptr, *M1 = nullptr, *M2 = nullptr;
  uint8_t* is_kernel = nullptr;
  bool weights_set = false;
  // activations / gradients (sized for np_max sites)
  float *x_tm = nullptr, *lout[2] = {nullptr, nullptr}, *dlout[2] = {nullptr, nullptr};
  struct Dir { float *pre, *gates, *hbuf, *cbuf, *dZ; } dir[2][2] = {};      // hbuf / cbuf: 35 slabs (train_kernels.cuh)
  float *a3 = nullptr, *da3 = nullptr, *a4 = nullptr, *a4d = nullptr, *da4 = nullptr, *a5[4] = {}, *a5d[4] = {}, *da5[4] = {};
  float *zall = nullptr, *dzall = nullptr, *probs = nullptr, *target = nullptr;
  void* x_in = nullptr;
  uint8_t* mask[6] = {};                                               // lstm2 [33][n][256], l4 [n][192], l5_1..4 [n][96]
  double* d_loss = nullptr;                                            // [4 focal sums, L2 sum, gradient sum of squares]
  int64_t last_n = 0, last_np = 0;
  bool lstm_pending = false;
  bool one_sync = false;                                               // inside clairb_trainer_step / deferred mode: the parts do not synchronise
  bool deferred = false;                                               // clairb_trainer_set_deferred
  float* loss_tail = nullptr;                                          // caller-owned [4] floats: the focal sums, written by forward_backward
  float rates[6] = {0.5f, 0.5f, 0.2f, 0.2f, 0.2f, 0.2f};
};

namespace {

using namespace clairb;

std::string g_trainer_error;

int tfail(clairb_trainer* t, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (t) t->err = buf; else g_trainer_error = buf;
  return code;
}
#define TR_TRY(t, call)                                                                              \
  do {                                                                                               \
    cudaError_t _st = (call);                                                                        \
    if (_st != cudaSuccess)                                                                          \
      return tfail((t), _st == cudaErrorMemoryAllocation ? CLAIRB_ENOMEM : CLAIRB_ECUDA, "%s failed: %s (%s:%d)", #call, \
                   cudaGetErrorString(_st), __FILE__, __LINE__);                                     \
  } while (0)

inline unsigned blocks_for(int64_t count, int threads = 256) { return (unsigned)((count + threads - 1) / threads); }
// one cluster of SEQ_CTAS CTAs per `rows` sites (lstm_seq_forward<rows> / lstm_seq_backward<rows>)
inline unsigned seq_grid(int64_t np, int rows = clairb::train::SEQ_ROWS) { return (unsigned)((np + rows - 1) / rows * clairb::train::SEQ_CTAS); }
// Sites per cluster of the sequence kernels.  Forward: always 32 - two CTAs per SM, one computes while the other waits for its
// exchange (measured at 512 and 2048 sites: 3.00 / 8.45 ms per step against 3.26 / 8.83 with 64).  Backward: 32 while both
// directions' clusters fit the machine at two CTAs per SM, 64 beyond (its [64 x 128] weight tile is loaded per CTA and its
// fragments are reused over fewer sites at 32: 8.63 against 8.45 ms at 2048 sites).
inline int seq_rows_backward(int64_t np) { return seq_grid(np, 32) * 2 <= 2 * 148 ? 32 : 64; }

void trainer_layout(clairb_trainer* t) {
  const int head_n[4] = {21, 3, 33, 33};
  const char* head_names[4] = {"Y_base_change_logits", "Y_genotype_logits", "Y_indel_length_logits_1", "Y_indel_length_logits_2"};
  int64_t off = 0;
  auto add = [&](const std::string& name, int64_t rows, int64_t cols) {
    t->index[name] = (int)t->params.size();
    t->params.push_back({name, off, rows, cols});
    off += rows * cols;
  };
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      char buf[256];
      snprintf(buf, sizeof buf, "LSTM%d/stack_bidirectional_rnn/cell_0/bidirectional_rnn/%s/cudnn_compatible_lstm_cell/", l + 1, d ? "bw" : "fw");
      add(std::string(buf) + "kernel", (l ? 2 * H : F_IN) + H, G4);
      add(std::string(buf) + "bias", 1, G4);
    }
  t->dense_off = off;
  for (int c = 0; c < 2 * H; ++c) {
    add("L3/Unit_" + std::to_string(c) + "/kernel", T_STEPS, L3_UNITS);
    add("L3/Unit_" + std::to_string(c) + "/bias", 1, L3_UNITS);
  }
  add("L4/kernel", L3_K, L4_UNITS);
  add("L4/bias", 1, L4_UNITS);
  for (int k = 0; k < 4; ++k) {
    add("L5_" + std::to_string(k + 1) + "/kernel", L4_UNITS, L5_UNITS);
    add("L5_" + std::to_string(k + 1) + "/bias", 1, L5_UNITS);
  }
  for (int k = 0; k < 4; ++k) {
    add(std::string("Prediction/") + head_names[k] + "/kernel", L5_UNITS, head_n[k]);
    add(std::string("Prediction/") + head_names[k] + "/bias", 1, head_n[k]);
  }
  t->n_params = off;
}

const clairb_trainer::Param& tp(clairb_trainer* t, const std::string& name) { return t->params[t->index.at(name)]; }
std::string lstm_prefix(int l, int d) {
  char buf[256];
  snprintf(buf, sizeof buf, "LSTM%d/stack_bidirectional_rnn/cell_0/bidirectional_rnn/%s/cudnn_compatible_lstm_cell/", l + 1, d ? "bw" : "fw");
  return buf;
}

void trainer_free(clairb_trainer* t) {
  cudaSetDevice(t->device);
  auto drop = [](auto*& p) { cudaFree(p); p = nullptr; };
  drop(t->P); drop(t->G_own); drop(t->M1); drop(t->M2); drop(t->is_kernel);
  drop(t->x_tm); drop(t->lout[0]); drop(t->lout[1]); drop(t->dlout[0]); drop(t->dlout[1]);
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      auto& q = t->dir[l][d];
      drop(q.pre); drop(q.gates); drop(q.hbuf); drop(q.cbuf); drop(q.dZ);
    }
  drop(t->a3); drop(t->da3); drop(t->a4); drop(t->a4d); drop(t->da4);
  for (int k = 0; k < 4; ++k) { drop(t->a5[k]); drop(t->a5d[k]); drop(t->da5[k]); }
  drop(t->zall); drop(t->dzall); drop(t->probs); drop(t->target); drop(t->d_loss);
  cudaFree(t->x_in); t->x_in = nullptr;
  for (int i = 0; i < 6; ++i) drop(t->mask[i]);
  if (t->st) cudaStreamDestroy(t->st);
  if (t->st2) cudaStreamDestroy(t->st2);
  for (int k = 2; k < 4; ++k)
    if (t->st_head[k]) cudaStreamDestroy(t->st_head[k]);
  for (int k = 0; k < 4; ++k)
    if (t->ev_head[k]) cudaEventDestroy(t->ev_head[k]);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
}

// backward through the two directions of one BiLSTM layer: dlout[l] (gradient of the layer's output) -> parameter gradients,
// and (layer 2 only) the gradient of the layer's input accumulated into dlout[0]
int lstm_layer_backward(clairb_trainer* t, int l, int64_t np) {
  using namespace clairb::train;
  cudaStream_t st = t->st;
  const int K = l ? 2 * H : F_IN;
  const int64_t rows = (int64_t)T_STEPS * np;
  const float* in = l ? t->lout[0] : t->x_tm;
  TR_TRY(t, cudaEventRecord(t->ev_fork, st));
  TR_TRY(t, cudaStreamWaitEvent(t->st2, t->ev_fork, 0));
  for (int d = 0; d < 2; ++d) {
    cudaStream_t sd = d ? t->st2 : st;
    auto& q = t->dir[l][d];
    const auto& pk = tp(t, lstm_prefix(l, d) + "kernel");
    const auto& pb = tp(t, lstm_prefix(l, d) + "bias");
    if (seq_rows_backward(np) == 32)
      lstm_seq_backward<32><<<seq_grid(np, 32), 256, seq_bwd_smem(32), sd>>>(t->dlout[l], d * H, q.gates, q.cbuf, t->P + pk.off + (size_t)K * G4, q.dZ,
                                                                               t->G + pb.off, (int)np, d);
    else
      lstm_seq_backward<64><<<seq_grid(np, 64), 256, seq_bwd_smem(64), sd>>>(t->dlout[l], d * H, q.gates, q.cbuf, t->P + pk.off + (size_t)K * G4, q.dZ,
                                                                               t->G + pb.off, (int)np, d);
    ++t->launches;
    float* gk = t->G + pk.off;
    // dW_x = in^T . dZ (all steps at once, both in time order), dW_h = h_prev^T . dZ: h of the step before time t is slab t (fw) / t + 2 (bw)
    gemm(true, false, K, G4, (int)rows, in, K, q.dZ, G4, 0.f, gk, G4, sd, &t->launches, true);
    gemm(true, false, H, G4, (int)rows, q.hbuf + (d ? 2 : 0) * (size_t)np * H, H, q.dZ, G4, 0.f, gk + (size_t)K * G4, G4, sd, &t->launches, true);
    // d(input) = dZ . W_x^T, the two directions added: fw writes dlout[0] on its stream, bw adds to it behind the join
    if (l == 1 && d == 0) gemm(false, true, (int)rows, K, G4, q.dZ, G4, t->P + pk.off, G4, 0.f, t->dlout[0], K, sd, &t->launches);
  }
  TR_TRY(t, cudaEventRecord(t->ev_join, t->st2));
  TR_TRY(t, cudaStreamWaitEvent(st, t->ev_join, 0));
  if (l == 1) {
    const auto& pk = tp(t, lstm_prefix(1, 1) + "kernel");
    gemm(false, true, (int)rows, K, G4, t->dir[1][1].dZ, G4, t->P + pk.off, G4, 1.f, t->dlout[0], K, st, &t->launches);
  }
  return CLAIRB_OK;
}

}  // namespace

extern "C" {

int clairb_trainer_create(int device, int64_t max_batch, clairb_trainer** out) {
  using namespace clairb::train;
  if (!out || max_batch <= 0) return tfail(nullptr, CLAIRB_EINVAL, "clairb_trainer_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return tfail(nullptr, CLAIRB_ENODEVICE, "clairb_trainer_create: CUDA device %d not available (%d visible)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return tfail(nullptr, CLAIRB_ENODEVICE, "clairb_trainer_create: device %d is not compute capability 10.x (sm_100a only, no fallback)", device);
  clairb_trainer* t = new clairb_trainer();
  t->device = device;
  t->max_batch = max_batch;
  t->np_max = (max_batch + ROWS - 1) / ROWS * ROWS;
  trainer_layout(t);
  auto bail = [&](int rc) {
    g_trainer_error = t->err;
    trainer_free(t);
    delete t;
    return rc;
  };
  const size_t np = (size_t)t->np_max, TS = T_STEPS;
#define TC_TRY(call)                                                                 \
  do {                                                                               \
    cudaError_t _st = (call);                                                        \
    if (_st != cudaSuccess) {                                                        \
      tfail(t, CLAIRB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_st));       \
      return bail(_st == cudaErrorMemoryAllocation ? CLAIRB_ENOMEM : CLAIRB_ECUDA);  \
    }                                                                                \
  } while (0)
#define TC_ALLOC(p, count) TC_TRY(cudaMalloc((void**)&(p), (size_t)(count) * sizeof(*(p))))
  TC_TRY(cudaSetDevice(device));
  TC_TRY(cudaFuncSetAttribute(sgemm_big<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG_SMEM));
  TC_TRY(cudaFuncSetAttribute(sgemm_big<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG_SMEM));
  TC_TRY(cudaFuncSetAttribute(sgemm_big<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG_SMEM));
  TC_TRY(cudaFuncSetAttribute(sgemm_big<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG_SMEM));
  TC_TRY(cudaFuncSetAttribute(l3_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
  TC_TRY(cudaFuncSetAttribute(l3_backward_input, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
  TC_TRY(cudaFuncSetAttribute(l3_backward_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, L3_SMEM));
  TC_TRY(cudaFuncSetAttribute(lstm_seq_forward<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, seq_fwd_smem(32)));
  TC_TRY(cudaFuncSetAttribute(lstm_seq_backward<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, seq_bwd_smem(32)));
  TC_TRY(cudaFuncSetAttribute(lstm_seq_backward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, seq_bwd_smem(64)));
  TC_TRY(cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
  TC_TRY(cudaStreamCreateWithFlags(&t->st2, cudaStreamNonBlocking));
  t->st_head[0] = t->st;
  t->st_head[1] = t->st2;
  for (int k = 2; k < 4; ++k) TC_TRY(cudaStreamCreateWithFlags(&t->st_head[k], cudaStreamNonBlocking));
  for (int k = 0; k < 4; ++k) TC_TRY(cudaEventCreateWithFlags(&t->ev_head[k], cudaEventDisableTiming));
  TC_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  TC_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
  TC_ALLOC(t->P, t->n_params); TC_ALLOC(t->G_own, t->n_params); TC_ALLOC(t->M1, t->n_params); TC_ALLOC(t->M2, t->n_params);
  TC_ALLOC(t->is_kernel, t->n_params);
  t->G = t->G_own;
  TC_TRY(cudaMemset(t->M1, 0, t->n_params * sizeof(float)));
  TC_TRY(cudaMemset(t->M2, 0, t->n_params * sizeof(float)));
  TC_TRY(cudaMemset(t->G, 0, t->n_params * sizeof(float)));
  {
    std::vector<uint8_t> flag((size_t)t->n_params, 0);
    for (const auto& p : t->params)
      if (p.rows > 1 || p.name.find("bias") == std::string::npos) std::fill(flag.begin() + p.off, flag.begin() + p.off + p.rows * p.cols, 1);
    TC_TRY(cudaMemcpy(t->is_kernel, flag.data(), flag.size(), cudaMemcpyHostToDevice));
  }
  TC_ALLOC(t->x_tm, TS * np * F_IN);
  TC_TRY(cudaMalloc(&t->x_in, np * SITE_ELEMS * sizeof(float)));
  for (int l = 0; l < 2; ++l) {
    TC_ALLOC(t->lout[l], TS * np * 2 * H);
    TC_ALLOC(t->dlout[l], TS * np * 2 * H);
    for (int d = 0; d < 2; ++d) {
      auto& q = t->dir[l][d];
      TC_ALLOC(q.pre, TS * np * G4); TC_ALLOC(q.gates, TS * np * G4);
      TC_ALLOC(q.hbuf, (TS + 2) * np * H); TC_ALLOC(q.cbuf, (TS + 2) * np * H);
      TC_ALLOC(q.dZ, TS * np * G4);
    }
  }
  TC_ALLOC(t->a3, np * L3_K); TC_ALLOC(t->da3, np * L3_K);
  TC_ALLOC(t->a4, np * L4_UNITS); TC_ALLOC(t->a4d, np * L4_UNITS); TC_ALLOC(t->da4, np * L4_UNITS);
  for (int k = 0; k < 4; ++k) { TC_ALLOC(t->a5[k], np * L5_UNITS); TC_ALLOC(t->a5d[k], np * L5_UNITS); TC_ALLOC(t->da5[k], np * L5_UNITS); }
  TC_ALLOC(t->zall, np * N_OUT); TC_ALLOC(t->dzall, np * N_OUT); TC_ALLOC(t->probs, np * N_OUT); TC_ALLOC(t->target, np * N_OUT);
  TC_ALLOC(t->d_loss, 8);
  const size_t mask_count[6] = {TS * np * 2 * H, np * L4_UNITS, np * L5_UNITS, np * L5_UNITS, np * L5_UNITS, np * L5_UNITS};
  for (int i = 0; i < 6; ++i) TC_ALLOC(t->mask[i], mask_count[i]);
#undef TC_ALLOC
#undef TC_TRY
  *out = t;
  return CLAIRB_OK;
}

int64_t clairb_trainer_num_params(const clairb_trainer* t) { return t ? t->n_params : 0; }
int64_t clairb_trainer_dense_offset(const clairb_trainer* t) { return t ? t->dense_off : 0; }
int64_t clairb_trainer_kernel_launches(const clairb_trainer* t) { return t ? t->launches : 0; }
const char* clairb_trainer_last_error(const clairb_trainer* t) { return t ? t->err.c_str() : g_trainer_error.c_str(); }

int clairb_trainer_set_weight(clairb_trainer* t, const char* tf_name, const float* data, const int64_t* shape, int rank) {
  if (!t) return CLAIRB_EINVAL;
  if (!tf_name || !data || !shape || rank < 1 || rank > 2) return tfail(t, CLAIRB_EINVAL, "trainer_set_weight: bad arguments");
  auto it = t->index.find(tf_name);
  if (it == t->index.end()) return tfail(t, CLAIRB_EWEIGHTS, "trainer_set_weight: %s is not a variable of the graph", tf_name);
  const auto& p = t->params[it->second];
  const int64_t rows = rank == 2 ? shape[0] : 1, cols = rank == 2 ? shape[1] : shape[0];
  if (rows != p.rows || cols != p.cols) return tfail(t, CLAIRB_EWEIGHTS, "trainer_set_weight: %s has the wrong shape", tf_name);
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaMemcpy(t->P + p.off, data, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice));
  t->weights_set = true;
  return CLAIRB_OK;
}

// which = 0 weights, 1 gradients of the last step, 2 / 3 Adam moments
int clairb_trainer_get(clairb_trainer* t, int which, const char* tf_name, float* out, int64_t count) {
  if (!t || !tf_name || !out || which < 0 || which > 3) return CLAIRB_EINVAL;
  auto it = t->index.find(tf_name);
  if (it == t->index.end()) return tfail(t, CLAIRB_EWEIGHTS, "trainer_get: %s is not a variable of the graph", tf_name);
  const auto& p = t->params[it->second];
  if (count != p.rows * p.cols) return tfail(t, CLAIRB_EINVAL, "trainer_get: %s holds %lld values", tf_name, (long long)(p.rows * p.cols));
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaStreamSynchronize(t->st));
  const float* src = which == 0 ? t->P : which == 1 ? t->G : which == 2 ? t->M1 : t->M2;
  TR_TRY(t, cudaMemcpy(out, src + p.off, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  return CLAIRB_OK;
}

int clairb_trainer_set_grad_buffer(clairb_trainer* t, float* dev_ptr) {
  if (!t) return CLAIRB_EINVAL;
  t->G = dev_ptr ? dev_ptr : t->G_own;
  return CLAIRB_OK;
}

// Data-parallel plumbing without host synchronisation between the parts of a step.  `stream` is the CUDA stream every part is
// enqueued on (the second direction's stream is joined back into it before a part returns): a caller makes its communication
// stream wait on it and makes it wait on the collectives.  In deferred mode forward_backward / backward_lstm only enqueue (the
// `losses` argument is not written; read them with clairb_trainer_read_losses after clairb_trainer_apply, which synchronises).
// `loss_tail`: 4 caller-owned device floats that receive the focal-loss sums at the end of forward_backward - placed right
// behind the gradient buffer they are summed over the ranks by the same all-reduce.
void* clairb_trainer_stream(clairb_trainer* t) { return t ? (void*)t->st : nullptr; }

int clairb_trainer_set_deferred(clairb_trainer* t, int on, float* loss_tail) {
  if (!t) return CLAIRB_EINVAL;
  t->deferred = on != 0;
  t->loss_tail = on ? loss_tail : nullptr;
  return CLAIRB_OK;
}

int clairb_trainer_read_losses(clairb_trainer* t, double* losses) {
  if (!t || !losses) return CLAIRB_EINVAL;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaStreamSynchronize(t->st));
  TR_TRY(t, cudaMemcpy(losses, t->d_loss, 5 * sizeof(double), cudaMemcpyDeviceToHost));
  losses[4] *= 0.5;                                          // sum ||v||^2 / 2 (tf.nn.l2_loss)
  return CLAIRB_OK;
}

int clairb_trainer_set_dropout_rates(clairb_trainer* t, const float* rates6) {
  if (!t || !rates6) return CLAIRB_EINVAL;
  for (int i = 0; i < 6; ++i) {
    if (!(rates6[i] >= 0.f && rates6[i] < 1.f)) return tfail(t, CLAIRB_EINVAL, "trainer_set_dropout_rates: rates must lie in [0, 1)");
    t->rates[i] = rates6[i];
  }
  return CLAIRB_OK;
}

// Forward in training mode, loss, and the backward pass down to the gradient of the LSTM2 output: on return the gradients of
// the dense layers (flat offsets >= clairb_trainer_dense_offset) are complete - a data-parallel caller starts their all-reduce
// now - and losses[0..4] hold the four focal-loss sums and the L2 sum without lambda.  clairb_trainer_backward_lstm finishes.
int clairb_trainer_forward_backward(clairb_trainer* t, const void* x_host, int dtype, const float* y_host, int64_t n,
                                    const uint8_t* const* masks, uint64_t seed, double* losses) {
  using namespace clairb::train;
  if (!t) return CLAIRB_EINVAL;
  if (!t->weights_set) return tfail(t, CLAIRB_EINVAL, "train step before the weights were set");
  if (!x_host || !y_host || (!losses && !t->deferred) || n <= 0 || n > t->max_batch) return tfail(t, CLAIRB_EINVAL, "train step: bad n or buffers");
  if (dtype != CLAIRB_DTYPE_F32 && dtype != CLAIRB_DTYPE_I16) return tfail(t, CLAIRB_EINVAL, "unknown dtype %d", dtype);
  TR_TRY(t, cudaSetDevice(t->device));
  cudaStream_t st = t->st;
  const int64_t np = (n + ROWS - 1) / ROWS * ROWS;
  const int64_t rows = (int64_t)T_STEPS * np;
  t->last_n = n;
  t->last_np = np;
  const size_t eb = dtype == CLAIRB_DTYPE_I16 ? 2 : 4;
  // inputs (padding sites: zero tensors, zero targets -> zero loss and zero gradient)
  TR_TRY(t, cudaMemsetAsync(t->x_in, 0, (size_t)np * SITE_ELEMS * eb, st));
  TR_TRY(t, cudaMemcpyAsync(t->x_in, x_host, (size_t)n * SITE_ELEMS * eb, cudaMemcpyHostToDevice, st));
  TR_TRY(t, cudaMemsetAsync(t->target, 0, (size_t)np * N_OUT * sizeof(float), st));
  TR_TRY(t, cudaMemcpyAsync(t->target, y_host, (size_t)n * N_OUT * sizeof(float), cudaMemcpyHostToDevice, st));
  if (dtype == CLAIRB_DTYPE_I16) input_time_major<int16_t><<<blocks_for(np * SITE_ELEMS), 256, 0, st>>>((const int16_t*)t->x_in, t->x_tm, (int)np);
  else input_time_major<float><<<blocks_for(np * SITE_ELEMS), 256, 0, st>>>((const float*)t->x_in, t->x_tm, (int)np);
  ++t->launches;
  // dropout masks: the caller's (parity tests) or drawn from the seed
  const int64_t mask_count[6] = {rows * 2 * H, np * L4_UNITS, np * L5_UNITS, np * L5_UNITS, np * L5_UNITS, np * L5_UNITS};
  const int64_t mask_real[6] = {0, n * L4_UNITS, n * L5_UNITS, n * L5_UNITS, n * L5_UNITS, n * L5_UNITS};
  bool masks_forked = false;
  for (int i = 0; i < 6; ++i) {
    if (masks && masks[i]) {
      if (i == 0) {
        // [33][n][256] -> [33][np][256]
        TR_TRY(t, cudaMemsetAsync(t->mask[0], 1, (size_t)mask_count[0], st));
        TR_TRY(t, cudaMemcpy2DAsync(t->mask[0], (size_t)np * 2 * H, masks[0], (size_t)n * 2 * H, (size_t)n * 2 * H, T_STEPS, cudaMemcpyHostToDevice, st));
      } else {
        TR_TRY(t, cudaMemsetAsync(t->mask[i], 1, (size_t)mask_count[i], st));
        TR_TRY(t, cudaMemcpyAsync(t->mask[i], masks[i], (size_t)mask_real[i], cudaMemcpyHostToDevice, st));
      }
    } else {
      // drawn on a side stream while the LSTMs run; joined in front of the first consumer (the dropout behind LSTM2)
      if (!masks_forked) {
        TR_TRY(t, cudaEventRecord(t->ev_head[0], st));
        TR_TRY(t, cudaStreamWaitEvent(t->st_head[2], t->ev_head[0], 0));
        masks_forked = true;
      }
      make_mask<<<blocks_for(mask_count[i]), 256, 0, t->st_head[2]>>>(t->mask[i], mask_count[i], t->rates[i], seed, (uint64_t)i);
      ++t->launches;
    }
  }
  if (masks_forked) TR_TRY(t, cudaEventRecord(t->ev_head[0], t->st_head[2]));
  TR_TRY(t, cudaMemsetAsync(t->d_loss, 0, 8 * sizeof(double), st));
  // ---- forward ----
  for (int l = 0; l < 2; ++l) {
    const int K = l ? 2 * H : F_IN;
    const float* in = l ? t->lout[0] : t->x_tm;             // LSTM1_dropout_rate is 0 (clair/model.py:95): no mask between the layers
    TR_TRY(t, cudaEventRecord(t->ev_fork, st));
    TR_TRY(t, cudaStreamWaitEvent(t->st2, t->ev_fork, 0));
    for (int d = 0; d < 2; ++d) {
      cudaStream_t sd = d ? t->st2 : st;
      auto& q = t->dir[l][d];
      const auto& pk = tp(t, lstm_prefix(l, d) + "kernel");
      const auto& pb = tp(t, lstm_prefix(l, d) + "bias");
      gemm(false, false, (int)rows, G4, K, in, K, t->P + pk.off, G4, 0.f, q.pre, G4, sd, &t->launches);
      // the state before the first step: slab 0 (fw walks t = 0..32) and slab 34 (bw walks t = 32..0)
      TR_TRY(t, cudaMemsetAsync(q.hbuf + (d ? 34 : 0) * (size_t)np * H, 0, (size_t)np * H * sizeof(float), sd));
      TR_TRY(t, cudaMemsetAsync(q.cbuf + (d ? 34 : 0) * (size_t)np * H, 0, (size_t)np * H * sizeof(float), sd));
      lstm_seq_forward<32><<<seq_grid(np, 32), 256, seq_fwd_smem(32), sd>>>(q.pre, t->P + pk.off + (size_t)K * G4, t->P + pb.off, q.gates, q.cbuf, q.hbuf,
                                                                              t->lout[l], d * H, (int)np, d);
      ++t->launches;
    }
    TR_TRY(t, cudaEventRecord(t->ev_join, t->st2));
    TR_TRY(t, cudaStreamWaitEvent(st, t->ev_join, 0));
  }
  if (masks_forked) TR_TRY(t, cudaStreamWaitEvent(st, t->ev_head[0], 0));
  if (t->rates[0] > 0.f) {
    dropout_scale<<<blocks_for(rows * 2 * H), 256, 0, st>>>(t->lout[1], t->mask[0], 1.f / (1.f - t->rates[0]), rows * 2 * H);
    ++t->launches;
  }
  const auto& p3 = tp(t, "L3/Unit_0/kernel");
  const int l3_sites = (int)((np + 15) / 16 + 15) / 16 * 16;              // ~16 site chunks x 8 channel groups = 128 CTAs
  const dim3 l3_grid(2 * H / L3_LANES, (unsigned)((np + l3_sites - 1) / l3_sites));
  l3_forward<<<l3_grid, 256, L3_SMEM, st>>>(t->lout[1], t->P + p3.off, t->a3, (int)np, l3_sites);
  ++t->launches;
  auto alpha_ab = [](float rate, float* a, float* b) {
    const double q = 1.0 - rate, al = (double)ALPHA_DROPOUT;
    *a = (float)std::sqrt(1.0 / (q * ((1.0 - q) * al * al + 1.0)));
    *b = (float)(-(double)*a * (1.0 - q) * al);
  };
  const auto& p4k = tp(t, "L4/kernel");
  const auto& p4b = tp(t, "L4/bias");
  gemm(false, false, (int)np, L4_UNITS, L3_K, t->a3, L3_K, t->P + p4k.off, L4_UNITS, 0.f, t->a4, L4_UNITS, st, &t->launches);
  bias_selu<<<blocks_for(np * L4_UNITS), 256, 0, st>>>(t->a4, t->P + p4b.off, np, L4_UNITS);
  TR_TRY(t, cudaMemcpyAsync(t->a4d, t->a4, (size_t)np * L4_UNITS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ++t->launches;
  if (t->rates[1] > 0.f) {
    float a, b;
    alpha_ab(t->rates[1], &a, &b);
    alpha_dropout_forward<<<blocks_for(np * L4_UNITS), 256, 0, st>>>(t->a4d, t->mask[1], a, b, np * L4_UNITS);
    ++t->launches;
  }
  const int head_n[4] = {21, 3, 33, 33}, head_off[4] = {0, 21, 24, 57};
  const char* head_names[4] = {"Y_base_change_logits", "Y_genotype_logits", "Y_indel_length_logits_1", "Y_indel_length_logits_2"};
  float* zk[4];
  // the four heads are independent chains of small kernels (launch-bound): each on its own stream, forked behind L4 and joined
  // in front of the loss; the same again for their backward chains
  auto fork_heads = [&]() -> int {
    TR_TRY(t, cudaEventRecord(t->ev_fork, st));
    for (int k = 1; k < 4; ++k) TR_TRY(t, cudaStreamWaitEvent(t->st_head[k], t->ev_fork, 0));
    return CLAIRB_OK;
  };
  auto join_heads = [&]() -> int {
    for (int k = 1; k < 4; ++k) {
      TR_TRY(t, cudaEventRecord(t->ev_head[k], t->st_head[k]));
      TR_TRY(t, cudaStreamWaitEvent(st, t->ev_head[k], 0));
    }
    return CLAIRB_OK;
  };
  if (int rc = fork_heads()) return rc;
  for (int k = 0; k < 4; ++k) {
    cudaStream_t sh = t->st_head[k];
    const auto& p5k = tp(t, "L5_" + std::to_string(k + 1) + "/kernel");
    const auto& p5b = tp(t, "L5_" + std::to_string(k + 1) + "/bias");
    const auto& phk = tp(t, std::string("Prediction/") + head_names[k] + "/kernel");
    const auto& phb = tp(t, std::string("Prediction/") + head_names[k] + "/bias");
    gemm(false, false, (int)np, L5_UNITS, L4_UNITS, t->a4d, L4_UNITS, t->P + p5k.off, L5_UNITS, 0.f, t->a5[k], L5_UNITS, sh, &t->launches);
    bias_selu<<<blocks_for(np * L5_UNITS), 256, 0, sh>>>(t->a5[k], t->P + p5b.off, np, L5_UNITS);
    TR_TRY(t, cudaMemcpyAsync(t->a5d[k], t->a5[k], (size_t)np * L5_UNITS * sizeof(float), cudaMemcpyDeviceToDevice, sh));
    ++t->launches;
    if (t->rates[2 + k] > 0.f) {
      float a, b;
      alpha_ab(t->rates[2 + k], &a, &b);
      alpha_dropout_forward<<<blocks_for(np * L5_UNITS), 256, 0, sh>>>(t->a5d[k], t->mask[2 + k], a, b, np * L5_UNITS);
      ++t->launches;
    }
    // head k: its post-SELU logits live in their own [np][n_k] block of da5-sized scratch (da3 is free until the backward)
    zk[k] = t->da3 + (size_t)np * head_off[k];
    gemm(false, false, (int)np, head_n[k], L5_UNITS, t->a5d[k], L5_UNITS, t->P + phk.off, head_n[k], 0.f, zk[k], head_n[k], sh, &t->launches);
    bias_selu<<<blocks_for(np * head_n[k]), 256, 0, sh>>>(zk[k], t->P + phb.off, np, head_n[k]);
    ++t->launches;
    // side by side in zall [np][90] for the loss kernel
    TR_TRY(t, cudaMemcpy2DAsync(t->zall + head_off[k], N_OUT * sizeof(float), zk[k], head_n[k] * sizeof(float), head_n[k] * sizeof(float), (size_t)np,
                                cudaMemcpyDeviceToDevice, sh));
  }
  if (int rc = join_heads()) return rc;
  TR_TRY(t, cudaMemsetAsync(t->dzall, 0, (size_t)np * N_OUT * sizeof(float), st));      // padding sites: no gradient
  focal_loss_heads<<<blocks_for(np, 128), 128, 0, st>>>(t->zall, t->target, t->probs, t->dzall, t->d_loss, (int)n);
  sum_squares<<<256, 256, 0, st>>>(t->P, t->is_kernel, t->n_params, t->d_loss + 4);
  t->launches += 2;
  // ---- backward: heads, L5, L4, slice-dense ----
  TR_TRY(t, cudaMemsetAsync(t->G + t->dense_off, 0, (size_t)(t->n_params - t->dense_off) * sizeof(float), st));
  if (int rc = fork_heads()) return rc;
  for (int k = 0; k < 4; ++k) {
    cudaStream_t sh = t->st_head[k];
    float* dzk = t->da3 + (size_t)np * (N_OUT + head_off[k]);          // scratch behind the four logit blocks, one block per head
    float* da4k = t->da3 + (size_t)np * (2 * N_OUT + k * L4_UNITS);    // this head's share of d(a4); added up behind the join
    const auto& p5k = tp(t, "L5_" + std::to_string(k + 1) + "/kernel");
    const auto& p5b = tp(t, "L5_" + std::to_string(k + 1) + "/bias");
    const auto& phk = tp(t, std::string("Prediction/") + head_names[k] + "/kernel");
    const auto& phb = tp(t, std::string("Prediction/") + head_names[k] + "/bias");
    // gradient w.r.t. the head's pre-activation: dz (from the loss, in dzall) * selu'(z)
    TR_TRY(t, cudaMemcpy2DAsync(dzk, head_n[k] * sizeof(float), t->dzall + head_off[k], N_OUT * sizeof(float), head_n[k] * sizeof(float), (size_t)np,
                                cudaMemcpyDeviceToDevice, sh));
    selu_backward<<<blocks_for(np * head_n[k]), 256, 0, sh>>>(dzk, zk[k], np * head_n[k]);
    ++t->launches;
    gemm(true, false, L5_UNITS, head_n[k], (int)np, t->a5d[k], L5_UNITS, dzk, head_n[k], 0.f, t->G + phk.off, head_n[k], sh, &t->launches, true);
    column_sums<<<dim3(blocks_for(head_n[k], 128), 16), 128, 0, sh>>>(dzk, np, head_n[k], t->G + phb.off);
    gemm(false, true, (int)np, L5_UNITS, head_n[k], dzk, head_n[k], t->P + phk.off, head_n[k], 0.f, t->da5[k], L5_UNITS, sh, &t->launches);
    ++t->launches;
    if (t->rates[2 + k] > 0.f) {
      float a, b;
      alpha_ab(t->rates[2 + k], &a, &b);
      alpha_dropout_backward<<<blocks_for(np * L5_UNITS), 256, 0, sh>>>(t->da5[k], t->mask[2 + k], a, np * L5_UNITS);
      ++t->launches;
    }
    selu_backward<<<blocks_for(np * L5_UNITS), 256, 0, sh>>>(t->da5[k], t->a5[k], np * L5_UNITS);
    ++t->launches;
    gemm(true, false, L4_UNITS, L5_UNITS, (int)np, t->a4d, L4_UNITS, t->da5[k], L5_UNITS, 0.f, t->G + p5k.off, L5_UNITS, sh, &t->launches, true);
    column_sums<<<dim3(blocks_for(L5_UNITS, 128), 16), 128, 0, sh>>>(t->da5[k], np, L5_UNITS, t->G + p5b.off);
    ++t->launches;
    gemm(false, true, (int)np, L4_UNITS, L5_UNITS, t->da5[k], L5_UNITS, t->P + p5k.off, L5_UNITS, 0.f, da4k, L4_UNITS, sh, &t->launches);
  }
  if (int rc = join_heads()) return rc;
  sum_four<<<blocks_for(np * L4_UNITS), 256, 0, st>>>(t->da3 + (size_t)np * 2 * N_OUT, np * L4_UNITS, t->da4);
  ++t->launches;
  if (t->rates[1] > 0.f) {
    float a, b;
    alpha_ab(t->rates[1], &a, &b);
    alpha_dropout_backward<<<blocks_for(np * L4_UNITS), 256, 0, st>>>(t->da4, t->mask[1], a, np * L4_UNITS);
    ++t->launches;
  }
  selu_backward<<<blocks_for(np * L4_UNITS), 256, 0, st>>>(t->da4, t->a4, np * L4_UNITS);
  ++t->launches;
  gemm(true, false, L3_K, L4_UNITS, (int)np, t->a3, L3_K, t->da4, L4_UNITS, 0.f, t->G + p4k.off, L4_UNITS, st, &t->launches, true);
  column_sums<<<dim3(blocks_for(L4_UNITS, 128), 16), 128, 0, st>>>(t->da4, np, L4_UNITS, t->G + p4b.off);
  ++t->launches;
  gemm(false, true, (int)np, L3_K, L4_UNITS, t->da4, L4_UNITS, t->P + p4k.off, L4_UNITS, 0.f, t->da3, L3_K, st, &t->launches);
  l3_backward_input<<<l3_grid, 256, L3_SMEM, st>>>(t->da3, t->a3, t->P + p3.off, t->dlout[1], (int)np, l3_sites);
  l3_backward_weights<<<l3_grid, 352, L3_SMEM, st>>>(t->lout[1], t->da3, t->G + p3.off, (int)np, l3_sites);
  t->launches += 2;
  if (t->rates[0] > 0.f) {
    dropout_scale<<<blocks_for(rows * 2 * H), 256, 0, st>>>(t->dlout[1], t->mask[0], 1.f / (1.f - t->rates[0]), rows * 2 * H);
    ++t->launches;
  }
  if (t->loss_tail) {
    store_loss_sums<<<1, 32, 0, st>>>(t->d_loss, t->loss_tail);
    ++t->launches;
  }
  TR_TRY(t, cudaGetLastError());
  t->lstm_pending = true;
  if (t->one_sync || t->deferred) return CLAIRB_OK;
  TR_TRY(t, cudaMemcpyAsync(losses, t->d_loss, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  losses[4] *= 0.5;                                          // sum ||v||^2 / 2 (tf.nn.l2_loss)
  return CLAIRB_OK;
}

// BPTT through LSTM2 and LSTM1: completes the gradient buffer (flat offsets < clairb_trainer_dense_offset).
int clairb_trainer_backward_lstm(clairb_trainer* t) {
  if (!t) return CLAIRB_EINVAL;
  if (!t->lstm_pending) return tfail(t, CLAIRB_EINVAL, "backward_lstm: no forward_backward call is pending");
  TR_TRY(t, cudaSetDevice(t->device));
  // the LSTM block of the gradient buffer starts from zero: bias gradients and split contractions are added into it
  TR_TRY(t, cudaMemsetAsync(t->G, 0, (size_t)t->dense_off * sizeof(float), t->st));
  if (int rc = lstm_layer_backward(t, 1, t->last_np)) return rc;
  if (int rc = lstm_layer_backward(t, 0, t->last_np)) return rc;
  TR_TRY(t, cudaGetLastError());
  if (!t->one_sync && !t->deferred) TR_TRY(t, cudaStreamSynchronize(t->st));
  t->lstm_pending = false;
  return CLAIRB_OK;
}

// L2 gradient, global-norm clip, Adam (clair/model.py:689-694, 717-728).  `step` is the 1-based Adam step (bias corrections).
// grad_norm (optional) receives the global norm before clipping.
int clairb_trainer_apply(clairb_trainer* t, float learning_rate, float l2_lambda, float clip_norm, int64_t step, double* grad_norm) {
  using namespace clairb::train;
  if (!t) return CLAIRB_EINVAL;
  if (t->lstm_pending) return tfail(t, CLAIRB_EINVAL, "apply: clairb_trainer_backward_lstm has not run for the pending step");
  if (step < 1 || !(clip_norm > 0.f)) return tfail(t, CLAIRB_EINVAL, "apply: step must be >= 1 and clip_norm > 0");
  TR_TRY(t, cudaSetDevice(t->device));
  cudaStream_t st = t->st;
  if (l2_lambda != 0.f) {
    add_l2_gradient<<<blocks_for(t->n_params), 256, 0, st>>>(t->G, t->P, t->is_kernel, l2_lambda, t->n_params);
    ++t->launches;
  }
  TR_TRY(t, cudaMemsetAsync(t->d_loss + 5, 0, sizeof(double), st));
  sum_squares<<<256, 256, 0, st>>>(t->G, nullptr, t->n_params, t->d_loss + 5);
  const double lr_t = (double)learning_rate * std::sqrt(1.0 - std::pow(0.999, (double)step)) / (1.0 - std::pow(0.9, (double)step));
  adam_update<<<blocks_for(t->n_params), 256, 0, st>>>(t->P, t->G, t->M1, t->M2, t->d_loss + 5, clip_norm, (float)lr_t, t->n_params);
  t->launches += 2;
  TR_TRY(t, cudaGetLastError());
  double ss = 0.0;
  TR_TRY(t, cudaMemcpyAsync(&ss, t->d_loss + 5, sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  if (grad_norm) *grad_norm = std::sqrt(ss);
  return CLAIRB_OK;
}

// The whole optimisation step of Clair.train (clair/model.py:913-945) - the three calls above back to back on the device, ONE
// synchronisation at the end (the single-GPU path; a data-parallel caller needs the two points in between for its all-reduce).
int clairb_trainer_step(clairb_trainer* t, const void* x_host, int dtype, const float* y_host, int64_t n, const uint8_t* const* masks, uint64_t seed,
                        float learning_rate, float l2_lambda, float clip_norm, int64_t step, double* losses, double* grad_norm) {
  if (!t) return CLAIRB_EINVAL;
  if (!losses) return tfail(t, CLAIRB_EINVAL, "train step: bad n or buffers");
  t->one_sync = true;
  int rc = clairb_trainer_forward_backward(t, x_host, dtype, y_host, n, masks, seed, losses);
  if (!rc) rc = clairb_trainer_backward_lstm(t);
  t->one_sync = false;
  if (rc) { t->lstm_pending = false; return rc; }
  rc = clairb_trainer_apply(t, learning_rate, l2_lambda, clip_norm, step, grad_norm);      // synchronises
  if (rc) return rc;
  TR_TRY(t, cudaMemcpy(losses, t->d_loss, 5 * sizeof(double), cudaMemcpyDeviceToHost));
  losses[4] *= 0.5;                                          // sum ||v||^2 / 2 (tf.nn.l2_loss)
  return CLAIRB_OK;
}

// probabilities [n][90] of the last forward (training phase: with its dropout masks)
int clairb_trainer_get_probabilities(clairb_trainer* t, float* out, int64_t n) {
  if (!t || !out || n != t->last_n) return CLAIRB_EINVAL;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaMemcpy(out, t->probs, (size_t)n * N_OUT * sizeof(float), cudaMemcpyDeviceToHost));
  return CLAIRB_OK;
}

int clairb_trainer_destroy(clairb_trainer* t) {
  if (!t) return CLAIRB_EINVAL;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  trainer_free(t);
  delete t;
  return CLAIRB_OK;
}

}  // extern "C"
