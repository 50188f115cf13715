// TEST INFRASTRUCTURE.  Compiles the __host__ __device__ halves of csrc/create_tensor_kernels.cuh (read_range, fold_read,
// base_row) for the CPU so that the per-site rule the kernel applies can be checked against the reference-generated golden
// rows on a machine without a GPU (tests/test_widen_create_tensor.py, -m "not gpu").  Never linked into libclair_b200.so.
#include "../../clair_b200/csrc/create_tensor_kernels.cuh"

namespace {
struct PlainAdd {
  int* cnt;
  __host__ __device__ void add(int i) { cnt[i] += 1; }
};
}  // namespace

namespace {
struct HostStage {
  int ops[clairb::ct::STAGE_OPS];
  uint32_t words[clairb::ct::STAGE_WORDS];
  void put_op(int j, int lc) { ops[j] = lc; }
  void put_word(int w, uint32_t v) { words[w] = v; }
  void converge() const {}
  int op(int j) const { return ops[j]; }
  uint8_t byte(int i) const { return (uint8_t)(words[i >> 2] >> (8 * (i & 3))); }
};
}  // namespace

// The kernel's -DCLAIRB_CT_STAGED path: stage_read + fold_read_staged, fold_read_ops for the reads that do not fit the stage.  `seq` must be
// 4-byte aligned and carry 64 readable bytes behind its end (as the device buffer does).  staged_reads / general_reads: how many
// (site, read) pairs took which walk.
extern "C" int ct_host_sites_staged(const int32_t* read_pos, const int32_t* read_end, const int32_t* read_maxend, const int32_t* read_op0,
                                    const uint8_t* read_strand, int32_t n_reads, const int32_t* op_ref, const int32_t* op_qry,
                                    const int32_t* op_len, const uint8_t* seq, const uint8_t* ref, int32_t ref_start0, int32_t ref_len,
                                    const int32_t* centers, int32_t n_centers, int left_edge, int32_t* counts, int32_t* opened,
                                    int64_t* staged_reads, int64_t* general_reads) {
  clairb::ct::Alignments a{read_pos, read_end, read_maxend, read_op0, read_strand, op_ref, op_qry, op_len,
                           seq,      ref,      ref_start0,  ref_len,  n_reads};
  *staged_reads = *general_reads = 0;
  for (int ci = 0; ci < n_centers; ++ci) {
    PlainAdd add{counts + (size_t)ci * clairb::ct::ELEMS};
    int first, last, n = 0;
    uint8_t win[clairb::ct::N_POS];
    for (int i = 0; i < clairb::ct::N_POS; ++i) win[i] = clairb::ct::window_row(a, centers[ci], i);
    clairb::ct::read_range(a, centers[ci], left_edge != 0, first, last);
    for (int r = first; r < last; ++r) {
      HostStage st;
      clairb::ct::Staged sg;
      const int state = clairb::ct::stage_read(a, r, centers[ci], left_edge != 0, st, sg);
      if (state == 1) { clairb::ct::fold_read_staged(sg, centers[ci], read_strand[r] ? 16 : 0, win, st, add); ++*staged_reads; }
      else if (state == 2) { clairb::ct::fold_read_ops(a, r, centers[ci], left_edge != 0, win, add); ++*general_reads; }
      n += state != 0;
    }
    opened[ci] = n;
    for (int cell = 0; cell < clairb::ct::CELLS; ++cell) {
      int ch[4];
      clairb::ct::channels_from_slots(add.cnt + cell * 4, ch);
      for (int j = 0; j < 4; ++j) add.cnt[cell * 4 + j] = ch[j];
    }
  }
  return 0;
}

extern "C" int ct_host_sites(const int32_t* read_pos, const int32_t* read_end, const int32_t* read_maxend, const int32_t* read_op0,
                             const uint8_t* read_strand, int32_t n_reads, const int32_t* op_ref, const int32_t* op_qry,
                             const int32_t* op_len, const uint8_t* seq, const uint8_t* ref, int32_t ref_start0, int32_t ref_len,
                             const int32_t* centers, int32_t n_centers, int left_edge, int32_t* counts, int32_t* opened) {
  clairb::ct::Alignments a{read_pos, read_end, read_maxend, read_op0, read_strand, op_ref, op_qry, op_len,
                           seq,      ref,      ref_start0,  ref_len,  n_reads};
  for (int ci = 0; ci < n_centers; ++ci) {
    PlainAdd add{counts + (size_t)ci * clairb::ct::ELEMS};
    int first, last, n = 0;
    uint8_t win[clairb::ct::N_POS];
    for (int i = 0; i < clairb::ct::N_POS; ++i) win[i] = clairb::ct::window_row(a, centers[ci], i);
    clairb::ct::read_range(a, centers[ci], left_edge != 0, first, last);
    for (int r = first; r < last; ++r) n += clairb::ct::fold_read(a, r, centers[ci], left_edge != 0, win, add) ? 1 : 0;
    // every read outside [first, last) must be rejected by the rule itself
    static int scratch[clairb::ct::ELEMS];
    PlainAdd sink{scratch};
    for (int r = 0; r < n_reads; ++r)
      if ((r < first || r >= last) && clairb::ct::fold_read(a, r, centers[ci], left_edge != 0, win, sink)) return 1 + ci;
    opened[ci] = n;
    // event slots -> the reference's channels, as the kernel does when it writes the row
    for (int cell = 0; cell < clairb::ct::CELLS; ++cell) {
      int ch[4];
      clairb::ct::channels_from_slots(add.cnt + cell * 4, ch);
      for (int j = 0; j < 4; ++j) add.cnt[cell * 4 + j] = ch[j];
    }
  }
  return 0;
}
