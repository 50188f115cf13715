/* clair_b200.h — C-ABI of the B200-native replacement for Clair's batched forward path.
 *
 * The reference (HKU-BAL/Clair) is pure Python over TensorFlow 1.13; the path replaced here is
 *   Clair.__init__/init/restore_parameters/predict/close      reference clair/model.py:58,807,1016,946,872
 *   (graph: clair/model.py:400-622, activation: clair/selu.py:26-30)
 * i.e. everything `session.run(self.Y)` does for the "2BiLSTM" structure.  There is no FFI in the
 * reference; this header is the boundary a maintainer binds with ctypes.
 *
 * Conventions: plain pointers and sizes only; every entry point returns 0 on success and a
 * non-zero CLAIRB_E* code on failure (message via clairb_last_error); nothing here calls
 * exit() or throws.  One handle drives one GPU.  Calls may come from any host thread; the
 * synchronous calls of one handle are serialised inside the library (the reference keeps exactly
 * one predict in flight: clair/call_var.py:1340-1352), clairb_predict_async keeps many in flight.
 *
 * Tensor layouts
 *   input  x   : [n,33,8,4] row-major (position, ACGTacgt row, channel), already
 *                channel-subtracted as clair/utils.py:96-98 yields it; dtype per CLAIRB_DTYPE_*.
 *   output out : [n,90] float32 row-major = softmax probabilities of the four heads concatenated
 *                21 (gt21) | 3 (genotype) | 33 (indel length 1) | 33 (indel length 2)
 *                (clair/model.py:581-622, clair/task/main.py:10-29).
 */
#ifndef CLAIR_B200_H
#define CLAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLAIRB_OK            0
#define CLAIRB_EINVAL        1   /* bad argument / shape / state            */
#define CLAIRB_ECUDA         2   /* a CUDA runtime call or kernel failed     */
#define CLAIRB_ENOMEM        3   /* host or device allocation failed         */
#define CLAIRB_ENODEVICE     4   /* no sm_100 device visible                 */
#define CLAIRB_EWEIGHTS      5   /* unknown / missing / mis-shaped variable  */

#define CLAIRB_DTYPE_F32     0   /* what tensor_generator_from yields (clair/utils.py:84)     */
#define CLAIRB_DTYPE_I16     1   /* compact transport of the same integer counts (lossless)   */

#define CLAIRB_N_OUT         90  /* 21 + 3 + 33 + 33                                           */
#define CLAIRB_SITE_ELEMS    1056 /* 33*8*4, clair/utils.py:68-69                              */

/* stages of the forward graph whose activations can be read back for parity tests */
#define CLAIRB_LAYER_LSTM1   1   /* [33,n,256]  clair/model.py:423-430 (time-major like TF)    */
#define CLAIRB_LAYER_LSTM2   2   /* [33,n,256]  clair/model.py:443-450                          */
#define CLAIRB_LAYER_L3      3   /* [n,30,256]  clair/model.py:464-471                          */
#define CLAIRB_LAYER_L4      4   /* [n,192]     clair/model.py:482-488                          */
#define CLAIRB_LAYER_LOGITS  5   /* [n,90]      post-SELU head outputs, clair/model.py:581-618  */

typedef struct clairb_engine clairb_engine;

/* Replaces Clair.__init__ + tf.Session creation (clair/model.py:58-192).
 * device: CUDA ordinal.  max_sites: largest n any later call will pass (device workspace and
 * pinned staging are sized once from it).  batch_sites: sites per predict-batch when a call
 * carries several batches (param.predictBatchSize = 1000, shared/param.py:16); tiles never
 * straddle a batch.  Fails with CLAIRB_ENODEVICE when the device is not compute capability 10.x
 * — there is no CPU or other-architecture fallback. */
int clairb_create(int device, int64_t max_sites, int batch_sites, clairb_engine** out);

/* Replaces tf.train.Saver.restore variable assignment (clair/model.py:1016-1020).
 * tf_name is the TF-1.13 variable name (e.g. "L4/kernel", "L3/Unit_17/bias",
 * "LSTM1/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/kernel");
 * data is float32 row-major of the given shape.  Copies the data. */
int clairb_set_weight(clairb_engine* e, const char* tf_name, const float* data,
                      const int64_t* shape, int rank);

/* Checks that every variable of the forward graph was set, builds the device-side operand
 * layouts (gate-column permutation, fp16 hi/lo split, tile blocking) and uploads them. */
int clairb_finalize_weights(clairb_engine* e);

/* Replaces Clair.predict (clair/model.py:946-966): host buffers in, host buffers out.
 * x_host: n sites of dtype `dtype`; out_host: [n,90] float32.  Copies host->device, runs the
 * forward, copies device->host and returns when out_host is fully written (the caller's next
 * pipeline stage reads it immediately: clair/call_var.py:1334-1338).  n may be any value in
 * [1, max_sites]; it is processed as ceil(n/batch_sites) predict-batches, several in flight.
 * The library does not retain x_host or out_host. */
int clairb_predict(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_host);

/* The same call with the reference's own return layout: four arrays [n,21] [n,3] [n,33] [n,33]
 * (clair/model.py:963, clair/task/main.py:10-29).  The heads kernel writes head-major chunk buffers, so the
 * host side only copies contiguous blocks (splitting packed rows on the host costs as much as half the forward). */
int clairb_predict_split(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_gt21,
                         float* out_genotype, float* out_indel_1, float* out_indel_2);

/* ---- many predict calls in flight (SURVEY.md 8b "predict_async / _wait", 8d C2 ">= 8 batches in flight") -----------
 * The reference's loop hands predict() one batch of param.predictBatchSize = 1000 sites per call
 * (clair/call_var.py:1340-1344, shared/param.py:16); 1000 sites occupy 8 of the 74 CTA pairs a B200 holds.
 * clairb_predict_async queues the call and returns at once with a ticket; a worker thread owned by the handle packs the
 * sites of consecutive queued calls into full device chunks (sites are independent, tiles run through call boundaries)
 * and runs the same copy-overlapped pipeline as clairb_predict.  clairb_predict_wait blocks until every output of that
 * ticket is written and returns the call's status; results are bit-identical to the synchronous calls.
 *   outputs  : the four head arrays [n,21] [n,3] [n,33] [n,33] (as clairb_predict_split), or - with out_genotype,
 *              out_indel_1, out_indel_2 all NULL - packed [n,90] rows in out_gt21 (as clairb_predict)
 *   ref_base / decision : both NULL, or both given for the first-choice decision records (as clairb_predict_decide)
 *   n        : any value >= 1 (not bounded by max_sites)
 * The library reads x_host / ref_base and writes the outputs until clairb_predict_wait(ticket) has returned: the caller
 * keeps them alive and untouched until then.  Tickets complete in submission order; every ticket must be waited for
 * exactly once.  Pinned x_host (clairb_host_alloc) keeps the copies asynchronous.  Synchronous calls on the same handle
 * are serialised behind the queued ones. */
int clairb_predict_async(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_gt21,
                         float* out_genotype, float* out_indel_1, float* out_indel_2, const uint8_t* ref_base,
                         int32_t* decision, int64_t* ticket);
int clairb_predict_wait(clairb_engine* e, int64_t ticket);

/* Host input, device output: as clairb_predict, but the packed [n,90] rows are left in device memory at out_dev (the
 * send buffer of the multi-GPU gather) instead of being copied back.  Returns when out_dev is written. */
int clairb_predict_to_device(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_dev);

/* The queued form of it (tickets as clairb_predict_async): consecutive calls keep the copy / compute pipeline of the handle
 * full, so a gather of slice j (NCCL, the caller's business) overlaps the forward of slice j+1.  clairb_predict_wait
 * returns when the rows of that ticket are in out_dev. */
int clairb_predict_async_to_device(clairb_engine* e, const void* x_host, int dtype, int64_t n, float* out_dev,
                                   int64_t* ticket);

/* Same forward with both buffers already resident in device memory, enqueued on `stream`
 * (a cudaStream_t; NULL = legacy default stream) without synchronising the host.  Used by
 * bench.py for the device-resident number and by the multi-GPU shard path. */
int clairb_predict_device(clairb_engine* e, const void* x_dev, int dtype, int64_t n,
                          float* out_dev, void* stream);

/* ---- first-choice variant decision (the head of the reference's VCF stage) --------------
 * Replaces, per site, possible_outcome_probabilites_from (clair/call_var.py:589-690) plus the first pass of
 * output_from's selection loop (clair/call_var.py:732-760): which of the ~1.2 k outcome products is the
 * maximum, in the reference's float32 arithmetic and with its tie order (first category in elif order, then
 * list.index).  The string assembly of REF/ALT, the indel-base lookups and the rare retry iterations stay with
 * the caller (clair/call_var.py:762-929).
 *   ref_base : [n] uint8, 0..3 = A C G T: BASE2ACGT[reference_sequence[16]] (clair/call_var.py:718)
 *   decision : [n][CLAIRB_DECISION_WORDS] int32 records
 *              [0] category 0..9 in the order of output_from's flags tuple (clair/call_var.py:931-937):
 *                  reference, homo SNP, hetero SNP, homo Ins, hetero ACGT+Ins, hetero InsIns, homo Del,
 *                  hetero ACGT+Del, hetero DelDel, Ins+Del
 *              [1],[2] variant lengths: homo Ins/Del and ACGT+Ins/Del: (length, 0); InsIns / DelDel: the
 *                  (shorter, longer) tuple; Ins+Del: (deletion length, insertion length) as
 *                  hetero_InsDel_tuples_from stores it (clair/call_var.py:411-424)
 *              [3] aux: reference / SNP: the gt21 index of the label (clair/task/gt21.py:27-48);
 *                  ACGT+Ins/Del: the hetero base 0..3
 *              [4] maximum probability, float32 bits     [5] read depth
 *                  sum(x[16,:,delete] + x[16,:,reference]) (clair/call_var.py:1021-1024), float32 bits
 *              [6] quality score of that call, int(round(max(-10 log10(e) ln((1-p+1e-300)/(p+1e-300)) + 16, 0)^2))
 *                  (quality_score_from, clair/call_var.py:568-586; p = float32 product, the rest in float64)
 *              [7] supporting-read count of that call (clair/call_var.py:1100-1151), float32 bits; the allele
 *                  frequency the VCF row prints is [7] / [5] in float64, capped at 1 (:1152-1154)
 *              [6] and [7] assume the first choice stands (indel-base lookups come back non-empty, :780-929).
 */
#define CLAIRB_DECISION_WORDS 8

/* Forward + decision in one pass: as clairb_predict, and the decision kernel runs on every chunk while its
 * probabilities and input tensor are still in device memory. */
int clairb_predict_decide(clairb_engine* e, const void* x_host, int dtype, int64_t n,
                          const uint8_t* ref_base, float* out_host, int32_t* decision);

/* The same with the four-array return layout of Clair.predict (see clairb_predict_split). */
int clairb_predict_split_decide(clairb_engine* e, const void* x_host, int dtype, int64_t n,
                                const uint8_t* ref_base, float* out_gt21, float* out_genotype,
                                float* out_indel_1, float* out_indel_2, int32_t* decision);

/* Decision alone, from probabilities the caller already holds (e.g. an ensemble average,
 * clair/post_processing/ensemble.py).  x_host may be NULL (read depth is then reported as 0). */
int clairb_decide(clairb_engine* e, const float* probs_host, const uint8_t* ref_base, const void* x_host,
                  int dtype, int64_t n, int32_t* decision);

/* ---- CreateTensor on the device ------------------------------------------------------------
 * Replaces the CIGAR walk and generate_tensor of dataPrepScripts/CreateTensor.py (:283-366 and :29-65): alignments in,
 * one [33,8,4] block of int16 counts per candidate site out, the counts the reference prints as text (:57-62).
 * The host side (clair_b200/create_tensor.py) keeps what the reference does per SAM row before the walk - the
 * mapping-quality filter (:264) and the per-POS depth cap (:274-281) - and encodes the surviving reads:
 *   read_pos    [n_reads]    0-based POS, ascending (a coordinate-sorted BAM)
 *   read_end    [n_reads]    one past the last reference position covered by an M / = / X / D op
 *   read_op0    [n_reads+1]  index of each read's first op (prefix array)
 *   read_strand [n_reads]    (FLAG & 16) != 0                                   (:262)
 *   op_ref/op_qry/op_len [n_ops]  per kept CIGAR op: reference position at its start, offset of its first base in
 *               `seq` (soft clips already skipped, :290-291), and length << 2 | code with code CLAIRB_OP_M (M,=,X),
 *               CLAIRB_OP_I, CLAIRB_OP_D.  N / H / P ops advance nothing in the reference and are not encoded.
 *   seq         query bases of all reads back to back (any case, :261)
 *   ref         reference_sequence as `samtools faidx` returned it (any case, :149), its first base being the 0-based
 *               contig position ref_start0 (reference_start_0_based, :223)
 * centers: [n_centers] 1-based candidate positions, strictly ascending, already restricted to ctgStart..ctgEnd (:83) and
 *   to positions whose window starts inside `ref` (:55).
 * flags: CLAIRB_CT_LEFT_EDGE (the reference's default; clear it for --stop_consider_left_edge) |
 *        CLAIRB_CT_SUBTRACT (write channels 1..3 minus channel 0, what tensor_generator_from feeds the network,
 *        clair/utils.py:96-98, instead of the raw counts).
 * x_host: [n_centers][1056] int16 or NULL (the block then only stays resident on the device for
 *   clairb_predict_created).  meta_host: [n_centers][2] int32 = number of reads that opened the site's window (0: the
 *   reference prints no row for it) and the depth at the centre position (what --minCoverage is compared with, :55).
 * Not modelled: the 5,000,000-record memory guard (available_slots, :180,285-286). */
#define CLAIRB_OP_M 0
#define CLAIRB_OP_I 1
#define CLAIRB_OP_D 2
#define CLAIRB_CT_LEFT_EDGE 1
#define CLAIRB_CT_SUBTRACT  2

typedef struct clairb_alignments {
  const int32_t* read_pos;
  const int32_t* read_end;
  const int32_t* read_op0;
  const uint8_t* read_strand;
  int64_t n_reads;
  const int32_t* op_ref;
  const int32_t* op_qry;
  const int32_t* op_len;
  int64_t n_ops;
  const uint8_t* seq;
  int64_t seq_len;
  const uint8_t* ref;
  int64_t ref_len;
  int32_t ref_start0;
} clairb_alignments;

int clairb_create_tensors(clairb_engine* e, const clairb_alignments* a, const int32_t* centers, int64_t n_centers,
                          int flags, int16_t* x_host, int32_t* meta_host);

/* Forward over rows of the tensor block the last clairb_create_tensors call (with CLAIRB_CT_SUBTRACT) left on the
 * device: rows[i] indexes that call's centers.  The tensors never exist in host memory - the text hop between
 * CreateTensor.py and call_var.py (clair/callVarBam.py:191-200) and the host->device copy are gone.
 * out_host: [n,90] float32 as clairb_predict. */
int clairb_predict_created(clairb_engine* e, const int64_t* rows, int64_t n, float* out_host);

/* Host-only (no device): one Blosc1 frame of the reference's training / evaluation bins -> its bytes (a pickled numpy
 * array: blosc.pack_array(array, cname='lz4hc', clevel=9, shuffle=NOSHUFFLE), clair/utils.py:47-48; read back by
 * blosc.unpack_array in decompress_array, clair/utils.py:223-262).  Reads LZ4 / stored frames, split or not, byte-shuffled
 * or not.  With dst == NULL only *nbytes (the uncompressed size from the header) is reported. */
int clairb_blosc_decompress(const void* src, int64_t src_len, void* dst, int64_t dst_cap, int64_t* nbytes);

/* Host-only (no device): the text rows CreateTensor.py prints (dataPrepScripts/CreateTensor.py:57-62), one per site:
 * "<ctg_name> <position> <reference[window_start .. +33)> <1056 counts>\n" from raw counts x [n][1056] int16 (the x_host of
 * clairb_create_tensors without CLAIRB_CT_SUBTRACT).  With out == NULL only *out_len (bytes needed) is reported. */
int clairb_format_tensor_rows(const char* ctg_name, const int64_t* positions, const char* reference, int64_t reference_len,
                              const int64_t* window_start, const int16_t* x, int64_t n, char* out, int64_t out_cap,
                              int64_t* out_len);

/* Host-only (no device): the VCF rows output_with prints for reference / SNP calls (clair/call_var.py:1184-1197),
 * "<ctg>\t<pos>\t.\t<REF>\t<ALT>\t<QUAL>\t<FILTER>\t.\tGT:GQ:DP:AF\t<GT>:<QUAL>:<DP>:<AF %.4f>", n rows joined by '\n'.
 *   ctg_blob / ctg_off [n+1] : contig names back to back and their offsets      pos [n]
 *   ref [n] : one character      alt [n][4] : "X" or "X,Y", NUL-terminated
 *   filter_code [n] : 0 ".", 1 "PASS", 2 "LowQual" (filtration_value_from, :70-75)
 *   gt_code [n] : 0 "0/0", 1 "1/1", 2 "0/1", 3 "1/2", 4 "0", 5 "1" (clair/task/genotype.py:3; the haploid modes, :1164-1166)
 * row_end [n] : offset one past each row.  With out == NULL only *out_len (bytes needed, +1 for out_cap) is reported. */
int clairb_format_vcf_rows(int64_t n, const char* ctg_blob, const int32_t* ctg_off, const int64_t* pos, const uint8_t* ref,
                           const uint8_t* alt, const int32_t* quality, const uint8_t* filter_code, const uint8_t* gt_code,
                           const int32_t* depth, const double* af, char* out, int64_t out_cap, int64_t* out_len,
                           int64_t* row_end);

/* Host-only (no device): `samtools view` text -> the arrays of clairb_alignments.  Replaces what the reference does per
 * SAM row before and while it walks the CIGAR string (dataPrepScripts/CreateTensor.py:251-296): '@' rows skipped, the
 * mapping-quality filter (:264), the per-POS depth cap (:274-281) and the CIGAR grammar (:283-366).
 * state: [3] int32 carried across the blocks of one stream = previous_position, depthCap (:249-250; start both at 0) and
 *   the POS of the last kept read (start at -2147483648); updated only by a filling call.
 * With read_pos == NULL nothing is written: the call only reports n_reads / n_ops / n_bases so that the caller can size
 * the arrays (read_op0 holds n_reads + 1 entries), e.g. in pinned memory from clairb_host_alloc.
 * Rows must be complete lines; CLAIRB_EINVAL for a malformed row, unsorted rows, or a CIGAR that consumes more bases
 * than SEQ holds (the reference raises IndexError there); message via clairb_last_error(NULL). */
int clairb_encode_sam(const char* text, int64_t text_len, int min_mq, int dcov, int32_t* state,
                      int64_t cap_reads, int64_t cap_ops, int64_t cap_bases,
                      int32_t* read_pos, int32_t* read_end, int32_t* read_op0, uint8_t* read_strand,
                      int32_t* op_ref, int32_t* op_qry, int32_t* op_len, uint8_t* seq,
                      int64_t* n_reads, int64_t* n_ops, int64_t* n_bases);

/* Host-only (no device): replaces the per-row `row.split()` + `np.array(columns, float32)` of
 * tensor_generator_from (clair/utils.py:81-98) for one predict-batch of text rows (CreateTensor.py:60-65).
 * Parses complete '\n'-terminated lines of `text` until max_rows lines are read; rows whose centre base
 * seq[16] is not an IUPAC code are dropped (clair/utils.py:90), kept rows are written to x_out
 * ([kept][1056], CLAIRB_DTYPE_F32 or _I16) with channel 0 subtracted from channels 1..3
 * (clair/utils.py:96-98).  info_off: [max_rows][6] byte offsets into `text` (start, end of ctg, pos, seq) of
 * every kept row.  rows_read / rows_kept / consumed (bytes) report progress.  A malformed row is
 * CLAIRB_EINVAL (the reference raises on it); message via clairb_last_error(NULL). */
int clairb_decode_rows(const char* text, int64_t text_len, int64_t max_rows, int dtype, void* x_out,
                       int32_t* info_off, int64_t* rows_read, int64_t* rows_kept, int64_t* consumed);

/* Parity hook: activations of the most recent clairb_predict* call at one stage of the graph,
 * copied to host as float32 in the reference's own axis order (see CLAIRB_LAYER_*).
 * out_host must hold layer_elems(layer) * n floats. */
int clairb_get_layer(clairb_engine* e, int layer, float* out_host, int64_t n);

/* Pinned host memory for async staging (replaces the pageable numpy buffers of
 * clair/utils.py:85).  Buffers from here make clairb_predict's copies truly asynchronous. */
int clairb_host_alloc(void** ptr, int64_t bytes);
int clairb_host_free(void* ptr);

/* Number of kernels this library launched on the handle since creation (bench.py reports the
 * difference over the timed region as gpu_launches). */
int64_t clairb_kernel_launches(const clairb_engine* e);

/* Per-kernel device timing for the roofline report: while enabled, every kernel launch is
 * bracketed by CUDA events on its own stream.  clairb_read_profile writes a JSON array
 * [{"kernel": name, "launches": L, "ms": total}] into `json` and resets the counters. */
int clairb_set_profiling(clairb_engine* e, int enabled);
int clairb_read_profile(clairb_engine* e, char* json, int64_t json_len);

/* Static description of the build: "clair_b200 <version> sm_100a <engine>" */
const char* clairb_version(void);

/* Message of the last failure on this handle (or of the last failed clairb_create when e is
 * NULL).  The pointer stays valid until the next failing call. */
const char* clairb_last_error(const clairb_engine* e);

/* Replaces Clair.close / __del__ (clair/model.py:872,1149). */
int clairb_destroy(clairb_engine* e);

/* ---- training step ------------------------------------------------------------------------------
 * Replaces, per call of Clair.train (clair/model.py:913-945): the forward in training phase (tf.layers.dropout behind LSTM2,
 * selu.dropout_selu behind L4 / L5_k: :434-578, clair/selu.py:43-74), the focal loss of the four heads (:783-805), the L2 term
 * (:689-694), the gradients of every trainable variable (BPTT through both BiLSTMs), clip_by_global_norm(5.0) and the Adam
 * update (:717-728).  fp32 on the device.  One clairb_trainer drives one GPU; parameters, gradients and Adam moments are flat
 * buffers in the order of the TF variable list (LSTM1 fw / bw, LSTM2 fw / bw: kernel, bias; L3/Unit_0..255; L4; L5_1..4; heads).
 * A step is three calls, so that a data-parallel caller can all-reduce the gradient buffer (set_grad_buffer: a device buffer of
 * its own, e.g. a torch tensor handed to NCCL) in two pieces that overlap the backward pass:
 *   clairb_trainer_forward_backward : forward, loss, backward through heads / L5 / L4 / slice-dense; on return the gradients at
 *                                     flat offsets >= clairb_trainer_dense_offset() are complete; losses[5] = the four focal-loss
 *                                     sums and sum ||kernel||^2 / 2
 *   clairb_trainer_backward_lstm    : BPTT through LSTM2 and LSTM1 (offsets < dense_offset)
 *   clairb_trainer_apply            : g += lambda * w on kernels, clip by global norm, Adam step `step` (1-based)
 *   clairb_trainer_step             : the three above back to back, one synchronisation (single GPU)
 * masks: NULL (drawn on the device from `seed`, one stream per dropout) or six uint8 keep-masks [33][n][256], [n][192],
 * 4 x [n][96] (parity tests: TensorFlow's random stream cannot be reproduced, both sides are handed the same masks).
 * clairb_trainer_get: which = 0 weights, 1 gradients of the last step, 2 / 3 Adam moments, by TF variable name. */
typedef struct clairb_trainer clairb_trainer;
int clairb_trainer_create(int device, int64_t max_batch, clairb_trainer** out);
int clairb_trainer_set_weight(clairb_trainer* t, const char* tf_name, const float* data, const int64_t* shape, int rank);
int clairb_trainer_get(clairb_trainer* t, int which, const char* tf_name, float* out, int64_t count);
int clairb_trainer_set_grad_buffer(clairb_trainer* t, float* dev_ptr);
int clairb_trainer_set_dropout_rates(clairb_trainer* t, const float* rates6);
int64_t clairb_trainer_num_params(const clairb_trainer* t);
int64_t clairb_trainer_dense_offset(const clairb_trainer* t);
int clairb_trainer_forward_backward(clairb_trainer* t, const void* x_host, int dtype, const float* y_host, int64_t n,
                                    const uint8_t* const* masks, uint64_t seed, double* losses);
int clairb_trainer_backward_lstm(clairb_trainer* t);
int clairb_trainer_apply(clairb_trainer* t, float learning_rate, float l2_lambda, float clip_norm, int64_t step,
                         double* grad_norm);
/* The three calls above as ONE call with one device synchronisation: what Clair.train(batchX, batchY) does per batch
 * (clair/model.py:913-945) on one GPU.  losses[5] and grad_norm as above. */
int clairb_trainer_step(clairb_trainer* t, const void* x_host, int dtype, const float* y_host, int64_t n,
                        const uint8_t* const* masks, uint64_t seed, float learning_rate, float l2_lambda, float clip_norm,
                        int64_t step, double* losses, double* grad_norm);
/* Data-parallel plumbing without a host synchronisation between the parts of a step: the stream the parts are enqueued on
 * (cudaStream_t; a caller makes its communication stream wait on it, and it wait on the collectives); deferred mode, in which
 * forward_backward / backward_lstm only enqueue (losses are not written - clairb_trainer_read_losses after clairb_trainer_apply);
 * loss_tail = 4 caller-owned device floats that receive the focal-loss sums at the end of forward_backward, e.g. right behind
 * the gradient buffer so that the gradient all-reduce sums them over the ranks too. */
void* clairb_trainer_stream(clairb_trainer* t);
int clairb_trainer_set_deferred(clairb_trainer* t, int on, float* loss_tail);
int clairb_trainer_read_losses(clairb_trainer* t, double* losses);
int clairb_trainer_get_probabilities(clairb_trainer* t, float* out, int64_t n);
int64_t clairb_trainer_kernel_launches(const clairb_trainer* t);
const char* clairb_trainer_last_error(const clairb_trainer* t);
int clairb_trainer_destroy(clairb_trainer* t);

#ifdef __cplusplus
}
#endif
#endif /* CLAIR_B200_H */
