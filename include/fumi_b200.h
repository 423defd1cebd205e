/*
 * fumi_b200.h -- C ABI of libfumi_b200.so: the B200 (sm_100a) implementation of FuMI's
 * episodic inner-loop adaptation path.
 *
 * The reference (s-a-malik/fumi) is pure Python and has no FFI; its boundary for this path is
 * the Python call surface listed in SURVEY.md section 8(b).  Each entry point below names the
 * reference code it replaces (paths relative to the reference root).  The host-side mirror of
 * that call surface lives in fumi_b200/ (Python) and binds these symbols through ctypes
 * (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - Plain C types only.  Unless a parameter is marked HOST, every pointer is a DEVICE pointer
 *     owned by the caller (typically a torch tensor's data_ptr()); nothing is allocated or freed
 *     behind the caller's back except the opaque host-side sampler handle.
 *   - `stream` is a cudaStream_t passed as void*; all device work is enqueued on it and the call
 *     returns without synchronising.
 *   - Every function returns 0 on success or a negative fumi_status; the message of the last
 *     failure on the calling thread is available from fumi_last_error().  There is no CPU
 *     fallback: without a CUDA device the compute entry points fail with FUMI_ERR_CUDA.
 *   - Matrices are row-major and dense.  Arithmetic is fp32 (reference: torch float32);
 *     indices, labels and predictions are int64 (reference: torch int64).
 *   - Shapes: B tasks per call, N ways, NK support rows and NQ query rows per task (grouped by
 *     class in tuple order, as the torchmeta collate produces them), D image-feature dim,
 *     H0/H1 hidden dims of the adapted image MLP (only 256/64, the reference default
 *     --im_hid_dim, is compiled), HD = H1 + 1 (head weights + bias, fumi.py:76-79).
 */
#ifndef FUMI_B200_H_
#define FUMI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUMI_B200_ABI_VERSION 1

typedef enum {
    FUMI_OK = 0,
    FUMI_ERR_ARG = -1,         /* bad argument / unsupported shape */
    FUMI_ERR_CUDA = -2,        /* CUDA runtime error (includes: no device) */
    FUMI_ERR_UNSUPPORTED = -3, /* configuration the reference accepts but this build does not */
    FUMI_ERR_DATA = -4         /* data-dependent failure (e.g. class smaller than K+Q) */
} fumi_status;

/* Per-call description of the episode (reference flags: utils/utils.py:80-90,120-128,160-179). */
typedef struct {
    int32_t num_ways;      /* N   --num_ways                                              */
    int32_t num_support;   /* NK  = N * --num_shots                                       */
    int32_t num_query;     /* NQ  = N * query shots (--num_shots_test | int(100/N))       */
    int32_t hid0;          /* H0  --im_hid_dim[0] (256)                                   */
    int32_t hid1;          /* H1  --im_hid_dim[1] (64)                                    */
    int32_t steps;         /* --num_train_adapt_steps | --num_test_adapt_steps            */
    float step_size;       /* --step_size (inner SGD learning rate alpha)                 */
    float dropout_p;       /* --dropout in train mode, 0 in eval mode / MAML              */
    uint64_t dropout_seed; /* counter-based masks: f(seed, task, pass, layer, row, col)   */
    int64_t task_offset;   /* global index of task 0 of this call (dropout counters, multi-GPU shards) */
    int32_t first_order;   /* maml.py:177 --first_order (FuMI always 0, fumi.py:176)      */
    int32_t reserved;
} fumi_episode_cfg;

int fumi_abi_version(void);
const char* fumi_last_error(void);
/* Number of SMs of the current device (grid sizing); negative fumi_status on failure. */
int fumi_device_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense layers shared by the path.
 * fumi_linear_fwd: y[M,N] = act(x[M,K] . w[N,K]^T + bias[N])        (bias may be NULL)
 *   replaces F.linear + ReLU of: the first image layer im_net.linear0 applied to every feature
 *   row once per outer step (fumi.py:215 via MetaLinear; hoisted out of the per-task loop, see
 *   DESIGN.md "Gram form"), the hypernetwork layers (fumi.py:70-107,109-113) and AM3's
 *   image_encoder / g / h (am3.py:105-126).  act: 0 none, 1 ReLU, 2 tanh, 3 sigmoid.
 *   precision: 0 = fp32 FMA, 1 = tcgen05 3xTF32 split (fp32-accurate tensor-core path).
 * fumi_linear_wgrad: dw[N,K] (+)= dy[M,N]^T . x[M,K]; db[N] (+)= column sums of dy (db may be NULL)
 *   replaces the autograd weight gradients of the same layers in outer_loss.backward()
 *   (fumi.py:192).
 * fumi_linear_dgrad: dx[M,K] = (dy[M,N] . w[N,K]) * (gate ? gate[M,K] > 0 : 1)
 *   (ReLU backward through the hypernetwork hidden layer).
 * ---------------------------------------------------------------------------------------- */
int fumi_linear_fwd(const float* x, const float* w, const float* bias, float* y,
                    int64_t M, int64_t N, int64_t K, int32_t act, int32_t precision, void* stream);
int fumi_linear_wgrad(const float* dy, const float* x, float* dw, float* db,
                      int64_t M, int64_t N, int64_t K, int32_t accumulate, int32_t precision, void* stream);
int fumi_linear_dgrad(const float* dy, const float* w, const float* gate, float* dx,
                      int64_t M, int64_t N, int64_t K, void* stream);
/* y = y * (1 - out*out) style activation backward for act 2 (tanh): dy *= 1 - y^2, in place. */
int fumi_tanh_bwd(const float* y, float* dy, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * tcgen05 tensor-core path of the same dense layers, fp32-accurate "3xTF32":
 *   x = hi + lo with hi = tf32(x);   A.B^T ~= A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T
 * fumi_split_tf32: hi[i] = tf32-rounded x[i] (fp32 container), lo[i] = x[i] - hi[i].  Static data
 *   (the HBM feature bank) is split once; weights once per outer step.
 * fumi_transpose_split_tf32: x[R,C] -> hiT/loT [C, ldt] (ldt >= R, multiple of 4; pad zero-filled):
 *   K-major operand planes for the weight-gradient contraction over rows (dW0 = d_proj^T X).
 * fumi_gemm_tf32x3: C[M,N] (=|+=) act(A[M,K] . B[N,K]^T + bias[N]); both operands K-contiguous with
 *   leading dimensions lda/ldb (floats, multiples of 4, planes 16-byte aligned).  TMA-fed tcgen05.mma
 *   (kind::tf32, M128 N256 K8) with the accumulator in TMEM.  split_k: 0 = auto, n = split K into n
 *   slices (each slice writes its partial tile to an internal workspace and a second pass sums them in slice order:
 *   deterministic, no floating-point atomics; bias/act must then be off, as with accumulate).
 * Replaces: F.linear of the hypernetwork (fumi.py:70-107,109-113) and of im_net.linear0 over the bank
 * (fumi.py:215), and the linear0.weight gradient of outer_loss.backward() (fumi.py:192).
 * ---------------------------------------------------------------------------------------- */
int fumi_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
int fumi_transpose_split_tf32(const float* x, float* hiT, float* loT, int64_t R, int64_t C, int64_t ldt,
                              void* stream);
int fumi_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                     const float* bias, float* c, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldc, int32_t act, int32_t accumulate, int32_t split_k,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * tcgen05 "3 x fp16" contraction: same contract as fumi_gemm_tf32x3 on fp16 operand planes --
 * half the bytes per element pair (hi + lo = 4 B) and K = 16 per MMA at twice the tf32 rate.  The
 * planes hold x * 2^k with k chosen from the DEVICE scalar max|x| (fumi_absmax) so that the largest
 * element lands in [2^13, 2^14); the epilogue undoes both scales exactly.  Used for the two
 * bank-sized contractions (im_net.linear0 over the feature bank, fumi.py:215, and its weight gradient,
 * fumi.py:192); the static bank is split once.
 *   fumi_absmax(x, n, out)                 out[0] = max |x[i]|   (device scalar, stream-ordered)
 *   fumi_split_f16 / _transpose_split_f16  fp32 -> fp16 (hi, lo) planes, row-major / transposed [C, ldt], ldt % 8 == 0
 *   fumi_gemm_f16x3                        C (=|+=) act(A . B^T + bias); lda / ldb in fp16 elements, % 8 == 0
 * ---------------------------------------------------------------------------------------- */
int fumi_absmax(const float* x, int64_t n, float* out, void* stream);
int fumi_split_f16(const float* x, const float* absmax, void* hi, void* lo, int64_t n, void* stream);
int fumi_transpose_split_f16(const float* x, const float* absmax, void* hiT, void* loT, int64_t R, int64_t C,
                             int64_t ldt, void* stream);
int fumi_gemm_f16x3(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                    const float* a_absmax, const float* b_absmax, const float* bias, float* c,
                    int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                    int32_t act, int32_t accumulate, int32_t split_k, void* stream);

/* ------------------------------------------------------------------------------------------
 * Episode Gram blocks (the HBM-bound gather of the path).
 * gram[b, i, j] = <feats[row(b,i)], feats[sup_rows[b,j]]>, i over the NK support rows then the NQ
 * query rows of task b.  Each sampled feature row is read from HBM once per task.
 * Replaces the per-task re-application of the adapted first layer, F.linear(x, W0 - alpha*dW0...)
 * (fumi.py:161,178 with torchmeta's updated 'linear0.weight'): see DESIGN.md "Gram form".
 * Also the GPU-resident replacement of the loader's per-sample feature copies
 * (dataset/data.py:545,571-577): rows are gathered straight from the HBM feature bank.
 * ---------------------------------------------------------------------------------------- */
int fumi_gram(const float* feats, int64_t num_rows, int64_t D,
              const int64_t* sup_rows, const int64_t* qry_rows,
              int64_t B, int32_t NK, int32_t NQ, float* gram, void* stream);
/* Same blocks from the bank's fp16 (hi, lo) planes (fumi_split_f16 + its absmax scalar): tcgen05 kind::f16,
 * K = 16 per MMA, no split arithmetic in the kernel.  NK <= 32, NK + NQ <= 192, D % 64 == 0 (else use fumi_gram). */
int fumi_gram_f16(const void* feats_hi, const void* feats_lo, const float* absmax, int64_t num_rows, int64_t D,
                  const int64_t* sup_rows, const int64_t* qry_rows, int64_t B, int32_t NK, int32_t NQ,
                  float* gram, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused inner loop + query scoring   (fumi.py:148-185, maml.py:158-183)
 *   proj        [R, H0]   feature rows through the meta-initial first layer, no bias
 *   sup_rows    [B, NK]   row of proj for each support sample;  qry_rows [B, NQ]
 *   sup_y,qry_y           labels in [0, N)
 *   gram        [B, NK+NQ, NK] from fumi_gram
 *   b0 [H0], w1 [H1,H0], b1 [H1]  meta-initial im_net.linear0.bias / linear1.*
 *   head_table  [Rh, HD]  head initialisations: hypernetwork outputs (FuMI, fumi.py:156,198-212)
 *                         or the shared lin_final [weight|bias] (MAML, maml.py:28)
 *   head_rows   [B, N]    row of head_table initialising label i of task b; NULL = rows 0..N-1
 *                         for every task (MAML)
 * outputs
 *   logits [B,NQ,N], preds [B,NQ] (argmax, lowest index on ties: torch.max, fumi.py:180),
 *   task_loss [B] (mean CE of the task's queries, fumi.py:182), task_acc [B] (fumi.py:185,329-331)
 *   stash: NULL for inference; otherwise fumi_episode_stash_floats(cfg)*B floats that
 *          fumi_episode_bwd consumes (adapted state + per-step activations).  The adapted state
 *          (hp, W1, b1, b0 and the first-layer coefficient matrix S with W0_task = W0 - alpha S^T X)
 *          can be read back from it for parity dumps: fumi_stash_layout().
 * ---------------------------------------------------------------------------------------- */
int64_t fumi_episode_stash_floats(const fumi_episode_cfg* cfg);
typedef struct {
    int64_t per_task;   /* floats per task                                   */
    int64_t S;          /* [NK,H0]  offset of the first-layer coefficient S  */
    int64_t w1t;        /* [H0,H1]  adapted linear1.weight, TRANSPOSED       */
    int64_t b0;         /* [H0]     adapted linear0.bias                     */
    int64_t b1;         /* [H1]     adapted linear1.bias                     */
    int64_t head;       /* [N,HD]   adapted head                             */
    int64_t steps;      /* start of the per-step activation records          */
    int64_t per_step;
    /* per-step record, offsets relative to the record (4-byte words).  format 0: H0 [NK,H0] and H1 [NK,H1] as fp32.
     * format 1 (NK <= 32, the tensor-core kernels): H1 fp32; H0 as the forward's fp16 operand planes -- hi and lo
     * [NK,H0] halves with H0 = (hi + lo) * 2^-e, e = ((int32*)record)[rec_exp] -- which the backward copies straight
     * into its tiles. */
    int64_t format;
    int64_t rec_h1;
    int64_t rec_exp;    /* -1 in format 0 */
    int64_t rec_h0_hi;  /* format 0: the fp32 H0 block */
    int64_t rec_h0_lo;  /* -1 in format 0 */
} fumi_stash_layout_t;
int fumi_stash_layout(const fumi_episode_cfg* cfg, fumi_stash_layout_t* out /* HOST */);

int fumi_episode_fwd(const fumi_episode_cfg* cfg, int64_t B,
                     const float* proj, const int64_t* sup_rows, const int64_t* qry_rows,
                     const int64_t* sup_y, const int64_t* qry_y, const float* gram,
                     const float* b0, const float* w1, const float* b1,
                     const float* head_table, const int64_t* head_rows,
                     float* logits, int64_t* preds, float* task_loss, float* task_acc,
                     float* stash, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hand-written backward of the same fused step: exact second-order meta-gradient
 * (outer_loss.backward() through create_graph=True inner steps, fumi.py:165-176,190-192;
 * maml.py:173-177,188-190).  Gradients are of  loss_scale * sum_b task_loss[b].
 *   d_proj     [R, H0]   += dLoss/d proj rows (atomic scatter-add; zero it first)
 *   d_head     [B, N, HD] = dLoss/d head init of each task
 *   d_b0_parts [P, H0], d_w1_parts [P, H1*H0] (layout [H1,H0]), d_b1_parts [P, H1]:
 *       per-CTA partial sums, P = fumi_episode_bwd_parts(); reduce with fumi_reduce_parts.
 * ---------------------------------------------------------------------------------------- */
int fumi_episode_bwd_parts(void);
int fumi_episode_bwd(const fumi_episode_cfg* cfg, int64_t B,
                     const float* proj, const int64_t* sup_rows, const int64_t* qry_rows,
                     const int64_t* sup_y, const int64_t* qry_y, const float* gram,
                     const float* stash, float loss_scale,
                     float* d_proj, float* d_head,
                     float* d_b0_parts, float* d_w1_parts, float* d_b1_parts, void* stream);
/* out[n] (+)= sum_p parts[p, n]  (deterministic order) */
int fumi_reduce_parts(const float* parts, int64_t P, int64_t n, float* out, int32_t accumulate, void* stream);
/* table[rows[i], :] += src[i, :]  for i < n   (scatter-add of per-task head gradients into
 * the per-class hypernetwork-output gradient; deterministic when rows are unique) */
int fumi_scatter_add_rows(const float* src, const int64_t* rows, int64_t n, int64_t width,
                          float* table, void* stream);
/* loss = sum(task_loss)/B, acc = sum(task_acc)/B  (fumi.py:187-188) -> out[0], out[1] */
int fumi_reduce_loss_acc(const float* task_loss, const float* task_acc, int64_t B, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Outer-loop optimizer step: torch.optim.Adam(lr, weight_decay) as built by init_optim
 * (utils/utils.py:280-283): L2 term added to the gradient, bias correction with 1-based `step`.
 * decoupled != 0 gives AdamW (utils.py:289-290).  One fused launch over the flat parameter buffer.
 * ---------------------------------------------------------------------------------------- */
int fumi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay,
                   int64_t step, int32_t decoupled, void* stream);

/* ------------------------------------------------------------------------------------------
 * AM3 meta-test scoring (am3.py:159-200; utils/utils.py:302-402), eval mode.
 *   emb        [R, P]   image_encoder(feature rows)       (fumi_linear_fwd)
 *   text_proto [C, P]   g(text) per class row             (fumi_linear_fwd x2)
 *   lamda      [C]      sigmoid(h(g(text))) per class row; lamda_fixed < 0 = use it, 0 / 1 override
 *   class_rows [B, N]   class row of label i
 * outputs: dist [B,NQ,N] squared distances, preds [B,NQ] (argmin, lowest index on ties),
 *          task_loss [B] = sum over the task's queries of CE(-dist) (divide by B*NQ for the mean).
 * ---------------------------------------------------------------------------------------- */
int fumi_am3_score(const float* emb, const float* text_proto, const float* lamda,
                   const int64_t* sup_rows, const int64_t* qry_rows,
                   const int64_t* sup_y, const int64_t* qry_y, const int64_t* class_rows,
                   int64_t B, int32_t N, int32_t NK, int32_t NQ, int32_t P, int32_t lamda_fixed,
                   float* protos, float* dist, int64_t* preds, float* task_loss, void* stream);

/* Confusion counts of a flat prediction array: counts[t * N + p] += #(y == t, pred == p)  (int64 [N, N], zero it first).
 * Replaces the host pass of utils.get_preds / AM3.evaluate (utils/utils.py:323-326: sklearn accuracy_score and
 * precision_recall_fscore_support over every query prediction) by its sufficient statistics. */
int fumi_confusion_counts(const int64_t* y, const int64_t* pred, int64_t n, int32_t N, int64_t* counts, void* stream);

/* AM3 meta-training (am3.py:154-196, 215-305): backward of fumi_am3_score's loss = loss_scale * sum of the query CEs.
 *   protos / dist  outputs of fumi_am3_score for the same batch
 *   d_emb [R,P]    += d loss / d image embedding rows (atomic scatter-add; zero it first)
 *   d_tproto [B,N,P], d_lamda [B,N]  gradients wrt the class text prototype / lamda of label i of task b
 *                  (scatter them into the class tables with fumi_scatter_add_rows by class_rows)
 * fumi_dropout_apply: x[r,c] *= keep(seed, layer, r, c) / (1 - p), the counter-based mask of the episode kernels (also
 *   its own backward when applied to the gradient) -- AM3's Dropout inside g / h (am3.py:66-88) in train mode.
 * fumi_sigmoid_bwd: dy *= y (1 - y)   (lamda = sigmoid(h(.)), am3.py:125). */
int fumi_am3_bwd(const float* emb, const float* text_proto, const float* lamda,
                 const int64_t* sup_rows, const int64_t* qry_rows,
                 const int64_t* sup_y, const int64_t* qry_y, const int64_t* class_rows,
                 int64_t B, int32_t N, int32_t NK, int32_t NQ, int32_t P, int32_t lamda_fixed,
                 const float* protos, const float* dist, float loss_scale,
                 float* d_emb, float* d_tproto, float* d_lamda, void* stream);
int fumi_dropout_apply(float* x, int64_t rows, int64_t cols, uint64_t seed, uint32_t layer, float p, void* stream);
int fumi_sigmoid_bwd(const float* y, float* dy, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Episodic task sampler (HOST side, native): bit-exact restatement of
 * BatchMetaDataLoader(ClassSplitter(InatAnim(...), shuffle=True, K, Q).seed(0)) --
 * dataset/data.py:73-84,125-188,377-414 + torchmeta 1.7.0.
 * The three generator streams of the reference are explicit:
 *   py_state    HOST uint32[625]  CPython `random` MT19937 state (random.getstate()[1])
 *   torch_state HOST uint32[626]  torch CPU generator: 624 key words, then next index, then `left`
 *                                  (unpacked from torch.get_rng_state() by the host wrapper)
 *   the split's own RandomState(0) stream lives inside the handle.
 * Both external states are read and written back, so the host program's other consumers
 * (model init, torch.randperm elsewhere) stay in sequence.
 * All pointers of fumi_sampler_create / _new_iterator / _next / _plan are HOST pointers.
 * ---------------------------------------------------------------------------------------- */
typedef struct fumi_sampler fumi_sampler;
/* class_offsets[C+1], class_image_ids[class_offsets[C]]: ascending image ids of split-class c. */
int fumi_sampler_create(const int64_t* class_offsets, const int64_t* class_image_ids, int64_t C,
                        int32_t N, int32_t K, int32_t Q, fumi_sampler** out);
void fumi_sampler_destroy(fumi_sampler* s);
/* iter(loader): burns the DataLoader base seed (2 x u32) from the torch stream. */
int fumi_sampler_new_iterator(fumi_sampler* s, uint32_t* torch_state);
/* One meta-batch.  classes[B,N] split-class per tuple position, label_perm[B,N] label of tuple
 * position p, sup_ids[B,N*K], qry_ids[B,N*Q] image ids, sup_y / qry_y labels,
 * head_class[B,N] split-class carrying label i, sup_rows / qry_rows = position of each image in
 * class_image_ids (row of the split's HBM feature matrix, which is stored in that order).
 * num_threads <= 0: hardware concurrency. */
int fumi_sampler_next(fumi_sampler* s, int64_t B, uint32_t* py_state, uint32_t* torch_state,
                      int64_t* classes, int64_t* label_perm, int64_t* sup_ids, int64_t* qry_ids,
                      int64_t* sup_y, int64_t* qry_y, int64_t* head_class,
                      int64_t* sup_rows, int64_t* qry_rows, int32_t num_threads);
/* Device-resident form of the same meta-batch, in two calls.
 * fumi_sampler_plan (HOST, sequential, cheap): advances the three sequential streams exactly as
 * fumi_sampler_next does -- class tuples from `random`, the split's RandomState(0) support / query
 * shuffles, torch.randperm(N) labels -- and leaves the independent hash-seeded per (task, class)
 * permutations (the bulk of the work: one MT19937 seeding + full Fisher-Yates over the class's
 * images each; torchmeta ClassSplitter_.__getitem__, RandomState((hash(tuple) + c + 0) % 2**32))
 * to the device.  Outputs (HOST): classes / label_perm / head_class as above, perm_seed[B,N] the
 * RandomState seed of tuple position p, picks[B,N,K+Q] the position in that permutation's first
 * K+Q entries taken by support slot k (picks[..,k]) and query slot q (picks[..,K+q]).
 * fumi_sampler_expand (DEVICE pointers, one kernel, one warp per (task, class)): seeds the
 * generator, runs numpy's backward Fisher-Yates over the class in shared memory and writes
 * sup_ids / qry_ids / sup_y / qry_y / sup_rows / qry_rows straight into HBM, so the index arrays the
 * episode kernels gather by never exist on the host.  class_offsets[C+1] / class_image_ids are the
 * device copies of the fumi_sampler_create tables; max_class_size = max_c n_c (sizes the shared
 * memory).  job_order (optional, from the plan) lists the (task, class) jobs longest class first: a job costs
 * its class size, so this keeps a long job from starting last.  Bit-identical to fumi_sampler_next. */
int fumi_sampler_plan(fumi_sampler* s, int64_t B, uint32_t* py_state, uint32_t* torch_state,
                      int64_t* classes, int64_t* label_perm, int64_t* head_class,
                      uint32_t* perm_seed, int32_t* picks, int32_t* job_order /* [B,N] or NULL */);
int fumi_sampler_expand(const int64_t* class_offsets, const int64_t* class_image_ids, int64_t max_class_size,
                        const int64_t* classes, const int64_t* label_perm, const uint32_t* perm_seed,
                        const int32_t* picks, const int32_t* job_order /* or NULL */,
                        int64_t B, int32_t N, int32_t K, int32_t Q,
                        int64_t* sup_ids, int64_t* qry_ids, int64_t* sup_y, int64_t* qry_y,
                        int64_t* sup_rows, int64_t* qry_rows, void* stream);
/* CPython hash(tuple of small non-negative ints) -- exposed for tests. */
int64_t fumi_py_tuple_hash(const int64_t* items, int64_t n);

/* ------------------------------------------------------------------------------------------
 * Diagnostics (no reference counterpart): per-phase SM-cycle counters of the episode kernels, summed over
 * CTAs by thread 0 of each.  fumi_debug_phase_profile(1) zeroes and enables them, (0) disables;
 * fumi_debug_read_phases copies the 64 counters to a HOST array (phase ids: csrc/episode.cu pc.mark).
 * ---------------------------------------------------------------------------------------- */
int fumi_debug_phase_profile(int enable);
int fumi_debug_read_phases(unsigned long long* out64);

#ifdef __cplusplus
}
#endif
#endif /* FUMI_B200_H_ */
