/*
 * mfas_b200.h -- C ABI of the B200-native MFAS candidate-training hot path.
 *
 * The reference (jperezrua/mfas) is pure Python/PyTorch and has no FFI; the seam this library
 * sits behind is the function-valued plug-in entry of the search driver,
 *     dataset_searchmethods['train_sampled_fun']      models/searchable.py:57,90,120
 * i.e. models/search/ntu_searchable.py::train_sampled_models (:23-102), plus the model class
 * Searchable_Skeleton_Image_Net (:178-301) and models/search/train_searchable/ntu.py
 * ::train_ntu_track_acc / test_ntu_track_acc (:14-125).  Each entry point below cites the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* or documented "device" is a DEVICE pointer
 *    owned by the caller (a torch tensor); the library never frees or retains it beyond the
 *    lifetime of the handle it was bound to;
 *  - every launch function takes a cudaStream_t (passed as void*) and is asynchronous on it;
 *  - return value: 0 = MFAS_OK, negative = error; mfas_last_error() gives the message for the
 *    calling thread; no C++ exception crosses the ABI;
 *  - one host thread per handle at a time.
 */
#ifndef MFAS_B200_H_
#define MFAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFAS_ABI_VERSION 5

#define MFAS_MAX_LAYERS 8     /* fusion steps per candidate (reference max_fusions default 4)   */
#define MFAS_MAX_BATCH 128    /* rows per batch (BASELINE configs use 8, 64, 128)               */
#define MFAS_MAX_HIDDEN 256   /* inner_representation_size                                       */
#define MFAS_MAX_CLASSES 64   /* num_outputs (60 NTU)                                            */
#define MFAS_NUM_TAPS 8       /* tap slots per modality: NTU 4 + 4 (ntu_searchable.py:291-292), MM-IMDB 2 + 4, AV-MNIST 5 + 3 */

enum {
  MFAS_OK = 0,
  MFAS_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  MFAS_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  MFAS_ERR_UNSUPPORTED = -3,  /* recognised but not built (e.g. a layer recipe the reference lacks) */
  MFAS_ERR_NOMEM = -4,
  MFAS_ERR_UNBOUND = -5       /* a candidate of the group has no arenas bound */
};

enum {
  MFAS_FLAG_BN = 1,        /* args.batchnorm : Linear -> act -> BatchNorm1d     (ntu_searchable.py:274-282) */
  MFAS_FLAG_DROPOUT = 2,   /* args.drpt>1e-10: ... -> Dropout(p)                (ntu_searchable.py:274-279) */
  MFAS_FLAG_ALPHAS = 4,    /* args.alphas    : AlphaScalarMultiplication gate   (aux_models.py:94-111)      */
  MFAS_FLAG_MULTITASK = 8, /* args.multitask : + cached backbone logits         (train_searchable/ntu.py:59-61) */
  MFAS_FLAG_MULTILABEL = 16, /* MM-IMDB head: WeightedCrossEntropyWithLogits (aux_models.py:129-147) on multi-hot
                              * targets instead of softmax-CE; the per-batch statistic is the sum of per-sample F1 of
                              * sigmoid(logits) > 0.3 (train_searchable/mmimdb.py:84,101) instead of #correct.  All candidates
                              * of a group share this flag. */
  MFAS_FLAG_PLAIN = 32      /* the AV-MNIST recipe (avmnist_searchable.py:276-285): Linear -> act [-> Dropout], never a BatchNorm;
                              * with it mfas_plan_layout accepts "no BatchNorm, no Dropout" (which the NTU network rejects) */
};

enum { MFAS_ACT_RELU = 0, MFAS_ACT_SIGMOID = 1, MFAS_ACT_LRELU = 2 };  /* conf[:,2], ntu_searchable.py:267-272 */

/* One split of the feature cache (replaces the backbone forward, ntu_searchable.py:211-225).
 * Tap t of a modality is a row-major [n_rows, d] fp32 matrix with leading dimension ld (floats). */
typedef struct mfas_cache_desc {
  int64_t n_rows;
  const float* ske[MFAS_NUM_TAPS];
  const float* rgb[MFAS_NUM_TAPS];
  int64_t ske_ld[MFAS_NUM_TAPS];
  int64_t rgb_ld[MFAS_NUM_TAPS];
  int32_t d_ske[MFAS_NUM_TAPS];
  int32_t d_rgb[MFAS_NUM_TAPS];
  const int64_t* labels;     /* [n_rows] class ids (single-label heads) */
  const float* logit_rgb;    /* [n_rows, C] backbone logits (multitask) or NULL */
  const float* logit_ske;
  const float* targets;      /* [n_rows, C] multi-hot fp32 targets (MFAS_FLAG_MULTILABEL; labels may then be NULL) or NULL */
  const float* pos_weight;   /* [C] positive-class weights of the weighted BCE (MFAS_FLAG_MULTILABEL) or NULL */
} mfas_cache_desc;

/* Where each tensor of one candidate lives inside its flat fp32 arenas.
 * Parameter arena (same offsets index the Adam m / v arenas and the optional grad arena):
 *   fusion_layers.l.0.weight [H,K_l] | .0.bias [H] | .2.weight [H] | .2.bias [H]   (l = 0..L-1)
 *   central_classifier.weight [C,H] | .bias [C] | alphas.l.alpha_x [1] (padded to 4)
 * Buffer arena: fusion_layers.l.2.running_mean [H] | running_var [H].
 * Offsets are in floats and 16-byte aligned.  State-dict names: SURVEY.md section 4. */
typedef struct mfas_layout {
  int32_t L, H, C, flags;
  int32_t conf[MFAS_MAX_LAYERS][3];   /* [ske tap, rgb tap, activation] per fusion step */
  int32_t d_ske[MFAS_MAX_LAYERS];     /* width of the selected taps */
  int32_t d_rgb[MFAS_MAX_LAYERS];
  int32_t K[MFAS_MAX_LAYERS];         /* Linear in_features = d_ske + d_rgb + (l>0)*H, ntu_searchable.py:261-264 */
  int64_t off_W[MFAS_MAX_LAYERS];
  int64_t off_b[MFAS_MAX_LAYERS];
  int64_t off_gamma[MFAS_MAX_LAYERS]; /* -1 when no BN */
  int64_t off_beta[MFAS_MAX_LAYERS];
  int64_t off_alpha[MFAS_MAX_LAYERS];
  int64_t off_Wc, off_bc;
  int64_t n_params;                   /* floats in the parameter arena */
  int64_t off_rm[MFAS_MAX_LAYERS];    /* -1 when no BN */
  int64_t off_rv[MFAS_MAX_LAYERS];
  int64_t n_bufs;                     /* floats in the buffer arena (>= 4) */
} mfas_layout;

/* Arenas of one candidate; all device pointers, caller-owned (views of torch tensors, so that
 * state_dict() stays PyTorch-native). grad may be NULL (gradients then never touch HBM). */
typedef struct mfas_arenas {
  float* params;      /* [n_params] */
  float* adam_m;      /* [n_params] exp_avg      */
  float* adam_v;      /* [n_params] exp_avg_sq   */
  float* grad;        /* [n_params] or NULL: raw dL/dp of the last train step (parity tests)   */
  float* bufs;        /* [n_bufs]   BN running stats */
  int64_t* nbt;       /* [L] num_batches_tracked      */
} mfas_arenas;

typedef struct mfas_adam_hparams {   /* torch.optim.Adam(params, lr, weight_decay=1e-4), ntu_searchable.py:65 */
  double beta1, beta2, eps, weight_decay;   /* python floats: torch forms 1-beta in fp64 before rounding to fp32 */
} mfas_adam_hparams;

typedef struct mfas_run_args {       /* one train_ntu_track_acc run for every candidate of a group */
  int32_t n_epochs;
  int32_t batch;                     /* args.batchsize; last batch of a pass may be short (drop_last=False) */
  const int32_t* perm_train;         /* device [n_cand][n_epochs][n_train] row order of each train pass */
  const int32_t* perm_dev;           /* device [n_cand][n_epochs][n_dev]  or NULL = identity */
  const float* step_size;            /* HOST [n_epochs*ceil(n_train/batch)]: lr_t / (1-beta1^t), formed in fp64 */
  const float* bc2_sqrt;             /* HOST, same length: sqrt(1-beta2^t) */
  int64_t adam_t0;                   /* optimiser steps taken before this run */
  double* stats;                     /* device [n_cand][n_epochs][4]: train loss sum, train correct, dev loss sum, dev correct
                                      * (MFAS_FLAG_MULTILABEL: "correct" is the sum of per-sample F1 scores) */
  double* best_acc;                  /* device [n_cand] best dev accuracy (strict >, starts at 0); MULTILABEL: best dev F1-samples */
  int32_t* best_epoch;               /* device [n_cand] epoch of best_acc or -1 */
  const double* best_acc_init;       /* HOST [n_cand] or NULL (= 0): the value best_acc starts from -- init_f1 of train_mmimdb_track_f1
                                      * (train_searchable/mmimdb.py:15,21): an epoch is snapshotted only when it beats it (strict >; a NaN
                                      * is never beaten, so the run ends on the incoming weights) */
} mfas_run_args;

typedef struct mfas_group* mfas_group_t;

/* ---- host-only helpers (no GPU needed) ---------------------------------------------------- */
int mfas_abi_version(void);
const char* mfas_last_error(void);

/* Layout of one candidate. conf: [L][3] ints. Mirrors _create_fc_layers / _create_alphas
 * (ntu_searchable.py:258-296). Fails with MFAS_ERR_UNSUPPORTED for drpt<1e-10 && !batchnorm, the
 * combination for which the reference has no layer recipe (UnboundLocalError). */
int mfas_plan_layout(int32_t L, const int32_t* conf, int32_t H, int32_t C, int32_t flags,
                     const int32_t d_ske[MFAS_NUM_TAPS], const int32_t d_rgb[MFAS_NUM_TAPS],
                     mfas_layout* out);

/* Algorithmic bytes / flops of one step; out[0..3] = train bytes,
 * eval bytes, fwd flops, bwd flops. */
int mfas_algorithmic_counts(const mfas_layout* lay, int32_t batch, double out[4]);

/* ---- group of candidates trained together on one device ------------------------------------
 * Replaces the per-candidate construct/.to(device) of train_sampled_models (ntu_searchable.py:38-72). */
int mfas_group_create(int32_t device, int32_t n_cand, const mfas_layout* layouts, int32_t batch_max,
                      float dropout_p, uint32_t dropout_seed, const int32_t* cand_ids /* [n_cand] or NULL */,
                      mfas_group_t* out);
int mfas_group_destroy(mfas_group_t g);
/* The workspace and partial-sum blocks of a destroyed group are parked for the next mfas_group_create on the same
 * device (the reference builds and drops a model per candidate, ntu_searchable.py:38-99; here the unit is a group per
 * train_sampled_models call).  This returns everything parked to the CUDA driver.  MFAS_POOL_MB caps it (default 4096). */
int mfas_release_cached_memory(void);
int mfas_group_bind(mfas_group_t g, int32_t cand, const mfas_arenas* arenas);
int mfas_group_set_adam(mfas_group_t g, const mfas_adam_hparams* hp);
/* Initial weights of every candidate of the group in ONE launch, the distributions of the reference constructor
 * (ntu_searchable.py:200, :267-282 via nn.Linear.reset_parameters / nn.BatchNorm1d, :202-204): Linear weight and bias
 * U(-1/sqrt(fan_in), 1/sqrt(fan_in)), BatchNorm weight 1 / bias 0 / running_mean 0 / running_var 1 / num_batches_tracked 0,
 * alpha N(0, 0.1) -- drawn from a counter-based generator keyed by (seed, candidate id, tensor, element), so a candidate
 * gets the same weights whichever group / rank / device trains it.  NOT the CPU stream of the reference constructor (for
 * seed parity with it the caller fills the arenas itself, mfas_host_uniform_fill).  Adam moments are zeroed. */
int mfas_group_init_params(mfas_group_t g, uint64_t seed, void* stream);
int mfas_group_num_launches(mfas_group_t g, int64_t* out);   /* kernels launched through g so far */
/* Which kernels serve this group: 1 = "tc" (tcgen05 tensor cores, 3xTF32; inner_representation_size a
 * multiple of 64), 0 = "ffma" (fp32 CUDA cores; any multiple of 16). Both are sm_100a CUDA in this
 * library; MFAS_ENGINE=ffma|tc in the environment forces one (bisecting / tests). */
int mfas_group_engine(mfas_group_t g, int32_t* out);
/* Synchronises the device and returns MFAS_ERR_CUDA if a kernel reported a failure (a bounded
 * tensor-core barrier wait that expired) since the group was created. */
int mfas_group_status(mfas_group_t g);

/* Forward of every candidate over one batch: Searchable_Skeleton_Image_Net.forward
 * (ntu_searchable.py:206-247) given cached taps. rows: device int32 row ids, candidate c reads
 * rows + c*rows_stride (stride 0 = shared batch). train!=0: BN batch statistics + running-stat
 * update (+dropout, keyed by step). d_logits: [n_cand][batch_max][C] or NULL.
 * d_loss / d_correct: [n_cand] mean CE loss and #correct of the batch, or NULL
 * (MFAS_FLAG_MULTILABEL: mean weighted BCE and the number of rows whose thresholded label set is exactly right). */
int mfas_forward(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_rows, int64_t rows_stride,
                 int32_t n_rows, int32_t train, int64_t step, float* d_logits, float* d_loss,
                 int32_t* d_correct, void* stream);

/* One optimiser step of every candidate: forward + CE + hand-derived backward + Adam(L2)
 * (train_searchable/ntu.py:46-69). step_size = lr/(1-beta1^t), bc2_sqrt = sqrt(1-beta2^t). */
int mfas_train_step(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_rows, int64_t rows_stride,
                    int32_t n_rows, float step_size, float bc2_sqrt, int64_t step, float* d_logits,
                    float* d_loss, int32_t* d_correct, void* stream);

/* The production path: num_epochs x (train pass + dev pass) for every candidate, best-dev
 * snapshot and final rollback, nothing returns to the host in between
 * (train_searchable/ntu.py:14-89).  The dev pass is an eval-mode pass (see mfas_eval_pass for its step width). */
int mfas_train_run(mfas_group_t g, const mfas_cache_desc* train, const mfas_cache_desc* dev,
                   const mfas_run_args* args, void* stream);

/* Accuracy pass in eval mode (test_ntu_track_acc, train_searchable/ntu.py:92-125).
 * d_perm: [n_cand][n_rows] or NULL = identity. d_out: [n_cand][2] loss sum, #correct.
 * Rows are independent in eval mode (running BatchNorm statistics, no dropout), so `batch` only bounds the step width:
 * on the tensor-core engine with batch_max <= 64 the pass runs 128 rows per step (the weights are streamed half as
 * often); #correct is unchanged, the loss sum differs by fp32 summation order.  MFAS_EVAL128=0 keeps `batch`. */
int mfas_eval_pass(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_perm, int32_t batch,
                   double* d_out, void* stream);

/* Measurement aid (bench.py): with profiling on, every train step of the tensor-core engine records CUDA events on the
 * launching stream around its three kernels; mfas_group_last_step_ms returns the device time in milliseconds of
 * {forward streaming kernel, fused chain kernel, backward streaming kernel} of the most recent train step. */
int mfas_group_set_profiling(mfas_group_t g, int32_t on);
int mfas_group_last_step_ms(mfas_group_t g, float* ms3);
/* Debug aid: groups created with MFAS_CHAIN_TIMELINE=1 in the environment record clock64() stamps at the phase
 * boundaries of the fused chain kernel (start, after each forward layer, head, after each backward layer);
 * copies out[n_cand][16] (synchronises the device). */
int mfas_group_chain_timeline(mfas_group_t g, int64_t* out, int32_t n_cand);

/* Host-side helper (no GPU): fills n_ops float arrays with U(from, to) draws taken from torch's global CPU generator
 * stream, bit-identical to torch.nn.init.uniform_/kaiming_uniform_ applied in the same order -- what the reference
 * constructor does for every candidate (models/search/ntu_searchable.py:200, :274-282 via nn.Linear.reset_parameters).
 * torch_rng_state is the byte buffer of torch.get_rng_state() (legacy 5056-byte mt19937 layout), advanced in place.
 * count[op] < 0 asks for -count[op] raw tempered 32-bit words instead (dst is then a uint32_t array): the caller turns
 * them into draws of other distributions (the Box-Muller normals of the scalar alphas, ntu_searchable.py:202-204).
 * Fills of 4 M words and more are pipelined over threads (MFAS_HOST_INIT_THREADS consumers, default 2; 0 = serial).
 * use_fma selects x*(to-from)+from as one fused multiply-add (what an FMA-contracting ATen build computes) or as two
 * roundings.  Returns MFAS_ERR_UNSUPPORTED if the buffer does not look like a seeded mt19937 state. */
int mfas_host_uniform_fill(uint8_t* torch_rng_state, int64_t state_bytes, int32_t n_ops, float* const* dst,
                           const int64_t* count, const float* from, const float* to, int32_t use_fma);

/* Host-only planning aid (no GPU): the tile list the persistent weight-gradient + Adam kernel of the tensor-core engine walks
 * for these candidates -- out[i] = {candidate, layer, first weight column, first output row} of tile i, 128 columns x 64 rows at
 * most, never across two concat sources [first-modality tap | second-modality tap | hidden]; with head_tiles the classifier
 * tiles (layer == L) follow.  out may be NULL to query the counts.  Replaces nothing in the reference (autograd has no tiles);
 * exported so that the tiling rules are testable without a GPU. */
int mfas_plan_bwd_tiles(const mfas_layout* layouts, int32_t n_cand, int32_t head_tiles, int32_t* out, int64_t max_tiles,
                        int64_t* n_tiles, int64_t* n_layer_tiles);

/* ---- feature-cache builder (the step before the path) ----------------------------------
 * Global average pooling of one backbone tap into its column slice of a cache matrix:
 *   d_out[b * out_ld + c] = mean over s of d_in[(b * C + c) * S + s]
 * GlobalPooling2D.forward (models/auxiliary/aux_models.py:58-64), which the reference applies to every selected tap on
 * every batch (models/search/ntu_searchable.py:224-225); here it runs once per sample when the cache is built.
 * d_in: device fp32 [B, C, S] contiguous (S = product of the trailing dims, 1 for a vector tap); asynchronous on stream. */
int mfas_global_pool(int32_t device, const float* d_in, int64_t B, int64_t C, int64_t S, float* d_out, int64_t out_ld,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFAS_B200_H_ */
