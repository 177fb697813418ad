/* fsar.h — C ABI of libfsar_sm100.so: the B200 (sm_100a) few-shot video inference path of CLIP-FSAR.
 *
 * The reference (alibaba-mmai-research/CLIP-FSAR) is pure Python/PyTorch and has no FFI of its own; the
 * boundary this library plugs into is the registered head module
 *     HEAD_REGISTRY.get(cfg.VIDEO.HEAD.NAME)(cfg=cfg)          models/base/models.py:40
 *     CNN_OTAM_CLIPFSAR.forward(inputs) -> {'logits','class_logits'}   models/base/few_shot.py:2772-2990
 * Each entry point below names the reference code it replaces. The Python binding a maintainer adds is the
 * ctypes stub in clip_fsar_b200/lib.py (shown in INTEGRATION.md).
 *
 * Conventions
 *   - return 0 on success, a negative FSAR_E_* code on failure; fsar_last_error() gives the message.
 *     Nothing throws across the ABI.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host synchronisation
 *     happens inside the *_forward / op calls (the *_host variants synchronise, by definition).
 *   - the caller owns every input/output buffer; the library owns the packed weights and its workspace.
 *   - pointers named *_dev are device pointers, *_host are host pointers. All tensors are contiguous fp32
 *     unless stated otherwise. One handle per device, not thread-safe per handle.
 *   - there is no CPU fallback: every entry point fails with FSAR_E_CUDA when no sm_100 device is present.
 *   - every entry point runs on the handle's device (cfg.device) and restores the caller's current device on return.
 *   - input validation that needs device data (labels) happens on the device: kernels clamp every index they form from
 *     caller labels and raise a flag in mapped host memory. *_collect_host reports it for the batch it collects
 *     (FSAR_E_INVALID, after its event wait); the stream-ordered *_forward calls, which never synchronise, report it at
 *     the NEXT call on the handle.
 */
#ifndef FSAR_H_
#define FSAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSAR_VERSION 100

enum {
    FSAR_OK = 0,
    FSAR_E_INVALID = -1,   /* bad argument / unsupported geometry */
    FSAR_E_CUDA = -2,      /* CUDA runtime / driver error, or no sm_100 device */
    FSAR_E_NOMEM = -3,     /* allocation failure */
    FSAR_E_NAME = -4,      /* unknown weight / buffer name */
    FSAR_E_STATE = -5      /* weights missing, capacity exceeded, ... */
};

typedef struct fsar_handle fsar_handle;

/* Geometry of the frame encoder (VisionTransformer.__init__, few_shot.py:655-669), the temporal
 * modulator (Transformer_v1 as built at few_shot.py:2736-2739) and capacities of the workspace. */
typedef struct fsar_config {
    int32_t image_size;      /* 224 */
    int32_t patch_size;      /* 16 (ViT-B/16), 14 (ViT-L/14) */
    int32_t width;           /* 768 / 1024; multiple of 128 */
    int32_t layers;          /* 12 / 24 */
    int32_t heads;           /* width / 64 */
    int32_t embed_dim;       /* 512 / 768 = mid_dim of the head */
    int32_t mod_heads;       /* 8 */
    int32_t mod_dim_head;    /* embed_dim / 8 */
    int32_t mod_mlp_dim;     /* 2048 (Transformer_v1 default) */
    int32_t mod_depth;       /* TRAIN.TRANSFORMER_DEPTH, default 1 */
    int32_t max_frames;      /* frames encoded per ViT pass; larger requests are chunked */
    int32_t max_videos;      /* capacity: support + query videos of one episode */
    int32_t max_tokens;      /* capacity: DATA.NUM_INPUT_FRAMES (T <= 32) */
    int32_t max_classes;     /* capacity: rows of text_features_{train,test} */
    int32_t max_batch;       /* capacity: episodes per fsar_episodes_* call (1 = single-episode use) */
    float otam_lambda;       /* 0.5, OTAM_cum_dist_v2 default (few_shot.py:2657) */
    int32_t device;          /* CUDA device ordinal */
} fsar_config;

/* One episode, the contents of the `inputs` dict of CNN_OTAM_CLIPFSAR.forward (few_shot.py:2773).
 * Frames of one video are adjacent; videos are in the (shuffled) order of the labels. */
typedef struct fsar_episode {
    const float* support_frames;       /* [n_support * T, 3, S, S] */
    const float* target_frames;        /* [n_target  * T, 3, S, S] */
    const float* support_labels;       /* [n_support] episode-local labels stored as fp32 (ssv2_few_shot.py:278-283) */
    const float* real_support_labels;  /* [n_support] dataset class ids as fp32, index text_features_test */
    int32_t n_support;                 /* way * shot videos */
    int32_t n_target;                  /* query videos */
    int32_t n_frames;                  /* T = DATA.NUM_INPUT_FRAMES */
    int32_t way;                       /* number of distinct support labels (torch.unique, few_shot.py:2965) */
    int32_t merge_before;              /* TRAIN.MERGE_BEFORE  (few_shot.py:2949) */
    int32_t single_direct;             /* TRAIN.SINGLE_DIRECT (few_shot.py:2979) */
    int32_t text_mode;                 /* 0 = visual path (default); 1 = TRAIN.EVAL_TEXT (few_shot.py:2835-2852);
                                          2 = TRAIN.COMBINE (2855-2930). Modes 1/2 return class_logits = None in the
                                          reference: the episode's slice of class_logits is zero-filled. */
    float text_coff;                   /* TRAIN.TEXT_COFF, exponent of the text probability in mode 2 (default 0.9) */
} fsar_episode;

/* Per-kernel-class device time of the calls issued between fsar_profile_begin/end (CUDA events on `stream`). */
#define FSAR_PROF_CLASSES 13
typedef struct fsar_profile {
    double ms[FSAR_PROF_CLASSES];        /* summed device milliseconds */
    int64_t launches[FSAR_PROF_CLASSES]; /* kernel launches */
    double flops[FSAR_PROF_CLASSES];     /* algorithmic FLOPs (GEMM / attention classes) */
    double bytes[FSAR_PROF_CLASSES];     /* algorithmic bytes (HBM-bound classes) */
} fsar_profile;
enum {
    FSAR_K_PATCH_GATHER = 0, FSAR_K_GEMM_PATCH = 1, FSAR_K_LAYERNORM = 2, FSAR_K_GEMM_QKV = 3,
    FSAR_K_ATTENTION = 4, FSAR_K_GEMM_OUT = 5, FSAR_K_GEMM_FC1 = 6, FSAR_K_GEMM_FC2 = 7,
    FSAR_K_FINAL_PROJ = 8, FSAR_K_HEAD_MISC = 9, FSAR_K_MODULATOR = 10, FSAR_K_COS_OTAM = 11,
    FSAR_K_LAST_BLOCK_CLS = 12   /* last block, CLS rows only: Q / out_proj / fc1 / fc2 GEMMs over n_frames rows, CLS attention, ln_2 */
};

int fsar_version(void);
const char* fsar_class_name(int kernel_class);

/* Lifetime. Replaces CNN_OTAM_CLIPFSAR.__init__ (few_shot.py:2695-2739) minus the CLIP checkpoint/text tower. */
int fsar_create(const fsar_config* cfg, fsar_handle** out);
void fsar_destroy(fsar_handle* h);
/* Message of the last failure on this handle (or of the last failed fsar_create when h == NULL). */
const char* fsar_last_error(const fsar_handle* h);

/* Weights. `name` is the reference head's state_dict key without the "head." prefix
 * (utils/checkpoint.py:329 load_state_dict(strict=False) looks these up), e.g.
 *   "backbone.conv1.weight", "backbone.transformer.resblocks.3.attn.in_proj_weight",
 *   "context2.layers.0.0.fn.to_q.weight", "scale",
 * plus the two non-state attributes "text_features_train" / "text_features_test" (few_shot.py:2720,2728;
 * numel / embed_dim rows). `data` holds `numel` fp32 values on the host (on_device == 0) or device.
 * GEMM weights are re-packed to the 16-bit tensor-core operand type inside the library.
 * Ordering: fsar_set_weight is a synchronising call. It waits for all work in flight on the device (no forward may still
 * be reading the weight), copies + re-packs, and waits again, so a forward enqueued afterwards on ANY stream sees the
 * new value. `data` may be released as soon as the call returns. */
int fsar_set_weight(fsar_handle* h, const char* name, const float* data, int64_t numel, int on_device);
/* Number of weights that have not been set yet (0 => ready); names via fsar_missing_weight(i). */
int fsar_missing_weights(const fsar_handle* h);
const char* fsar_missing_weight(const fsar_handle* h, int i);

/* VisionTransformer.forward (few_shot.py:671-688): frames [n,3,S,S] -> features [n, embed_dim]. */
int fsar_vit_forward(fsar_handle* h, const float* frames_dev, int n_frames, float* feats_dev, void* stream);

/* Transformer_v1.forward(x, x, x) (few_shot.py:990-999): x [n_seq, n_tok, embed_dim] -> same shape. */
int fsar_modulate(fsar_handle* h, const float* x_dev, int n_seq, int n_tok, float* out_dev, void* stream);

/* cos_sim + OTAM_cum_dist_v2 (+ transpose direction) + negation (few_shot.py:2970-2989):
 * q [Q,T,E], protos [way,T,E] -> logits [Q,way]. dists_dev [Q,way,T,T] and cum_dev [Q,way] may be NULL. */
int fsar_otam_logits(fsar_handle* h, const float* q_dev, const float* protos_dev, int Q, int way, int T,
                     int single_direct, float* logits_dev, float* dists_dev, float* cum_dev, void* stream);

/* CNN_OTAM_CLIPFSAR.forward, eval branch (few_shot.py:2932-2990), device-resident inputs.
 * logits_dev [n_target, way]; class_logits_dev [n_support + n_target, n_train_classes] (may be NULL). */
int fsar_episode_forward(fsar_handle* h, const fsar_episode* ep_dev, float* logits_dev, float* class_logits_dev,
                         void* stream);

/* Throughput form of the same forward: n_episodes independent episodes (a host array of fsar_episode holding DEVICE
 * pointers) in one call. All their frames are encoded in ViT passes of cfg.max_frames frames irrespective of episode
 * boundaries — with max_frames = 96 every ViT-B/16 GEMM runs whole waves of 256 x 256 tiles on 148 SMs — and the head
 * runs per episode. Results are concatenated: logits_dev [sum_i n_target_i * way_i], class_logits_dev (may be NULL)
 * [sum_i (n_support_i + n_target_i) * n_train_classes]. Same math per episode as fsar_episode_forward. */
int fsar_episodes_forward(fsar_handle* h, const fsar_episode* eps_dev, int n_episodes, float* logits_dev,
                          float* class_logits_dev, void* stream);
/* Host-buffer form of fsar_episodes_forward, pipelined over two slots (0 / 1): submit enqueues the host->device copies of
 * the frames and labels on the library's copy stream, the compute and the device->host copy of the results, and returns
 * WITHOUT waiting. Lifetime rule: every buffer an fsar_episode of the batch points to must stay allocated and unmodified
 * until fsar_episodes_collect_host(slot) has returned (with pinned memory the copies are still in flight when submit
 * returns). collect waits for the slot, reports label errors of that batch, and copies out the results. */
int fsar_episodes_submit_host(fsar_handle* h, int slot, const fsar_episode* eps_host, int n_episodes);
int fsar_episodes_collect_host(fsar_handle* h, int slot, float* logits_host, float* class_logits_host);

/* Same, with HOST buffers: host->device copies of the frames/labels and the device->host copy of the
 * results happen inside the call (pinned host memory makes them asynchronous up to the final sync).
 * This is what the reference runner does at runs/test_net_few_shot.py:61-62 + .item() at 174-178. */
int fsar_episode_forward_host(fsar_handle* h, const fsar_episode* ep_host, float* logits_host,
                              float* class_logits_host);
/* Pipelined variant: submit copies + compute of an episode into slot (0/1) without waiting, collect later.
 * Lets the copy of episode i+1 overlap the compute of episode i. */
int fsar_episode_submit_host(fsar_handle* h, int slot, const fsar_episode* ep_host);
int fsar_episode_collect_host(fsar_handle* h, int slot, float* logits_host, float* class_logits_host);

/* Frame pre-processing of the reference's test-time loader, fused on the device:
 * ToTensorVideo -> KineticsResizedCropFewshot(bilinear resize to resize_h x resize_w, centre crop image_size) ->
 * NormalizeVideo (datasets/base/ssv2_few_shot.py:633-642, datasets/utils/transformations.py:676-716).
 * frames_u8_dev uint8 [n_frames, H, W, 3] -> out_dev fp32 [n_frames, 3, image_size, image_size] (task-dict layout). */
int fsar_preprocess_u8(fsar_handle* h, const uint8_t* frames_u8_dev, int n_frames, int H, int W, int resize_h, int resize_w,
                       const float mean[3], const float std[3], float* out_dev, void* stream);
/* VisionTransformer.forward on RAW frames: the transform above is evaluated inside the patch gather (uint8 [H, W, 3] ->
 * 16-bit im2col rows of conv1), the fp32 [n, 3, S, S] crop is never materialised. Bit-identical to
 * fsar_preprocess_u8 followed by fsar_vit_forward. */
int fsar_vit_forward_u8(fsar_handle* h, const uint8_t* frames_u8_dev, int n_frames, int H, int W, int resize_h, int resize_w,
                        const float mean[3], const float std[3], float* feats_dev, void* stream);
/* fsar_episodes_submit_host with RAW uint8 frames: support_frames / target_frames of each episode point to HOST uint8
 * [videos * n_frames, H, W, 3]; the bytes are copied as they are and pre-processed on the device inside the patch gather
 * (4x less H2D traffic than fp32 crops at 224 x 224 sources, no fp32 intermediate in HBM). Collect with
 * fsar_episodes_collect_host; the lifetime rule of fsar_episodes_submit_host applies. */
int fsar_episodes_submit_host_u8(fsar_handle* h, int slot, const fsar_episode* eps_host_u8, int n_episodes, int H, int W,
                                 int resize_h, int resize_w, const float mean[3], const float std[3]);

/* CLIP text tower: CLIP.encode_text (few_shot.py:793-806) = token + positional embedding, `layers`
 * ResidualAttentionBlocks (619-640) under the causal mask of build_attention_mask (777-783), ln_final on the
 * end-of-text position (text.argmax(-1)), @ text_projection. It is what CNN_OTAM_CLIPFSAR.__init__ runs once to produce
 * text_features_{train,test} (few_shot.py:2714-2728) from tokenize("a photo of {class}") (393-429; the BPE tokenizer
 * stays on the host, this entry point takes token ids).
 * fsar_text_configure registers the tower's weights under their CLIP state_dict names with a "clip." prefix:
 *   clip.token_embedding.weight [vocab, width], clip.positional_embedding [context, width],
 *   clip.transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.{weight,bias},
 *   ln_1.{weight,bias}, ln_2.{weight,bias}, mlp.c_fc.{weight,bias}, mlp.c_proj.{weight,bias}},
 *   clip.ln_final.{weight,bias}, clip.text_projection [width, embed_dim]
 * (set them with fsar_set_weight; they are not needed by the episode entry points). The blocks run on the frame
 * encoder's kernels and workspace: width must be a multiple of 128 with heads * 64 == width, context_length <= 208. */
typedef struct fsar_text_config {
    int32_t width;           /* 512 (ViT-B/16, ViT-B/32), 768 (ViT-L/14) */
    int32_t layers;          /* 12 */
    int32_t heads;           /* width / 64 */
    int32_t context_length;  /* 77 */
    int32_t vocab_size;      /* 49408 */
} fsar_text_config;
int fsar_text_configure(fsar_handle* h, const fsar_text_config* cfg);
/* tokens_dev int32 [n_texts, context_length] (tokenize(), few_shot.py:393-429) -> out_dev fp32 [n_texts, embed_dim]. */
int fsar_text_encode(fsar_handle* h, const int32_t* tokens_dev, int n_texts, float* out_dev, void* stream);

/* Caller-side metrics (runs/test_net_few_shot.py:111 cross-entropy, 147 topks_correct, 151-160 per-class accuracy) kept
 * on the device: counters_dev int64[3] += {n_correct_top1, n_queries, round(sum CE * 1e6)}; per_class_dev (may be NULL)
 * int64[2 * way] += {hits per class | queries per class}. No host synchronisation; read the counters once per run. */
int fsar_metrics_update(fsar_handle* h, const float* logits_dev, const float* target_labels_dev, int Q, int way,
                        int64_t* counters_dev, int64_t* per_class_dev, void* stream);

/* Intermediates of the last episode (test taps): "support_feats" [S,T,E], "target_feats" [Q,T,E],
 * "mod_out" [rows,E], "protos" [way,T,E], "dists" [Q,way,T,T], "cum_dists" [Q,way], "class_index" (int32 [S]).
 * Copies min(numel, available) elements to host after synchronising `stream`; returns the element count. */
int64_t fsar_peek(fsar_handle* h, const char* name, void* dst_host, int64_t numel, void* stream);

/* Single operators (per-kernel parity tests call these through the ABI). 16-bit buffers use the library's
 * operand type (fsar_operand_dtype(): 0 = fp16, 1 = bf16). */
int fsar_operand_dtype(void);
int fsar_op_layernorm(fsar_handle* h, const float* x_dev, const float* gamma_dev, const float* beta_dev, int rows,
                      int dim, int out16, void* out_dev, void* stream);
/* C[M,N] = A16[M,K] * W16[N,K]^T with epilogue epi (0 store16, 1 quickgelu16, 2 resid32 (out += ...), 4 store32). */
int fsar_op_gemm(fsar_handle* h, const void* a16_dev, const void* w16_dev, const float* bias_dev, int M, int N, int K,
                 int epi, void* out_dev, void* stream);
/* qkv16 [n_frames * L, 3 * D] -> out16 [n_frames * L, D], D = heads * 64, L <= 257 (tcgen05 / TMEM attention core) */
int fsar_op_attention(fsar_handle* h, const void* qkv16_dev, int n_frames, int L, int heads, void* out16_dev,
                      void* stream);
int fsar_op_f32_to_16(fsar_handle* h, const float* src_dev, void* dst16_dev, int64_t numel, void* stream);

/* Instrumentation. */
int64_t fsar_launch_count(const fsar_handle* h);            /* kernels launched by this handle so far */
int fsar_profile_begin(fsar_handle* h);                     /* start per-kernel event timing */
int fsar_profile_end(fsar_handle* h, fsar_profile* out);    /* synchronises, fills out, stops timing */

#ifdef __cplusplus
}
#endif
#endif /* FSAR_H_ */
