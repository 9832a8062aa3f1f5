/* libbv2 -- C ABI of the Blackwell-native VITS2 inference engine (Bert-VITS2 v2.3 `SynthesizerTrn.infer`).
 *
 * The reference has no FFI/plugin layer: its seam is the Python class `models.SynthesizerTrn`
 * (reference models.py:811-1074) constructed by `infer.get_net_g` (reference infer.py:84-104) and driven by
 * `infer.infer` / `infer_multilang` (reference infer.py:302-318, 407-423).  This header is what a ctypes
 * binding behind that class binds (bert_vits2_b200/engine.py; INTEGRATION.md shows the reference-side stub).
 *
 * Conventions: every function returns 0 on success or a negative bv2_status; nothing throws or aborts across
 * the ABI; `bv2_last_error` returns a thread-unsafe, engine-owned message for the last failure.  The caller
 * owns every input/output buffer (device pointers unless noted, fp32 contiguous, reference tensor layouts
 * [B,C,T]); the engine owns weights and workspace.  Work is enqueued on the caller's `stream`
 * (a cudaStream_t passed as void*); the only host synchronisation is inside bv2_infer_begin (one read-back of
 * y_lengths, the same data-dependent length the reference syncs on at models.py:1058 / commons.py:120-121).
 * One engine per device; concurrent callers are serialised by an internal mutex (ctypes drops the GIL).
 * There is NO CPU fallback: creation fails if no sm_100 device is present.
 */
#ifndef BV2_H_
#define BV2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bv2_engine bv2_engine;

typedef enum {
    BV2_OK = 0,
    BV2_ERR_ARG = -1,      /* bad argument / unsupported configuration (reference raises ValueError) */
    BV2_ERR_STATE = -2,    /* call order violated (weights missing, not finalized, begin/finish mismatch) */
    BV2_ERR_CUDA = -3,     /* CUDA runtime error */
    BV2_ERR_INTERNAL = -4
} bv2_status;

#define BV2_MAX_UPS 8
#define BV2_MAX_RESBLOCK_KERNELS 4
#define BV2_MAX_DILATIONS 4

/* `hps.model` + ctor args of reference models.SynthesizerTrn.__init__ (models.py:816-841). */
typedef struct {
    int32_t n_vocab, num_tones, num_languages, bert_dim;
    int32_t inter_channels, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, window_size;
    int32_t gin_channels, n_speakers;
    int32_t n_flow_layer, n_layers_trans_flow, use_transformer_flow, flow_kernel_size, wn_layers;
    int32_t upsample_initial_channel, n_ups;
    int32_t upsample_rates[BV2_MAX_UPS], upsample_kernel_sizes[BV2_MAX_UPS];
    int32_t n_resblock_kernels, n_dilations;
    int32_t resblock_kernel_sizes[BV2_MAX_RESBLOCK_KERNELS];
    int32_t resblock_dilation_sizes[BV2_MAX_RESBLOCK_KERNELS][BV2_MAX_DILATIONS];
    int32_t sdp_filter, sdp_kernel, sdp_n_flows, sdp_dds_layers, sdp_num_bins;
    float sdp_tail_bound;
    int32_t dp_filter, dp_kernel, cond_layer_idx;
    int32_t generator_precision; /* 0 = fp32 SIMT convs; 1 = TF32 tcgen05 implicit-GEMM convs (flow + Generator); 2 = FP16-operand
                                    tcgen05 convs (same 11-bit significand as TF32, fp32 accumulate, fp32 activations in HBM) */
    int32_t n_flows;             /* couplings in `flow`: n_flow_layer for TransformerCouplingBlock (models.py:82-145), always 4 for
                                    ResidualCouplingBlock, whose n_layers argument receives n_flow_layer (models.py:403-445, 918-919) */
} bv2_config;

/* replaces: models.SynthesizerTrn(...).to(device) (reference infer.py:95-101) */
int bv2_create(bv2_engine** out, const bv2_config* cfg, int cuda_device);

/* replaces: load_state_dict for one entry of `G_*.pth["model"]` (reference utils.py:65-120).
 * `key` is the reference state_dict key (weight-norm as weight_g/weight_v); host_ptr is HOST memory;
 * dtype: 0 = fp32, 1 = fp16 (compress_model.py:49-52 checkpoints).  Unknown keys (enc_q.*) are ignored. */
int bv2_set_weight(bv2_engine* e, const char* key, const void* host_ptr, const int64_t* shape, int ndim, int dtype);

/* folds weight-norm (incl. ConvTranspose dim-0 = in_channels), folds Flip into coupling weights,
 * packs and uploads.  replaces: net_g.eval() + the per-call weight-norm re-evaluation of the reference. */
int bv2_finalize(bv2_engine* e);

/* Packed engine weight file (the reference's only checkpoint tool is compress_model.py:44-53, which drops enc_q and
 * casts to fp16).  bv2_save_packed dumps the finalized engine's weight arena: weight-norm and Flip already folded, SIMT and
 * tcgen05 operand images already packed for this configuration + precision.  bv2_load_packed replaces the
 * bv2_set_weight... + bv2_finalize sequence on a fresh engine created with the SAME bv2_config: it rebuilds the (cheap) layout
 * bookkeeping and fills device memory with ONE cudaMemcpy of the file image.  Mismatching configurations are rejected. */
int bv2_save_packed(bv2_engine* e, const char* path);
int bv2_load_packed(bv2_engine* e, const char* path);

/* ---- whole path: SynthesizerTrn.infer (reference models.py:1026-1074), split at the data-dependent length.
 * begin: emb_g -> enc_p -> sdp/dp -> durations.  Inputs as the reference passes them (int64 ids, fp32 feats).
 *   noise_w [B,2,T] replaces torch.randn at models.py:249.  w_ceil_override (optional, [B,T] fp32) teacher-forces
 *   durations for parity harnesses.  Writes y_lengths_host[B] (HOST) and *f_max = max(y_lengths).            */
int bv2_infer_begin(bv2_engine* e, int B, int T, const int64_t* x, const int64_t* x_lengths, const int64_t* sid,
                    const int64_t* tone, const int64_t* language, const float* bert, const float* ja_bert,
                    const float* en_bert, const float* noise_w, float noise_scale_w, float length_scale,
                    float sdp_ratio, const float* w_ceil_override, void* stream, int64_t* y_lengths_host,
                    int32_t* f_max);

/* finish: length regulation -> prior sample -> flow reverse -> Generator.  noise_z [B,inter,>=F] with row stride
 * noise_ld replaces torch.randn_like at models.py:1071.  Outputs (caller-allocated, any may be NULL except o):
 *   o [B,1,Fg*hop] with Fg = min(F, max_len), attn [B,1,F,T], y_mask [B,1,F], z/z_p/m_p/logs_p [B,inter,F]. */
int bv2_infer_finish(bv2_engine* e, const float* noise_z, int64_t noise_ld, float noise_scale, int32_t max_len,
                     float* o, float* attn, float* y_mask, float* z, float* z_p, float* m_p, float* logs_p,
                     void* stream);

/* Same as bv2_infer_finish, but the waveform leaves the engine as 16-bit PCM o16 [B,1,Fg*hop] (int16), converted exactly as the
 * reference's callers do with gradio.processing_utils.convert_to_16_bit_wav on every infer() result (reference webui.py:86,
 * 129, 198; hiyoriUI.py:343): per utterance, over its valid samples, data / abs(data).max() * 32767 -> astype(int16) in
 * float32 (samples past the valid length and all-zero utterances give 0).  Halves the D2H / peer-store bytes. */
int bv2_infer_finish_pcm16(bv2_engine* e, const float* noise_z, int64_t noise_ld, float noise_scale, int32_t max_len,
                           int16_t* o16, float* attn, float* y_mask, float* z, float* z_p, float* m_p, float* logs_p,
                           void* stream);
/* The same conversion for a waveform batch the caller already holds: wave [B,L] fp32 (device), n_valid [B] int64 (device,
 * may be NULL = L) -> out [B,L] int16 (device). */
int bv2_wave_to_pcm16(bv2_engine* e, int B, int64_t L, const float* wave, const int64_t* n_valid, int16_t* out, void* stream);

/* attn [B,1,F,T] (reference commons.generate_path, commons.py:126-140; returned by infer() at models.py:1074) materialised
 * on demand from the durations of the last bv2_infer_begin: callers that never read it (infer.py:302-318 does not) skip the
 * O(F*T) write by passing attn = NULL to bv2_infer_finish.  Valid until the next bv2_infer_begin on this engine. */
int bv2_attn_path(bv2_engine* e, float* attn, void* stream);

/* Size the workspace for batches up to (B, T tokens, F_cap frames) up front: afterwards no call within those bounds
 * allocates or synchronises the device (the workspace otherwise grows geometrically on first use of a larger shape). */
int bv2_reserve(bv2_engine* e, int B, int T, int F_cap);

/* ---- per-stage entry points (parity tests + microbenchmarks; same kernels as the whole path) ---------------
 * text encoder: outputs x [B,H,T], m_p/logs_p [B,inter,T] (reference models.py:377-400)                       */
int bv2_text_encoder(bv2_engine* e, int B, int T, const int64_t* x, const int64_t* x_lengths, const int64_t* sid,
                     const int64_t* tone, const int64_t* language, const float* bert, const float* ja_bert,
                     const float* en_bert, float* x_out, float* m_out, float* logs_out, void* stream);
/* duration predictors on a given encoder output x [B,H,T]: logw_sdp/logw_dp [B,1,T]
 * (reference models.py:197-204,245-256 and 285-299)                                                         */
int bv2_duration(bv2_engine* e, int B, int T, const float* x, const int64_t* x_lengths, const int64_t* sid,
                 const float* noise_w, float noise_scale_w, float* logw_sdp, float* logw_dp, void* stream);
/* flow reverse on z_p [B,inter,F] -> z (reference models.py:142-145 / 442-445)                               */
int bv2_flow_reverse(bv2_engine* e, int B, int F, const float* z_p, const int64_t* y_lengths, const int64_t* sid,
                     float* z, void* stream);
/* Generator: z [B,inter,F], g [B,gin] -> o [B,1,F*hop] (reference models.py:538-557)                          */
int bv2_generator(bv2_engine* e, int B, int F, const float* z, const float* g, float* o, void* stream);

/* Debug tap: copy a named internal stage buffer of the LAST call (converted to [B,C,T]) to HOST memory.
 * Returns the number of floats written, or a negative status.                                               */
int64_t bv2_debug_read(bv2_engine* e, const char* name, float* host_out, int64_t capacity);

/* Stage timing (CUDA events recorded on the caller's stream around "encoder_duration", "flow", "generator"):
 * enable with bv2_set_profiling(e, 1); bv2_stage_ms blocks on the stage's end event and returns the duration of
 * the stage in the LAST call, or a negative value if it was not recorded. */
int bv2_set_profiling(bv2_engine* e, int enable);
float bv2_stage_ms(bv2_engine* e, const char* stage);

/* Counters: kernels launched by the engine since creation / bytes of workspace in use. */
int64_t bv2_launch_count(const bv2_engine* e);
int64_t bv2_workspace_bytes(const bv2_engine* e);
int64_t bv2_workspace_grows(const bv2_engine* e); /* (re)allocations of the workspace arenas since creation */

/* Peer output slab -- the path's one exchange step on a multi-GPU node (the reference has no equivalent:
 * its inference is single-device, webui.py:31, 397-399).  The root rank owns a device slab and exports its CUDA IPC
 * handle; every other rank (one process per GPU) opens it and passes a slice of the mapped address as the `o` output
 * pointer of bv2_infer_finish / bv2_generator, so the conv_post+tanh epilogue stores the finished waveform straight
 * into the root's memory over NVLink/NVSwitch (peer stores, no staging copy, no NCCL payload).
 *   bv2_peer_slab_alloc : cudaMalloc `bytes` on `cuda_device`, zero it, return the pointer and the 64-byte IPC handle
 *   bv2_peer_slab_open  : map a slab exported by another process into this one (lazy peer access)
 *   bv2_peer_slab_close : unmap (opener side);  bv2_peer_slab_free : release (owner side)
 *   bv2_peer_write      : async device->device copy of `bytes` into a (possibly peer-mapped) slab address on `stream`
 * All return BV2_OK or a negative status; none of them needs an engine. */
#define BV2_IPC_HANDLE_BYTES 64
int bv2_peer_slab_alloc(int cuda_device, int64_t bytes, void** dptr, unsigned char* handle_out);
int bv2_peer_slab_open(int cuda_device, const unsigned char* handle, void** dptr);
int bv2_peer_slab_close(int cuda_device, void* dptr);
int bv2_peer_slab_free(int cuda_device, void* dptr);
int bv2_peer_write(int cuda_device, void* dst, const void* src, int64_t bytes, void* stream);

const char* bv2_last_error(const bv2_engine* e);
const char* bv2_version(void);
void bv2_destroy(bv2_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* BV2_H_ */
