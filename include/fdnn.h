/* fdnn.h — C ABI of libfast-dnn.so, the B200-native drop-in for fast-dnn's quantized inference path.
 *
 * This is the boundary the reference's JNI layer binds.  The eleven
 * Java_suskun_nn_QuantizedDnn_* symbols exported by the same library (csrc/jni_shim.cc) are thin
 * wrappers over these entry points; INTEGRATION.md shows the binding.  Every entry point cites the
 * reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions: plain pointers and sizes only; every function returns FDNN_OK (0) or a negative
 * FDNN_E* code and never throws or exits across the ABI (the reference crashes or exit(3)s on
 * failure, float_dnn.cc:171-188); fdnn_last_error() returns a thread-local message.  Caller input
 * buffers are never modified (the reference scales its private copy in place, dnn.cc:175-192,
 * jni_dnn.cc:44,58).  A model handle is immutable and may be shared by any number of host threads
 * (test/java/suskun/nn/MultiThreadedStressTest.java:48-67); a context belongs to one thread at a time.
 *
 * There is no CPU fallback: entry points that compute return FDNN_ENOGPU when no CUDA device is
 * usable.  fdnn_pack / fdnn_blob_* are host-only and work without a GPU.
 */
#ifndef FDNN_H
#define FDNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDNN_OK 0
#define FDNN_EINVAL (-1)  /* bad argument */
#define FDNN_EIO (-2)     /* file missing / unreadable / truncated */
#define FDNN_EFORMAT (-3) /* not a usable dnn.bin (see constraints below) */
#define FDNN_ENOGPU (-4)  /* no usable CUDA device; there is no CPU fallback */
#define FDNN_ECUDA (-5)   /* CUDA runtime error, text in fdnn_last_error() */
#define FDNN_ENOMEM (-6)

typedef struct fdnn_model fdnn_model; /* replaces dnn::QuantizedDnn*        (src/cpp/dnn.h:106-142) */
typedef struct fdnn_ctx fdnn_ctx;     /* replaces dnn::CalculationContext*  (src/cpp/dnn.h:144-208) */

const char *fdnn_last_error(void);
const char *fdnn_version(void);

/* ---- load (one-time) ------------------------------------------------------------------------
 * Network constraints inherited from the reference (dnn.cc:199,254,331; README.md:10,69): at least
 * two int8 layers (≥3 layers in the file), all hidden layers the same width, hidden width a
 * multiple of 16, any number of outputs.  Layer-0 input is zero-padded to a multiple of 4. */

/* Java_suskun_nn_QuantizedDnn_initialize (src/cpp/jni_dnn.cc:7-18): parse dnn.bin
 * (float_dnn.cc:18-69), quantize layers 1.. to int8 (dnn.cc:460-509), upload to CUDA `device`
 * (−1 = the calling thread's current device). */
int fdnn_load(const char *path, float cutoff, int device, fdnn_model **out);

/* Device group behind ONE handle (no counterpart in the reference, which is single-threaded CPU code):
 * the host parses and quantizes once, the packed blob is uploaded to devices[0] and delivered to the other devices by ONE
 * in-process ncclBroadcast (libnccl.so.2 is bound at run time; a group without it fails with FDNN_ECUDA), and every device
 * builds a replica.  fdnn_calculate on the returned handle shards its frames contiguously over the devices (whole tiles of
 * 128 frames, no per-frame collective); fdnn_ctx_new hands contexts out round-robin, one context lives on one device.
 * fdnn_load(path, cutoff, -1, …) does the same for the devices named by the environment variable FDNN_DEVICES
 * ("all" or "0,1,2,3"), which is how the JNI entry point — whose signature has no device argument, jni_dnn.cc:7-18 —
 * gets more than one GPU. */
int fdnn_load_devices(const char *path, float cutoff, const int *devices, int n_devices, fdnn_model **out);
int fdnn_device_count(const fdnn_model *model);      /* 1 unless the handle is a device group */
int fdnn_device_at(const fdnn_model *model, int i);  /* CUDA ordinal of the i-th device of the group */
long long fdnn_nccl_broadcast_count(void);           /* ncclBroadcast collectives issued by this library since load */
/* Host-only: how fdnn_calculate cuts a call of n_frames over n_devices — contiguous shards in device order, whole tiles of
 * 128 frames, sizes differing by at most one tile; first[d], count[d] for every device (count 0 = unused); returns the number
 * of devices used. */
int fdnn_shard_plan(int n_frames, int n_devices, int *first, int *count);

/* Host-only half of fdnn_load: parse + quantize into one relocatable blob (weights, biases,
 * multipliers, sigmoid LUT, saturation fix-up lists).  In a multi-GPU job rank 0 packs, the blob
 * is broadcast once (NCCL), and every rank calls fdnn_load_blob — SURVEY.md §8e. */
int fdnn_pack(const char *path, float cutoff, void **blob, size_t *size);
void fdnn_blob_free(void *blob);
/* `blob` may be a host pointer or a device pointer on `device` (e.g. an NCCL receive buffer). */
int fdnn_load_blob(const void *blob, size_t size, int device, fdnn_model **out);

/* ---- offline tooling either side of the path (host only, no GPU needed) -------------------------
 * FeedForwardNetwork.align + saveBinary (src/java/suskun/nn/FeedForwardNetwork.java:50-58,226-235,
 * 264-281,331-340): read a dnn.bin of any shape, zero-pad the input width to a multiple of
 * input_alignment (4) and every hidden width to a multiple of hidden_alignment (16), write a dnn.bin
 * that fdnn_load accepts.  Alignments are 1 … 4096 (FDNN_EINVAL otherwise).  The reference does this in Java only
 * (README.md:76 lists C++ as a TODO). */
int fdnn_align_dnn_bin(const char *in_path, const char *out_path, int input_alignment, int hidden_alignment);
/* FeedForwardNetwork.loadFromTextFile + saveBinary (FeedForwardNetwork.java:86-119,159-207,226-235): Kaldi nnet1
 * text model ("<AffineTransform> out in" blocks) plus the feature-transform text (optional <Splice> block,
 * shift, scale) → an unaligned dnn.bin; follow with fdnn_align_dnn_bin.  Java-only in the reference. */
int fdnn_import_kaldi_nnet1(const char *nnet_txt_path, const char *transform_txt_path, const char *out_dnn_bin_path);
/* Feature matrices: big-endian int32 frames, int32 dim, fp32 rows (BatchData.java:80-91,107-139;
 * float_dnn.cc:85-105).  *data is malloc'ed (release with fdnn_blob_free). */
int fdnn_feature_bin_read(const char *path, int *frames, int *dim, float **data);
int fdnn_feature_bin_write(const char *path, const float *data, int frames, int dim);
/* BatchData::dumpToFile(..., binary) (float_dnn.cc:128-164): native-endian uint32 n, uint32 d, fp32 rows */
int fdnn_output_dump_write(const char *path, const float *data, int frames, int dim);
/* BatchData::dumpToFile(..., binary = false): a row per line, values printed by `ostream << float` ("%g"), one blank between */
int fdnn_output_dump_write_txt(const char *path, const float *data, int frames, int dim);

/* Java_suskun_nn_QuantizedDnn_delete (jni_dnn.cc:128-133) */
int fdnn_free(fdnn_model *model);

/* inputDimension / outputDimension (jni_dnn.cc:20-33): padded input width, output width */
int fdnn_input_dim(const fdnn_model *model);
int fdnn_output_dim(const fdnn_model *model);
/* layerCount (jni_dnn.cc:150-156): number of layers in the file */
int fdnn_layer_count(const fdnn_model *model);
/* layerDimension (jni_dnn.cc:135-148): i==0 → nodes of file layer 0; i≥1 → nodes of file layer
 * i+1 (the reference indexes its int8 layer vector with i); −1 where the reference returns −1 or
 * reads out of bounds. */
int fdnn_layer_dim(const fdnn_model *model, int i);
int fdnn_hidden_dim(const fdnn_model *model);
int fdnn_device(const fdnn_model *model);
/* Tile policy of the tensor-core layers for SHORT batches (no counterpart in the reference, whose batchSize only blocks for
 * the CPU cache).  FDNN_POLICY_LATENCY (default): the narrowest tiles that fill the GPU in one wave — shortest time for
 * one caller.  FDNN_POLICY_THROUGHPUT: one step wider tiles, half as many CTAs that each move fewer operand bytes per
 * result — less SM time per frame, for callers that keep several contexts in flight on one model (the pattern of
 * MultiThreadedStressTest.java:48-61).  Results are identical.  A context takes the policy in force when it is CREATED
 * (fdnn_ctx_new; the workspaces fdnn_calculate pools are contexts too) and keeps it for life. */
#define FDNN_POLICY_LATENCY 0
#define FDNN_POLICY_THROUGHPUT 1
int fdnn_set_tile_policy(fdnn_model *model, int policy);

/* ---- full forward ---------------------------------------------------------------------------
 * Java_suskun_nn_QuantizedDnn_calculate (jni_dnn.cc:35-62) = CalculationContext::Calculate
 * (dnn.cc:162-165).  in: host [n × dim] fp32 row-major, dim must equal fdnn_input_dim; out: host
 * [n × output_dim] fp32 softmax rows.  Host pointers may be pageable or pinned (fdnn_host_alloc);
 * frames are streamed through the GPU in chunks with copies overlapped.  batch_hint is the
 * reference's CPU cache-blocking batchSize; results do not depend on it and it is ignored.
 * Re-entrant: concurrent calls on one model each use a private workspace. n == 0 is a no-op. */
int fdnn_calculate(fdnn_model *model, const float *in, int n, int dim, int batch_hint, float *out);
/* Same computation, results delivered piecewise: `sink` is called on the CALLING thread with rows [first_frame,
 * first_frame + n_frames) in page-locked memory that is only valid during the call — in frame order on one device, the
 * shards of a device group interleaved (each in order); every frame exactly once; a non-zero return aborts.  This is what Java_suskun_nn_QuantizedDnn_calculate uses to move scores from the transfer buffer straight into
 * the Java array (SetFloatArrayRegion) while the next sub-chunk is still crossing PCIe, instead of the reference's
 * malloc + copy + copy (jni_dnn.cc:49-58). */
typedef int (*fdnn_sink_fn)(void *user, int first_frame, int n_frames, const float *rows);
int fdnn_calculate_sink(fdnn_model *model, const float *in, int n, int dim, fdnn_sink_fn sink, void *user);

/* File to file: the data path of the reference's command-line driver (src/cpp/dnn.cc:55-78: BatchData(input_path) →
 * CalculationContext::Calculate → BatchData::dumpToFile) without holding either file in memory (BASELINE
 * config 4 is 1.76 GB of features and 32 GB of scores).  The feature file (big-endian int32 frames, int32 dim, fp32 rows,
 * float_dnn.cc:85-105) is read in chunks of chunk_frames (0 = 4096 per device; 1024 for the text dump) by a reader thread; the calling thread sends each
 * chunk through fdnn_calculate_sink and appends the scores, piece by piece as they come off the GPU, to out_path as the reference's
 * binary dump (FDNN_DUMP_BIN: native-endian uint32 frames, uint32 dim, fp32 rows) or text dump (FDNN_DUMP_TXT: a row per line,
 * values as printed by `ostream << float`), float_dnn.cc:128-164.  The file's dim must be the network's input width or its
 * unpadded width (at most 3 columns fewer, float_dnn.cc:32-33; the missing columns are zeros).  *frames_done (may be NULL) = rows
 * written, also on failure. */
#define FDNN_DUMP_BIN 1
#define FDNN_DUMP_TXT 0
int fdnn_calculate_file(fdnn_model *model, const char *feature_bin_path, const char *out_path, int out_format, int chunk_frames,
                        long long *frames_done);

/* ---- contexts: lazy output + device-resident pipelines --------------------------------------
 * Java_suskun_nn_QuantizedDnn_getContext (jni_dnn.cc:64-77): workspace for exactly n frames. */
int fdnn_ctx_new(fdnn_model *model, int n, int batch_hint, fdnn_ctx **out);
/* deleteLazyContext (jni_dnn.cc:119-126) */
int fdnn_ctx_free(fdnn_ctx *ctx);
int fdnn_ctx_frames(const fdnn_ctx *ctx);
int fdnn_ctx_output_dim(const fdnn_ctx *ctx);
int fdnn_ctx_input_dim(const fdnn_ctx *ctx);

/* calculateUntilOutput (jni_dnn.cc:79-95) = CalculateUntilLastHiddenLayer (dnn.cc:402-424).
 * in: host [ctx.n × input_dim].  Also computes the dense output-layer logits on the device so
 * that every later fdnn_ctx_lazy call is a masked softmax over resident data. */
int fdnn_ctx_until_output(fdnn_ctx *ctx, const float *in);
/* calculateLazy (jni_dnn.cc:97-117) = LazyOutputActivations (dnn.cc:355-392): mask[O], nonzero =
 * active; inactive logits are 0 and still enter the softmax denominator.  out: host [O]. */
int fdnn_ctx_lazy(fdnn_ctx *ctx, int idx, const int8_t *mask, float *out);
/* All ctx.n frames at once: masks host [n × O], out host [n × O] (BASELINE config 3). */
int fdnn_ctx_lazy_batch(fdnn_ctx *ctx, const int8_t *masks, float *out);

/* Device-resident variants: pointers are device memory on the model's device, work is enqueued on
 * `stream` (a cudaStream_t; NULL = the legacy default stream) and NOT synchronised.  n_frames ≤
 * ctx.n.  These are what a GPU-side pipeline calls and what bench.py times for `value`. */
int fdnn_ctx_forward_device(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, void *stream);
int fdnn_ctx_until_output_device(fdnn_ctx *ctx, const float *d_in, int n_frames, void *stream);
int fdnn_ctx_lazy_batch_device(fdnn_ctx *ctx, const int8_t *d_masks, int n_frames, float *d_out, void *stream);

/* ---- inspection (stage-level parity tests; not on the hot path) ------------------------------ */
/* u8 activations after hidden layer `layer` (0 = fp32 input layer … qlayers−2 = last hidden) of
 * the most recent forward on this context; host out [n_frames × H]. Only the last hidden layer
 * is retained unless the context was put in trace mode before the forward. */
int fdnn_ctx_set_trace(fdnn_ctx *ctx, int enable);
int fdnn_ctx_hidden(fdnn_ctx *ctx, int layer, int n_frames, uint8_t *out);
/* position-weighted 64-bit checksum of the same bytes, computed on the device (soak tests over more frames than are
 * worth copying back): equal digests ⇔ equal bytes up to a 2^-64 coincidence; any single-byte change changes it */
int fdnn_ctx_hidden_digest(fdnn_ctx *ctx, int layer, int n_frames, unsigned long long *digest);
/* dense output logits (dequantized + bias, before softmax) of the most recent until_output */
int fdnn_ctx_logits(fdnn_ctx *ctx, int n_frames, float *out);
/* quantized layer i (0-based among int8 layers): any of the out pointers may be NULL */
int fdnn_model_qlayer(const fdnn_model *model, int i, int *nodes, int *inputs, float *multiplier,
                      int8_t *weights, float *bias);
/* number of weight pairs per int8 layer that can trip pmaddubsw's int16 saturation (dnn.cc:337-340) */
int fdnn_model_fixup_count(const fdnn_model *model, int i);
/* 1 if layer i's epilogue uses the 3-op division that was exhaustively verified against IEEE
 * division for this layer's divisor at load, 0 if it uses __fdiv_rn */
int fdnn_model_fast_div(const fdnn_model *model, int i);
/* 1 if int8 layer i runs on the tcgen05 tensor-core kernel, 0 if on the dp4a kernel (narrow layers) */
int fdnn_model_uses_tensor_cores(const fdnn_model *model, int i);
int fdnn_sigmoid_lut(uint8_t out[1280]);

/* Bench/profiling aid (not on the hot path): `iters` forward passes over device-resident input on
 * the context's own stream with CUDA events between the kernels.  ms[0] = fp32 input layer,
 * ms[1..L-2] = hidden int8 layers, ms[L-1] = output int8 layer, ms[L] = softmax (L = layer count);
 * averages in milliseconds, synchronised on return. */
int fdnn_ctx_profile_stages(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, int iters, float *ms);

/* The pass as it normally runs (for batches that are one wave of tiles: input layer + ONE fused kernel for all int8 layers and
 * the softmax, csrc/qlayer_fused.cu): ms[0] = fp32 input layer, ms[1] = everything after it; *fused = 1 when that was the
 * single fused kernel.  fdnn_ctx_profile_stages above always times the layer-by-layer kernels. */
int fdnn_ctx_profile_pass(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, int iters, float *ms, int *fused);

/* Diagnostics of the certified tensor-core input layer (csrc/input_tc.cu): how many (frame, node) elements of the last pass
 * the error-bound certificate left to the exact CUDA-core path; *undecided = 0xffffffff when the context uses the plain
 * exact kernel.  Synchronises the context's stream. */
int fdnn_ctx_input_undecided(fdnn_ctx *ctx, unsigned *undecided);
/* Profiling aid: per-CTA phase timestamps of the tensor-core layer kernels.  enable=1 arms the
 * next forward pass; enable=0 copies [layers-1][1024][8] uint64 SM-clock stamps to `out` and disarms. */
int fdnn_ctx_timeline(fdnn_ctx *ctx, int enable, unsigned long long *out);

/* ---- pinned host memory for callers that want zero staging ---------------------------------- */
int fdnn_host_alloc(void **ptr, size_t bytes);
int fdnn_host_free(void *ptr);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
long long fdnn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FDNN_H */
