/* TEST INFRASTRUCTURE ONLY — CPU oracle for the fast-dnn quantized inference hot path.
 *
 * Plain-C restatement of the reference algorithm (ahmetaa/fast-dnn, src/cpp/dnn.cc and
 * src/cpp/float_dnn.cc); every function cites the reference lines it follows.  Parity pinned:
 * tests/test_oracle_vs_reference.py checks this port bit-for-bit against the compiled
 * reference (oracle/_ref) and against the committed golden vectors in tests/golden/ (which were
 * produced by the compiled reference, see tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (fast-dnn_b200/, libfast-dnn.so) never links, imports or calls it.
 */
#ifndef FDNN_ORACLE_H
#define FDNN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdo_net fdo_net;

/* float_dnn.cc:18-69 (parse) + dnn.cc:460-531 (quantize).  NULL on I/O or format error. */
fdo_net *fdo_load(const char *path, float cutoff);
void fdo_free(fdo_net *net);

int fdo_input_dim(const fdo_net *net);    /* padded to a multiple of 4 (float_dnn.cc:32-33) */
int fdo_output_dim(const fdo_net *net);
int fdo_hidden_dim(const fdo_net *net);   /* node count of file layer 0 */
int fdo_qlayer_count(const fdo_net *net); /* int8 layers = file layers − 1 (dnn.cc:516-520) */
int fdo_qlayer_nodes(const fdo_net *net, int i);
int fdo_qlayer_inputs(const fdo_net *net, int i);
float fdo_qlayer_multiplier(const fdo_net *net, int i);
const int8_t *fdo_qlayer_weights(const fdo_net *net, int i); /* [nodes][inputs] row-major */
const float *fdo_qlayer_bias(const fdo_net *net, int i);
const float *fdo_input_weights(const fdo_net *net);           /* [H][I] row-major */
const float *fdo_input_bias(const fdo_net *net);
const float *fdo_shift(const fdo_net *net);
const float *fdo_scale(const fdo_net *net);

/* dnn.cc:100-115 and dnn.h:35-42 */
void fdo_sigmoid_lut(uint8_t out[1280]);
uint8_t fdo_qsigmoid(float x);

/* dnn.cc:323-349 + 395-399: Σ over pairs of int16-saturated (a0·w0 + a1·w1), int32 accumulate */
int32_t fdo_node_sum(int K, const uint8_t *a, const int8_t *w);
/* same contraction without the int16 clamp (what a tensor-core IMMA computes); for tests */
int32_t fdo_node_sum_nosat(int K, const uint8_t *a, const int8_t *w);

/* dnn.cc:402-424 layer by layer.  out = [qlayer_count][n × H] u8: slot 0 = after the fp32 input
 * layer, slot j = after int8 layer j−1.  `in` is not modified. */
void fdo_hidden_trace(const fdo_net *net, const float *in, int n, uint8_t *out);
/* dnn.cc:402-424: last-hidden u8 activations [n × H] */
void fdo_until_output(const fdo_net *net, const float *in, int n, uint8_t *hidden, int threads);
/* dnn.cc:289-318 on the output layer: pre-bias dequantized activations [n × O] */
void fdo_output_linear(const fdo_net *net, const uint8_t *hidden, int n, float *lin, int threads);
/* dnn.cc:534-544 in place */
void fdo_softmax(float *row, int size);
/* dnn.cc:162-165 (Calculate) → softmax rows [n × O] */
void fdo_calculate(const fdo_net *net, const float *in, int n, float *out, int threads);
/* dnn.cc:355-392 for one frame given its last-hidden row; mask[O], nonzero = active */
void fdo_lazy(const fdo_net *net, const uint8_t *hidden_row, const int8_t *mask, float *out);

/* wall seconds of fdo_calculate over n frames with `threads` threads (for cpu_baseline "port") */
double fdo_time_calculate(const fdo_net *net, const float *in, int n, int threads, float *out);

#ifdef __cplusplus
}
#endif
#endif
