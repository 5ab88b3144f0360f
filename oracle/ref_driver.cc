// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// C wrapper over the UNMODIFIED reference implementation (ahmetaa/fast-dnn), which is
// compiled from the sources where they lie under /root/reference/src/cpp by oracle/Makefile
// into oracle/_ref/libfastdnn_ref.so.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load that library.
//
// The wrapper exposes the reference's public C++ classes stage by stage so that every
// intermediate of the hot path (quantized weights, LUT, last-hidden u8 activations, logits,
// softmax, lazy masked softmax) can be dumped and compared:
//   dnn::FloatDnn            /root/reference/src/cpp/float_dnn.cc:18-69
//   dnn::QuantizedDnn        /root/reference/src/cpp/dnn.cc:511-531
//   dnn::CalculationContext  /root/reference/src/cpp/dnn.cc:194-215, 402-454, 355-392
//
// `quantized_activations_` and the LUT are private members in the reference
// (dnn.h:45,197); this translation unit alone is compiled with private→public so they can be
// read.  The reference objects themselves are compiled untouched.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <iostream>
#include <cassert>
#include <cmath>
#include <x86intrin.h>

#define private public
#include "dnn.h"
#undef private

namespace dnn {
extern QuantizedSigmoid *qSigmoid;  // defined at dnn.cc:88
}
int ref_main(int argc, char *argv[]);  // dnn.cc:20 `main`, renamed by oracle/Makefile

namespace {

float *aligned_copy(const float *src, size_t count) {
  // The reference uses _mm_load_ps on the caller's buffer (dnn.cc:183,235) → 16-byte alignment.
  float *p = reinterpret_cast<float *>(dnn::aligned_malloc(16, sizeof(float) * std::max<size_t>(count, 4)));
  std::memcpy(p, src, sizeof(float) * count);
  return p;
}

}  // namespace

extern "C" {

void *ref_load(const char *path, float cutoff) {
  const dnn::FloatDnn float_dnn{std::string(path)};
  return new dnn::QuantizedDnn(float_dnn, cutoff);
}

void ref_free(void *h) { delete reinterpret_cast<dnn::QuantizedDnn *>(h); }

// The reference's own fp32 loader on its own (float_dnn.cc:18-69): what it reads back from a dnn.bin that OUR
// aligner / Kaldi importer wrote.
void *ref_float_load(const char *path) { return new dnn::FloatDnn(std::string(path)); }
void ref_float_free(void *h) { delete reinterpret_cast<dnn::FloatDnn *>(h); }
int ref_float_layer_count(void *h) { return (int) reinterpret_cast<dnn::FloatDnn *>(h)->layer_count(); }
int ref_float_layer_inputs(void *h, int i) { return (int) reinterpret_cast<dnn::FloatDnn *>(h)->layers()[i]->input_dimension(); }
int ref_float_layer_nodes(void *h, int i) { return (int) reinterpret_cast<dnn::FloatDnn *>(h)->layers()[i]->node_count(); }
void ref_float_layer(void *h, int i, float *w_out /*[nodes][inputs]*/, float *bias_out) {
  dnn::FloatLayer *l = reinterpret_cast<dnn::FloatDnn *>(h)->layers()[i];
  for (size_t n = 0; n < l->node_count(); ++n) std::memcpy(w_out + n * l->input_dimension(), l->weights()[n], l->input_dimension() * sizeof(float));
  std::memcpy(bias_out, l->bias(), l->node_count() * sizeof(float));
}
void ref_float_shift_scale(void *h, float *shift_out, float *scale_out) {
  auto *f = reinterpret_cast<dnn::FloatDnn *>(h);
  std::memcpy(shift_out, f->shift(), f->input_dimension() * sizeof(float));
  std::memcpy(scale_out, f->scale(), f->input_dimension() * sizeof(float));
}

int ref_input_dim(void *h) { return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->input_dimension(); }
int ref_output_dim(void *h) { return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->output_dimension(); }
// number of int8 layers (file layers − 1)
int ref_qlayer_count(void *h) { return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->layer_count(); }
int ref_hidden_dim(void *h) { return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->input_layer()->node_count(); }

int ref_qlayer_nodes(void *h, int i) {
  return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->layers()[i]->node_count();
}
int ref_qlayer_inputs(void *h, int i) {
  return (int) reinterpret_cast<dnn::QuantizedDnn *>(h)->layers()[i]->input_dimension();
}
float ref_qlayer_multiplier(void *h, int i) {
  return reinterpret_cast<dnn::QuantizedDnn *>(h)->layers()[i]->multiplier();
}
void ref_qlayer_weights(void *h, int i, int8_t *out) {
  auto *l = reinterpret_cast<dnn::QuantizedDnn *>(h)->layers()[i];
  std::memcpy(out, l->weights(), l->node_count() * l->input_dimension());
}
void ref_qlayer_bias(void *h, int i, float *out) {
  auto *l = reinterpret_cast<dnn::QuantizedDnn *>(h)->layers()[i];
  std::memcpy(out, l->bias(), l->node_count() * sizeof(float));
}
void ref_input_weights(void *h, float *w_out, float *bias_out, float *shift_out, float *scale_out) {
  auto *q = reinterpret_cast<dnn::QuantizedDnn *>(h);
  auto *l = q->input_layer();
  std::memcpy(w_out, l->weights(), l->node_count() * l->input_dimension() * sizeof(float));
  std::memcpy(bias_out, l->bias(), l->node_count() * sizeof(float));
  std::memcpy(shift_out, q->shift_, l->input_dimension() * sizeof(float));
  std::memcpy(scale_out, q->scale_, l->input_dimension() * sizeof(float));
}

void ref_sigmoid_lut(uint8_t *out /*[1280]*/) {
  std::memcpy(out, dnn::qSigmoid->lookup_, dnn::SIGMOID_LOOKUP_SIZE);
}
uint8_t ref_qsigmoid(float x) { return dnn::qSigmoid->get(x); }

// JNI `calculate` equivalent (jni_dnn.cc:35-62): per-call context, input not modified.
void ref_calculate(void *h, const float *in, int n, int dim, int batch, float *out) {
  auto *q = reinterpret_cast<dnn::QuantizedDnn *>(h);
  float *copy = aligned_copy(in, (size_t) n * dim);
  {
    dnn::BatchData data(copy, (size_t) n, (size_t) dim, true);  // owns + frees `copy`
    dnn::CalculationContext ctx(q, (size_t) n, (size_t) batch);
    dnn::BatchData *res = ctx.Calculate(data);
    std::memcpy(out, res->data(), (size_t) n * q->output_dimension() * sizeof(float));
    delete res;
  }
}

// Lazy context (jni_dnn.cc:64-126).
void *ref_ctx_new(void *h, int n, int batch) {
  return new dnn::CalculationContext(reinterpret_cast<dnn::QuantizedDnn *>(h), (size_t) n, (size_t) batch);
}
void ref_ctx_free(void *c) { delete reinterpret_cast<dnn::CalculationContext *>(c); }
void ref_ctx_until_output(void *c, const float *in) {
  auto *ctx = reinterpret_cast<dnn::CalculationContext *>(c);
  size_t n = ctx->input_count(), dim = ctx->dnn()->input_dimension();
  float *copy = aligned_copy(in, n * dim);
  dnn::BatchData data(copy, n, dim, true);
  ctx->CalculateUntilLastHiddenLayer(data);
}
// last-hidden u8 activations [n × H], frame-major (dnn.cc:207-208,269-270)
void ref_ctx_hidden(void *c, uint8_t *out) {
  auto *ctx = reinterpret_cast<dnn::CalculationContext *>(c);
  std::memcpy(out, ctx->quantized_activations_, ctx->hidden_node_count_ * ctx->input_count_);
}
// pre-bias dequantized output-layer activations [n × O] (what CalculateOutput computes at dnn.cc:434-437)
void ref_ctx_output_linear(void *c, float *out) {
  auto *ctx = reinterpret_cast<dnn::CalculationContext *>(c);
  size_t O = ctx->dnn()->output_dimension();
  for (size_t i = 0; i < ctx->input_count_; i += ctx->batch_size_) {
    ctx->QuantizedLayerActivations(*ctx->dnn()->output_layer(), i, &out[i * O]);
  }
}
void ref_ctx_lazy(void *c, int idx, const int8_t *mask, float *out) {
  auto *ctx = reinterpret_cast<dnn::CalculationContext *>(c);
  float *r = ctx->LazyOutputActivations((size_t) idx, reinterpret_cast<const char *>(mask));
  std::memcpy(out, r, ctx->dnn()->output_dimension() * sizeof(float));
}

// u8 activations after every hidden layer, stepping the reference's own public stage methods in
// the order CalculateUntilLastHiddenLayer does (dnn.cc:402-424).  out = [layers][n × H];
// layer 0 = after the fp32 input layer, layer j = after int8 layer j−1.
void ref_hidden_trace(void *h, const float *in, int n, int batch, uint8_t *out) {
  auto *q = reinterpret_cast<dnn::QuantizedDnn *>(h);
  size_t dim = q->input_dimension();
  float *copy = aligned_copy(in, (size_t) n * dim);
  dnn::BatchData data(copy, (size_t) n, dim, true);
  dnn::CalculationContext ctx(q, (size_t) n, (size_t) batch);
  size_t H = ctx.hidden_node_count_;
  q->ApplyShiftAndScale(data);
  for (size_t i = 0; i < (size_t) n; i += batch) {
    ctx.InputActivations(data, i);
    ctx.AddBias(q->input_layer()->bias());
    ctx.QuantizedSigmoid(i);
  }
  std::memcpy(out, ctx.quantized_activations_, H * n);
  for (size_t j = 0; j + 1 < q->layer_count(); ++j) {
    const dnn::QuantizedSimdLayer &layer = *q->layers()[j];
    for (size_t i = 0; i < (size_t) n; i += batch) {
      ctx.QuantizedLayerActivations(layer, i, ctx.activations_);
      ctx.AddBias(layer.bias());
      ctx.QuantizedSigmoid(i);
    }
    std::memcpy(out + (j + 1) * H * n, ctx.quantized_activations_, H * n);
  }
}

// CPU baseline timing: T threads, each with its own CalculationContext on a disjoint contiguous
// frame shard of one shared QuantizedDnn (test/java/suskun/nn/MultiThreadedStressTest.java:48-61).
// Timed region per thread = context construction + Calculate (what JNI calculate does,
// jni_dnn.cc:49-52).  Returns wall seconds for the slowest thread (steady_clock); the input
// copy is made outside the timed region.  If out != NULL the softmax rows are written there.
double ref_time_calculate(void *h, const float *in, int n, int dim, int batch, int threads, float *out) {
  auto *q = reinterpret_cast<dnn::QuantizedDnn *>(h);
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  size_t O = q->output_dimension();
  std::vector<float *> copies(threads);
  std::vector<int> begin(threads + 1);
  for (int t = 0; t <= threads; ++t) begin[t] = (int) ((long long) n * t / threads);
  for (int t = 0; t < threads; ++t)
    copies[t] = aligned_copy(in + (size_t) begin[t] * dim, (size_t) (begin[t + 1] - begin[t]) * dim);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> pool;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      ready.fetch_add(1);
      while (!go.load()) std::this_thread::yield();
      size_t cnt = (size_t) (begin[t + 1] - begin[t]);
      dnn::BatchData data(copies[t], cnt, (size_t) dim, true);
      dnn::CalculationContext ctx(q, cnt, (size_t) batch);
      dnn::BatchData *res = ctx.Calculate(data);
      if (out) std::memcpy(out + (size_t) begin[t] * O, res->data(), cnt * O * sizeof(float));
      delete res;
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// CPU baseline of the lazy path (BASELINE config 3): T threads, each with its own context over a disjoint frame shard
// of one shared QuantizedDnn; timed region per thread = context construction + CalculateUntilLastHiddenLayer + one
// LazyOutputActivations per frame with that frame's mask (the protocol of test/java/suskun/nn/FuncTest.java:92-133).
// masks = [n][O] bytes.  Returns wall seconds of the slowest thread.
double ref_time_lazy(void *h, const float *in, int n, int dim, int batch, const int8_t *masks, int threads, float *out) {
  auto *q = reinterpret_cast<dnn::QuantizedDnn *>(h);
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  size_t O = q->output_dimension();
  std::vector<float *> copies(threads);
  std::vector<int> begin(threads + 1);
  for (int t = 0; t <= threads; ++t) begin[t] = (int) ((long long) n * t / threads);
  for (int t = 0; t < threads; ++t)
    copies[t] = aligned_copy(in + (size_t) begin[t] * dim, (size_t) (begin[t + 1] - begin[t]) * dim);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      ready.fetch_add(1);
      while (!go.load()) std::this_thread::yield();
      size_t cnt = (size_t) (begin[t + 1] - begin[t]);
      dnn::BatchData data(copies[t], cnt, (size_t) dim, true);
      dnn::CalculationContext ctx(q, cnt, (size_t) batch);
      ctx.CalculateUntilLastHiddenLayer(data);
      float sink = 0;
      for (size_t i = 0; i < cnt; ++i) {
        float *r = ctx.LazyOutputActivations(i, reinterpret_cast<const char *>(masks + ((size_t) begin[t] + i) * O));
        if (out) std::memcpy(out + ((size_t) begin[t] + i) * O, r, O * sizeof(float));
        sink += r[0];
      }
      if (sink == 12345.678f) std::abort();
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  auto t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// The reference's command-line driver itself (dnn.cc:20-83, compiled with -Dmain=ref_main): model file, feature file →
// BatchData(input) → CalculationContext::Calculate → BatchData::dumpToFile(out, out_type == "BIN").  The file-to-file
// front end of the product (csrc/stream_file.cc) is checked against the files this writes.
int ref_cli(const char *model_path, const char *input_path, const char *out_path, const char *out_type) {
  std::string a0 = "fast-dnn", a1 = model_path, a2 = input_path, a3 = out_path, a4 = out_type;
  char *argv[5] = {&a0[0], &a1[0], &a2[0], &a3[0], &a4[0]};
  return ref_main(5, argv);
}

}  // extern "C"
