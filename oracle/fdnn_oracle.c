/* TEST INFRASTRUCTURE ONLY — see fdnn_oracle.h.  Scalar, no SIMD, no FMA (-ffp-contract=off). */
#define _POSIX_C_SOURCE 200809L
#include "fdnn_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  int nodes, inputs;
  int8_t *w;   /* [nodes][inputs] */
  float *bias; /* [nodes] */
  float multiplier;
} fdo_qlayer;

struct fdo_net {
  int in_dim;  /* padded ×4 */
  int hidden;  /* nodes of file layer 0 */
  float *w0;   /* [hidden][in_dim] */
  float *bias0;
  float *shift, *scale;
  int nq;
  fdo_qlayer *q;
};

static uint8_t g_lut[1280];
static pthread_once_t g_lut_once = PTHREAD_ONCE_INIT;

/* x86 cvttss2si / cvttsd2si semantics: out-of-range and NaN give INT_MIN ("integer indefinite").
 * The reference relies on this implicitly through static_cast<int>/<char> (dnn.h:36, dnn.cc:499). */
static int x86_trunc_to_int(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int) v;
}

/* dnn.cc:100-115 */
static void build_lut(void) {
  for (int i = -640; i < 640; ++i) {
    float k = i / 100.0f;
    float s = 1.0f / (1 + expf(-k));
    g_lut[i + 640] = (uint8_t) x86_trunc_to_int(roundf(s * 255.0f));
  }
}

void fdo_sigmoid_lut(uint8_t out[1280]) {
  pthread_once(&g_lut_once, build_lut);
  memcpy(out, g_lut, 1280);
}

/* dnn.h:35-42 — round() there is ::round(double) applied to the fp32 product input*100 */
uint8_t fdo_qsigmoid(float x) {
  pthread_once(&g_lut_once, build_lut);
  float t = x * 100;
  int k = x86_trunc_to_int(round((double) t));
  if (k <= -640) return 0;
  if (k >= 640) return 255;
  return g_lut[k + 640];
}

/* ---- loading: float_dnn.cc:166-212 (big-endian 4-byte words) ------------------------------- */
typedef struct {
  const unsigned char *p;
  size_t len, off;
  int bad;
} rd;

static uint32_t rd_u32(rd *r) {
  if (r->off + 4 > r->len) { r->bad = 1; return 0; }
  const unsigned char *b = r->p + r->off;
  r->off += 4;
  return ((uint32_t) b[0] << 24) | ((uint32_t) b[1] << 16) | ((uint32_t) b[2] << 8) | b[3];
}
static float rd_f32(rd *r) {
  uint32_t u = rd_u32(r);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static int pad_to(int num, int div) { /* float_dnn.cc:76-83 */
  int dif = div - num % div;
  return dif == div ? num : num + dif;
}

/* dnn.cc:148-160 */
static float abs_max(const float *f, size_t n, float lo, float hi) {
  float m = -3.402823466e+38F;
  for (size_t i = 0; i < n; ++i) {
    float v = f[i];
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    float a = (float) fabs(v);
    if (a > m) m = a;
  }
  return m;
}

/* dnn.cc:460-509.  Note the upper clip at :496-498 tests (minWeight > maxWeight) and is dead. */
static void quantize_layer(fdo_qlayer *q, const float *w, float cutoff) {
  float hi = cutoff, lo = -cutoff;
  float max = -3.402823466e+38F;
  for (int i = 0; i < q->nodes; ++i) {
    float nm = abs_max(w + (size_t) i * q->inputs, (size_t) q->inputs, lo, hi);
    if (nm > max) max = nm;
  }
  q->multiplier = roundf(127.0f / max);
  for (size_t i = 0; i < (size_t) q->nodes * q->inputs; ++i) {
    float f = w[i];
    if (f < lo) f = lo;
    if (lo > hi) f = hi;
    float prod = f * q->multiplier;
    q->w[i] = (int8_t) (uint8_t) (x86_trunc_to_int(roundf(prod)) & 0xff);
  }
}

void fdo_free(fdo_net *net) {
  if (!net) return;
  free(net->w0); free(net->bias0); free(net->shift); free(net->scale);
  for (int i = 0; i < net->nq; ++i) { free(net->q[i].w); free(net->q[i].bias); }
  free(net->q);
  free(net);
}

fdo_net *fdo_load(const char *path, float cutoff) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  rewind(f);
  unsigned char *buf = (unsigned char *) malloc(sz > 0 ? (size_t) sz : 1);
  if (!buf || fread(buf, 1, (size_t) sz, f) != (size_t) sz) { fclose(f); free(buf); return NULL; }
  fclose(f);
  rd r = {buf, (size_t) sz, 0, 0};

  int layers = (int) rd_u32(&r);
  if (r.bad || layers < 3 || layers > 1024) { free(buf); return NULL; } /* dnn.cc:199 needs ≥2 int8 layers */
  fdo_net *net = (fdo_net *) calloc(1, sizeof(fdo_net));
  net->nq = layers - 1;
  net->q = (fdo_qlayer *) calloc((size_t) net->nq, sizeof(fdo_qlayer));
  int actual_in = 0;
  for (int j = 0; j < layers && !r.bad; ++j) {
    int in = (int) rd_u32(&r), out = (int) rd_u32(&r);
    if (r.bad || in <= 0 || out <= 0 || (size_t) in * out > r.len) { r.bad = 1; break; }
    int pin = j == 0 ? pad_to(in, 4) : in;
    float *w = (float *) calloc((size_t) out * pin, sizeof(float));
    for (int o = 0; o < out; ++o)
      for (int i = 0; i < in; ++i) w[(size_t) o * pin + i] = rd_f32(&r);
    float *b = (float *) malloc((size_t) out * sizeof(float));
    for (int o = 0; o < out; ++o) b[o] = rd_f32(&r);
    if (j == 0) {
      actual_in = in;
      net->in_dim = pin; net->hidden = out; net->w0 = w; net->bias0 = b;
    } else {
      fdo_qlayer *q = &net->q[j - 1];
      q->nodes = out; q->inputs = in; q->bias = b;
      q->w = (int8_t *) malloc((size_t) out * in);
      quantize_layer(q, w, cutoff);
      free(w);
    }
  }
  if (!r.bad) {
    net->shift = (float *) calloc((size_t) net->in_dim, sizeof(float));
    net->scale = (float *) calloc((size_t) net->in_dim, sizeof(float));
    for (int i = 0; i < actual_in; ++i) net->shift[i] = rd_f32(&r);
    for (int i = 0; i < actual_in; ++i) net->scale[i] = rd_f32(&r);
  }
  free(buf);
  if (r.bad) { fdo_free(net); return NULL; }
  return net;
}

int fdo_input_dim(const fdo_net *n) { return n->in_dim; }
int fdo_output_dim(const fdo_net *n) { return n->q[n->nq - 1].nodes; }
int fdo_hidden_dim(const fdo_net *n) { return n->hidden; }
int fdo_qlayer_count(const fdo_net *n) { return n->nq; }
int fdo_qlayer_nodes(const fdo_net *n, int i) { return n->q[i].nodes; }
int fdo_qlayer_inputs(const fdo_net *n, int i) { return n->q[i].inputs; }
float fdo_qlayer_multiplier(const fdo_net *n, int i) { return n->q[i].multiplier; }
const int8_t *fdo_qlayer_weights(const fdo_net *n, int i) { return n->q[i].w; }
const float *fdo_qlayer_bias(const fdo_net *n, int i) { return n->q[i].bias; }
const float *fdo_input_weights(const fdo_net *n) { return n->w0; }
const float *fdo_input_bias(const fdo_net *n) { return n->bias0; }
const float *fdo_shift(const fdo_net *n) { return n->shift; }
const float *fdo_scale(const fdo_net *n) { return n->scale; }

/* ---- arithmetic ---------------------------------------------------------------------------- */

/* dnn.cc:323-349: pmaddubsw saturates each adjacent-pair sum to int16 before widening. */
int32_t fdo_node_sum(int K, const uint8_t *a, const int8_t *w) {
  int32_t s = 0;
  for (int k = 0; k < K; k += 2) {
    int32_t p = (int32_t) a[k] * w[k] + (int32_t) a[k + 1] * w[k + 1];
    if (p > 32767) p = 32767;
    if (p < -32768) p = -32768;
    s += p;
  }
  return s;
}

int32_t fdo_node_sum_nosat(int K, const uint8_t *a, const int8_t *w) {
  int32_t s = 0;
  for (int k = 0; k < K; ++k) s += (int32_t) a[k] * w[k];
  return s;
}

/* dnn.cc:175-192 (add, then mul) followed by dnn.cc:219-247 + 168-172: four lane partial sums
 * over k ≡ lane (mod 4), product rounded then added, combined as (l0+l1)+(l2+l3); then
 * dnn.cc:250-264 bias add and dnn.cc:267-286 LUT. */
static void input_layer_frame(const fdo_net *net, const float *x_raw, float *x, uint8_t *a) {
  int I = net->in_dim, H = net->hidden;
  for (int k = 0; k < I; ++k) {
    float v = x_raw[k] + net->shift[k];
    x[k] = v * net->scale[k];
  }
  for (int n = 0; n < H; ++n) {
    const float *w = net->w0 + (size_t) n * I;
    float l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    for (int k = 0; k < I; k += 4) {
      float m0 = x[k] * w[k], m1 = x[k + 1] * w[k + 1], m2 = x[k + 2] * w[k + 2], m3 = x[k + 3] * w[k + 3];
      l0 = l0 + m0; l1 = l1 + m1; l2 = l2 + m2; l3 = l3 + m3;
    }
    float h = (l0 + l1) + (l2 + l3);
    a[n] = fdo_qsigmoid(h + net->bias0[n]);
  }
}

/* dnn.cc:289-318: (float)sum / (multiplier*255.0f), IEEE division by an fp32 product */
static float dequant(int32_t s, float multiplier) {
  float coeff = multiplier * 255.0f;
  return (float) s / coeff;
}

static void hidden_layer_frame(const fdo_qlayer *q, const uint8_t *a, uint8_t *out) {
  for (int n = 0; n < q->nodes; ++n) {
    float lin = dequant(fdo_node_sum(q->inputs, a, q->w + (size_t) n * q->inputs), q->multiplier);
    out[n] = fdo_qsigmoid(lin + q->bias[n]);
  }
}

void fdo_hidden_trace(const fdo_net *net, const float *in, int n, uint8_t *out) {
  int I = net->in_dim, H = net->hidden;
  float *x = (float *) malloc((size_t) I * sizeof(float));
  for (int f = 0; f < n; ++f) {
    input_layer_frame(net, in + (size_t) f * I, x, out + (size_t) f * H);
    for (int j = 0; j + 1 < net->nq; ++j)
      hidden_layer_frame(&net->q[j], out + ((size_t) j * n + f) * H, out + ((size_t) (j + 1) * n + f) * H);
  }
  free(x);
}

/* dnn.cc:534-544: no max subtraction; sequential fp32 sum from 0 */
void fdo_softmax(float *row, int size) {
  float total = 0;
  for (int i = 0; i < size; ++i) {
    float d = expf(row[i]);
    row[i] = d;
    total += d;
  }
  for (int i = 0; i < size; ++i) row[i] = row[i] / total;
}

void fdo_lazy(const fdo_net *net, const uint8_t *hidden_row, const int8_t *mask, float *out) {
  const fdo_qlayer *q = &net->q[net->nq - 1];
  for (int i = 0; i < q->nodes; ++i) {
    if (mask[i] == 0) { out[i] = 0; continue; }
    float lin = dequant(fdo_node_sum(q->inputs, hidden_row, q->w + (size_t) i * q->inputs), q->multiplier);
    out[i] = lin + q->bias[i];
  }
  fdo_softmax(out, q->nodes);
}

/* ---- frame-parallel drivers ----------------------------------------------------------------- */
typedef struct {
  const fdo_net *net;
  const float *in;
  const uint8_t *hid_in;
  uint8_t *hid_out;
  float *out;
  int begin, end, mode; /* 0 until_output, 1 output_linear, 2 calculate */
} job;

static void *worker(void *arg) {
  job *j = (job *) arg;
  const fdo_net *net = j->net;
  int I = net->in_dim, H = net->hidden, O = fdo_output_dim(net);
  const fdo_qlayer *ql = &net->q[net->nq - 1];
  float *x = (float *) malloc((size_t) I * sizeof(float));
  uint8_t *a = (uint8_t *) malloc((size_t) H), *b = (uint8_t *) malloc((size_t) H);
  for (int f = j->begin; f < j->end; ++f) {
    const uint8_t *last;
    if (j->mode == 1) {
      last = j->hid_in + (size_t) f * H;
    } else {
      uint8_t *cur = a, *nxt = b;
      input_layer_frame(net, j->in + (size_t) f * I, x, cur);
      for (int l = 0; l + 1 < net->nq; ++l) {
        hidden_layer_frame(&net->q[l], cur, nxt);
        uint8_t *t = cur; cur = nxt; nxt = t;
      }
      last = cur;
      if (j->hid_out) memcpy(j->hid_out + (size_t) f * H, last, (size_t) H);
    }
    if (j->mode >= 1) {
      float *row = j->out + (size_t) f * O;
      for (int n = 0; n < O; ++n)
        row[n] = dequant(fdo_node_sum(ql->inputs, last, ql->w + (size_t) n * ql->inputs), ql->multiplier);
      if (j->mode == 2) { /* dnn.cc:442-449 */
        for (int n = 0; n < O; ++n) row[n] += ql->bias[n];
        fdo_softmax(row, O);
      }
    }
  }
  free(x); free(a); free(b);
  return NULL;
}

static void run(job base, int n, int threads) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) threads);
  job *jobs = (job *) malloc(sizeof(job) * (size_t) threads);
  for (int t = 0; t < threads; ++t) {
    jobs[t] = base;
    jobs[t].begin = (int) ((long long) n * t / threads);
    jobs[t].end = (int) ((long long) n * (t + 1) / threads);
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

void fdo_until_output(const fdo_net *net, const float *in, int n, uint8_t *hidden, int threads) {
  job j = {net, in, NULL, hidden, NULL, 0, 0, 0};
  run(j, n, threads);
}
void fdo_output_linear(const fdo_net *net, const uint8_t *hidden, int n, float *lin, int threads) {
  job j = {net, NULL, hidden, NULL, lin, 0, 0, 1};
  run(j, n, threads);
}
void fdo_calculate(const fdo_net *net, const float *in, int n, float *out, int threads) {
  job j = {net, in, NULL, NULL, out, 0, 0, 2};
  run(j, n, threads);
}

double fdo_time_calculate(const fdo_net *net, const float *in, int n, int threads, float *out) {
  float *tmp = out ? out : (float *) malloc((size_t) n * fdo_output_dim(net) * sizeof(float));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  fdo_calculate(net, in, n, tmp, threads);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (!out) free(tmp);
  return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}
