/* A JVM stand-in for timing the JNI entry point without a JVM (there is no JDK in the image).
 *
 * Drives Java_suskun_nn_QuantizedDnn_calculate of libfast-dnn.so the way the Java class does
 * (/root/reference/src/java/suskun/nn/QuantizedDnn.java:149-167 → jni_dnn.cc:35-62) from T host threads that share one
 * model handle (MultiThreadedStressTest.java:48-61), through a JNIEnv function table implemented here in C with what
 * HotSpot does for these calls: GetFloatArrayElements hands out a COPY of the array, NewFloatArray returns zero-filled heap
 * memory (a recycled, already-touched arena per thread, like a young generation), SetFloatArrayRegion is a memcpy into the
 * Java heap.  Slot numbers are those of the JNI specification (the same ones csrc/jni_min.h uses; tests/test_jni_shim.py
 * checks them against the reference's vendored jni.h).
 *
 *   gcc -O2 -shared -fPIC -o libjni_harness.so jni_harness.c -ldl -lpthread
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef const void *const *EnvTable;
typedef EnvTable *Env;

typedef struct {
  int len;
  int is_float;
  void *data;
} JArray;

enum { SLOT_FindClass = 6, SLOT_ThrowNew = 14, SLOT_GetStringUTFChars = 169, SLOT_ReleaseStringUTFChars = 170, SLOT_GetArrayLength = 171,
       SLOT_NewFloatArray = 181, SLOT_GetByteArrayElements = 184, SLOT_GetFloatArrayElements = 189, SLOT_ReleaseByteArrayElements = 192,
       SLOT_ReleaseFloatArrayElements = 197, SLOT_SetFloatArrayRegion = 213, TABLE_SIZE = 233 };

static __thread float *t_arena = NULL;   /* this thread's "young generation" for result arrays */
static __thread size_t t_arena_floats = 0;
static __thread JArray t_result;
static __thread int t_thrown = 0;

static void *j_FindClass(Env e, const char *name) { (void) e; (void) name; return (void *) 0x10; }
static int j_ThrowNew(Env e, void *cls, const char *msg) { (void) e; (void) cls; fprintf(stderr, "jni_harness: exception: %s\n", msg); t_thrown = 1; return 0; }
static const char *j_GetStringUTFChars(Env e, void *s, void *is_copy) { (void) e; (void) is_copy; return (const char *) s; }
static void j_ReleaseStringUTFChars(Env e, void *s, const char *c) { (void) e; (void) s; (void) c; }
static int j_GetArrayLength(Env e, void *a) { (void) e; return ((JArray *) a)->len; }
static void *j_NewFloatArray(Env e, int n) {
  (void) e;
  if ((size_t) n > t_arena_floats) {
    free(t_arena);
    t_arena = (float *) malloc((size_t) n * sizeof(float));
    t_arena_floats = (size_t) n;
  }
  memset(t_arena, 0, (size_t) n * sizeof(float)); /* Java arrays are born zeroed */
  t_result.len = n;
  t_result.is_float = 1;
  t_result.data = t_arena;
  return &t_result;
}
static void *j_GetArrayElements(Env e, void *a, void *is_copy) { /* HotSpot: a copy */
  (void) e; (void) is_copy;
  JArray *arr = (JArray *) a;
  size_t bytes = (size_t) arr->len * (arr->is_float ? 4 : 1);
  void *copy = malloc(bytes ? bytes : 1);
  memcpy(copy, arr->data, bytes);
  return copy;
}
static void j_ReleaseArrayElements(Env e, void *a, void *elems, int mode) { (void) e; (void) a; (void) mode; free(elems); }
static void j_SetFloatArrayRegion(Env e, void *a, int start, int n, const float *buf) {
  (void) e;
  memcpy((float *) ((JArray *) a)->data + start, buf, (size_t) n * sizeof(float));
}

static const void *g_table[TABLE_SIZE];
static EnvTable g_env_table = g_table;

static void init_table(void) {
  g_table[SLOT_FindClass] = (const void *) j_FindClass;
  g_table[SLOT_ThrowNew] = (const void *) j_ThrowNew;
  g_table[SLOT_GetStringUTFChars] = (const void *) j_GetStringUTFChars;
  g_table[SLOT_ReleaseStringUTFChars] = (const void *) j_ReleaseStringUTFChars;
  g_table[SLOT_GetArrayLength] = (const void *) j_GetArrayLength;
  g_table[SLOT_NewFloatArray] = (const void *) j_NewFloatArray;
  g_table[SLOT_GetByteArrayElements] = (const void *) j_GetArrayElements;
  g_table[SLOT_GetFloatArrayElements] = (const void *) j_GetArrayElements;
  g_table[SLOT_ReleaseByteArrayElements] = (const void *) j_ReleaseArrayElements;
  g_table[SLOT_ReleaseFloatArrayElements] = (const void *) j_ReleaseArrayElements;
  g_table[SLOT_SetFloatArrayRegion] = (const void *) j_SetFloatArrayRegion;
}

typedef int64_t (*init_fn)(Env, void *, void *, float);
typedef void *(*calc_fn)(Env, void *, int64_t, void *, int, int, int);
typedef void (*del_fn)(Env, void *, int64_t);

typedef struct {
  calc_fn calc;
  int64_t model;
  const float *frames;
  int n, dim, iters, failed;
  double checksum;
  pthread_barrier_t *start;
} Worker;

static void *worker_main(void *arg) {
  Worker *w = (Worker *) arg;
  JArray in = {w->n * w->dim, 1, (void *) w->frames};
  /* one untimed call: creates this thread's pooled workspace in the library and touches the arena */
  void *r = w->calc(&g_env_table, NULL, w->model, &in, w->n, w->dim, 10);
  if (!r || t_thrown) w->failed = 1;
  pthread_barrier_wait(w->start);
  pthread_barrier_wait(w->start); /* main thread reads the clock between the two */
  for (int i = 0; i < w->iters && !w->failed; ++i) {
    r = w->calc(&g_env_table, NULL, w->model, &in, w->n, w->dim, 10);
    if (!r || t_thrown) { w->failed = 1; break; }
    JArray *res = (JArray *) r;
    w->checksum += ((float *) res->data)[0] + ((float *) res->data)[res->len - 1];
  }
  free(t_arena);
  t_arena = NULL;
  t_arena_floats = 0;
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* seconds for `threads` callers × `iters` calls of calculate(n frames) on one shared model; < 0 on failure.
 * out_first (optional, n × outputs floats): the scores of one call, for checking. */
double jni_harness_calculate(const char *lib_path, const char *model_path, float cutoff, const float *frames, int n, int dim, int iters,
                             int threads, float *out_first, int out_floats) {
  init_table();
  void *lib = dlopen(lib_path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { fprintf(stderr, "jni_harness: %s\n", dlerror()); return -1.0; }
  init_fn init = (init_fn) dlsym(lib, "Java_suskun_nn_QuantizedDnn_initialize");
  calc_fn calc = (calc_fn) dlsym(lib, "Java_suskun_nn_QuantizedDnn_calculate");
  del_fn del = (del_fn) dlsym(lib, "Java_suskun_nn_QuantizedDnn_delete");
  if (!init || !calc || !del) return -2.0;
  int64_t model = init(&g_env_table, NULL, (void *) model_path, cutoff);
  if (!model || t_thrown) return -3.0;
  if (out_first) {
    JArray in = {n * dim, 1, (void *) frames};
    JArray *res = (JArray *) calc(&g_env_table, NULL, model, &in, n, dim, 10);
    if (!res || res->len != out_floats) return -4.0;
    memcpy(out_first, res->data, (size_t) out_floats * sizeof(float));
    free(t_arena);
    t_arena = NULL;
    t_arena_floats = 0;
  }
  if (threads < 1) threads = 1;
  pthread_barrier_t start;
  pthread_barrier_init(&start, NULL, (unsigned) threads + 1);
  Worker *w = (Worker *) calloc((size_t) threads, sizeof(Worker));
  pthread_t *tid = (pthread_t *) calloc((size_t) threads, sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) {
    w[t].calc = calc; w[t].model = model; w[t].frames = frames; w[t].n = n; w[t].dim = dim;
    w[t].iters = iters / threads + (t < iters % threads ? 1 : 0);
    w[t].start = &start;
    pthread_create(&tid[t], NULL, worker_main, &w[t]);
  }
  pthread_barrier_wait(&start);
  double t0 = now_s();
  pthread_barrier_wait(&start);
  int failed = 0;
  for (int t = 0; t < threads; ++t) { pthread_join(tid[t], NULL); failed |= w[t].failed; }
  double dt = now_s() - t0;
  del(&g_env_table, NULL, model);
  free(w);
  free(tid);
  pthread_barrier_destroy(&start);
  return failed ? -5.0 : dt;
}
