// The reference's JNI surface, re-pointed at the C ABI (include/fdnn.h).  Symbol names and JNI
// signatures are exactly those of /root/reference/src/cpp/suskun_nn_QuantizedDnn.h:15-96, so the
// unmodified Java class suskun.nn.QuantizedDnn (src/java/suskun/nn/QuantizedDnn.java:109-127)
// binds to this library when it is packaged as /resources/libfast-dnn.so (QuantizedDnn.java:31).
// Each function mirrors its counterpart in /root/reference/src/cpp/jni_dnn.cc (lines cited).
//
// Difference in failure behaviour: the reference crashes or exit(3)s; here a failed call throws
// java.lang.IllegalStateException carrying fdnn_last_error() and returns 0 / null.

#include <cstdlib>
#include <cstring>

#include "../../include/fdnn.h"
#include "jni_min.h"

namespace {

template <class Fn>
Fn slot(JNIEnvPtr env, int index) {
  return reinterpret_cast<Fn>(const_cast<void *>((*env)[index]));
}

void throw_state(JNIEnvPtr env, const char *what) {
  auto find = slot<jclass (*)(JNIEnvPtr, const char *)>(env, kJniFindClass);
  auto thrw = slot<jint (*)(JNIEnvPtr, jclass, const char *)>(env, kJniThrowNew);
  jclass cls = find(env, "java/lang/IllegalStateException");
  if (cls) thrw(env, cls, what);
}

jfloatArray to_java(JNIEnvPtr env, const float *data, jsize len) {
  auto mk = slot<jfloatArray (*)(JNIEnvPtr, jsize)>(env, kJniNewFloatArray);
  auto set = slot<void (*)(JNIEnvPtr, jfloatArray, jsize, jsize, const jfloat *)>(env, kJniSetFloatArrayRegion);
  jfloatArray arr = mk(env, len);
  if (arr) set(env, arr, 0, len, data);
  return arr;
}

}  // namespace

extern "C" {

// jni_dnn.cc:7-18
FDNN_JNIEXPORT jlong Java_suskun_nn_QuantizedDnn_initialize(JNIEnvPtr env, jobject, jstring path, jfloat cutoff) {
  auto get = slot<const char *(*) (JNIEnvPtr, jstring, jboolean *)>(env, kJniGetStringUTFChars);
  auto rel = slot<void (*)(JNIEnvPtr, jstring, const char *)>(env, kJniReleaseStringUTFChars);
  const char *chars = get(env, path, nullptr);
  if (!chars) return 0;  // OutOfMemoryError is already pending
  fdnn_model *model = nullptr;
  int rc = fdnn_load(chars, cutoff, -1, &model);
  rel(env, path, chars);
  if (rc != FDNN_OK) {
    throw_state(env, fdnn_last_error());
    return 0;
  }
  return reinterpret_cast<jlong>(model);
}

// jni_dnn.cc:20-25
FDNN_JNIEXPORT jint Java_suskun_nn_QuantizedDnn_inputDimension(JNIEnvPtr, jobject, jlong handle) {
  return fdnn_input_dim(reinterpret_cast<fdnn_model *>(handle));
}

// jni_dnn.cc:27-33
FDNN_JNIEXPORT jint Java_suskun_nn_QuantizedDnn_outputDimension(JNIEnvPtr, jobject, jlong handle) {
  return fdnn_output_dim(reinterpret_cast<fdnn_model *>(handle));
}

// jni_dnn.cc:35-62: the Java array is read, never written back (JNI_ABORT).  The reference mallocs a result buffer,
// computes into it and copies it into a new float[]; here the scores go from the library's page-locked transfer buffer
// straight into the float[] (SetFloatArrayRegion per 128-frame piece, on this thread, while later pieces are still
// crossing PCIe) — one host copy instead of two, and no pageable-memory staging inside the driver.
namespace {
struct JavaSink {
  JNIEnvPtr env;
  jfloatArray array;
  int out_dim;
};
int java_sink(void *user, int first_frame, int n_frames, const float *rows) {
  auto *js = static_cast<JavaSink *>(user);
  auto set = slot<void (*)(JNIEnvPtr, jfloatArray, jsize, jsize, const jfloat *)>(js->env, kJniSetFloatArrayRegion);
  set(js->env, js->array, jsize(first_frame) * js->out_dim, jsize(n_frames) * js->out_dim, rows);
  return 0;
}
}  // namespace

FDNN_JNIEXPORT jfloatArray Java_suskun_nn_QuantizedDnn_calculate(JNIEnvPtr env, jobject, jlong handle, jfloatArray input, jint count, jint dim,
                                                                 jint batch) {
  auto get = slot<jfloat *(*) (JNIEnvPtr, jfloatArray, jboolean *)>(env, kJniGetFloatArrayElements);
  auto rel = slot<void (*)(JNIEnvPtr, jfloatArray, jfloat *, jint)>(env, kJniReleaseFloatArrayElements);
  auto mk = slot<jfloatArray (*)(JNIEnvPtr, jsize)>(env, kJniNewFloatArray);
  auto length = slot<jsize (*)(JNIEnvPtr, jarray)>(env, kJniGetArrayLength);
  (void) batch;  // the reference's CPU cache-blocking batch size
  fdnn_model *model = reinterpret_cast<fdnn_model *>(handle);
  const int O = fdnn_output_dim(model);
  if (count < 0 || dim < 0 || O <= 0) {
    throw_state(env, "bad handle, frame count or dimension");
    return nullptr;
  }
  if (double(length(env, input)) < double(count) * double(dim)) {  // the reference reads count × dim floats whatever the array holds
    throw_state(env, "input array is shorter than count x dimension");
    return nullptr;
  }
  if (double(count) * double(O) > 2147483647.0) {  // a Java array holds at most 2^31 − 1 elements
    throw_state(env, "result does not fit a Java float[]");
    return nullptr;
  }
  jfloatArray result = mk(env, jsize(count) * jsize(O));
  if (!result) return nullptr;  // OutOfMemoryError is already pending
  jfloat *elements = get(env, input, nullptr);
  if (!elements) return nullptr;  // OutOfMemoryError is already pending
  JavaSink sink{env, result, O};
  const int rc = count == 0 ? FDNN_OK : fdnn_calculate_sink(model, elements, count, dim, java_sink, &sink);
  rel(env, input, elements, JNI_ABORT_MODE);
  if (rc != FDNN_OK) {
    throw_state(env, fdnn_last_error());
    return nullptr;
  }
  return result;
}

// jni_dnn.cc:64-77
FDNN_JNIEXPORT jlong Java_suskun_nn_QuantizedDnn_getContext(JNIEnvPtr env, jobject, jlong handle, jint count, jint batch) {
  fdnn_ctx *ctx = nullptr;
  if (fdnn_ctx_new(reinterpret_cast<fdnn_model *>(handle), count, batch, &ctx) != FDNN_OK) {
    throw_state(env, fdnn_last_error());
    return 0;
  }
  return reinterpret_cast<jlong>(ctx);
}

// jni_dnn.cc:79-95
FDNN_JNIEXPORT void Java_suskun_nn_QuantizedDnn_calculateUntilOutput(JNIEnvPtr env, jobject, jlong handle, jfloatArray input) {
  auto get = slot<jfloat *(*) (JNIEnvPtr, jfloatArray, jboolean *)>(env, kJniGetFloatArrayElements);
  auto rel = slot<void (*)(JNIEnvPtr, jfloatArray, jfloat *, jint)>(env, kJniReleaseFloatArrayElements);
  auto length = slot<jsize (*)(JNIEnvPtr, jarray)>(env, kJniGetArrayLength);
  fdnn_ctx *ctx = reinterpret_cast<fdnn_ctx *>(handle);
  // the reference assumes ctx.n × inputDimension floats (jni_dnn.cc:89-91) and reads past a shorter array
  if (double(length(env, input)) < double(fdnn_ctx_frames(ctx)) * double(fdnn_ctx_input_dim(ctx))) {
    throw_state(env, "input array is shorter than the context's frames x input dimension");
    return;
  }
  jfloat *elements = get(env, input, nullptr);
  if (!elements) return;  // OutOfMemoryError is already pending
  int rc = fdnn_ctx_until_output(ctx, elements);
  rel(env, input, elements, JNI_ABORT_MODE);
  if (rc != FDNN_OK) throw_state(env, fdnn_last_error());
}

// jni_dnn.cc:97-117: result length = mask length
FDNN_JNIEXPORT jfloatArray Java_suskun_nn_QuantizedDnn_calculateLazy(JNIEnvPtr env, jobject, jlong handle, jint index, jbyteArray mask) {
  auto get = slot<jbyte *(*) (JNIEnvPtr, jbyteArray, jboolean *)>(env, kJniGetByteArrayElements);
  auto rel = slot<void (*)(JNIEnvPtr, jbyteArray, jbyte *, jint)>(env, kJniReleaseByteArrayElements);
  auto length = slot<jsize (*)(JNIEnvPtr, jarray)>(env, kJniGetArrayLength);
  fdnn_ctx *ctx = reinterpret_cast<fdnn_ctx *>(handle);
  const jsize len = length(env, mask);
  if (len != fdnn_ctx_output_dim(ctx)) {  // the reference trusts the caller here and reads past a short mask
    throw_state(env, "mask length must equal the network's output dimension");
    return nullptr;
  }
  float *out = static_cast<float *>(std::malloc(len > 0 ? size_t(len) * sizeof(float) : 1));
  if (!out) {
    throw_state(env, "out of host memory");
    return nullptr;
  }
  jbyte *bytes = get(env, mask, nullptr);
  if (!bytes) {  // OutOfMemoryError is already pending
    std::free(out);
    return nullptr;
  }
  int rc = fdnn_ctx_lazy(ctx, index, bytes, out);
  rel(env, mask, bytes, JNI_ABORT_MODE);
  jfloatArray result = nullptr;
  if (rc == FDNN_OK)
    result = to_java(env, out, len);
  else
    throw_state(env, fdnn_last_error());
  std::free(out);
  return result;
}

// jni_dnn.cc:119-126
FDNN_JNIEXPORT void Java_suskun_nn_QuantizedDnn_deleteLazyContext(JNIEnvPtr, jobject, jlong handle) {
  fdnn_ctx_free(reinterpret_cast<fdnn_ctx *>(handle));
}

// jni_dnn.cc:128-133
FDNN_JNIEXPORT void Java_suskun_nn_QuantizedDnn_delete(JNIEnvPtr, jobject, jlong handle) { fdnn_free(reinterpret_cast<fdnn_model *>(handle)); }

// jni_dnn.cc:135-148
FDNN_JNIEXPORT jint Java_suskun_nn_QuantizedDnn_layerDimension(JNIEnvPtr, jobject, jlong handle, jint index) {
  return fdnn_layer_dim(reinterpret_cast<fdnn_model *>(handle), index);
}

// jni_dnn.cc:150-156
FDNN_JNIEXPORT jint Java_suskun_nn_QuantizedDnn_layerCount(JNIEnvPtr, jobject, jlong handle) {
  return fdnn_layer_count(reinterpret_cast<fdnn_model *>(handle));
}

}  // extern "C"
