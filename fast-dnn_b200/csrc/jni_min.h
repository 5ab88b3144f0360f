// Minimal, self-contained declaration of the part of the Java Native Interface that the eleven
// Java_suskun_nn_QuantizedDnn_* entry points need.  The image has no JDK, so instead of <jni.h>
// this header declares the JNIEnv function table as an array of slots and names the slot indices
// fixed by the JNI specification ("JNI Functions", interface function table); they are checked
// against the reference's vendored header (/root/reference/include/linux/jni.h) by
// tests/test_jni_shim.py when that tree is present.  Types follow jni_md.h for Linux x86-64
// (/root/reference/include/linux/jni_md.h:33-49): jint = int, jlong = long, jbyte = signed char.
#pragma once

#include <cstdint>

extern "C" {

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef float jfloat;
typedef jint jsize;
typedef uint8_t jboolean;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jfloatArray;
typedef jarray jbyteArray;

struct JNIEnv_min;
typedef const void *const *JNIEnvTable;  // *env → function table
typedef JNIEnvTable *JNIEnvPtr;          // JNIEnv*

enum JniSlot {
  kJniFindClass = 6,
  kJniThrowNew = 14,
  kJniGetStringUTFChars = 169,
  kJniReleaseStringUTFChars = 170,
  kJniGetArrayLength = 171,
  kJniNewFloatArray = 181,
  kJniGetByteArrayElements = 184,
  kJniGetFloatArrayElements = 189,
  kJniReleaseByteArrayElements = 192,
  kJniReleaseFloatArrayElements = 197,
  kJniSetFloatArrayRegion = 213,
  kJniTableSize = 233,
};

enum { JNI_ABORT_MODE = 2 };  // JNI_ABORT: free the buffer without copying back

#define FDNN_JNIEXPORT __attribute__((visibility("default")))
}
