/*
 * jni.h -- a MINIMAL stand-in for the JDK's <jni.h>, for this image only (it has no JDK).
 *
 * It declares exactly the types, macros and JNIEnv members that integration/mhapb_jni.c uses, with the names,
 * argument lists and calling convention of the JNI specification (Java SE, "JNI Functions"), so that the shim
 *   (a) compiles here (tests/test_jni_shim.py: gcc -fsyntax-only and a real shared-object build), and
 *   (b) can be driven on the GPU box by tests/cpp/jni_harness.c through a fake JNIEnv whose function table is filled
 *       with a small C implementation of these members.
 * The member ORDER of the real JNINativeInterface_ is not reproduced: a shim built against this header must not be
 * loaded into a JVM.  A maintainer builds integration/mhapb_jni.c against $JAVA_HOME/include/jni.h instead; nothing in
 * the shim depends on this file beyond the names below.
 */
#ifndef MHAPB_JNI_STUB_H
#define MHAPB_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_STUB_NOT_A_REAL_JDK 1

typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jthrowable;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    void *reserved0;
    jclass (JNICALL *FindClass)(JNIEnv *env, const char *name);
    jint (JNICALL *ThrowNew)(JNIEnv *env, jclass clazz, const char *msg);
    jboolean (JNICALL *ExceptionCheck)(JNIEnv *env);
    jsize (JNICALL *GetArrayLength)(JNIEnv *env, jarray array);
    jbyteArray (JNICALL *NewByteArray)(JNIEnv *env, jsize len);
    jbyte *(JNICALL *GetByteArrayElements)(JNIEnv *env, jbyteArray array, jboolean *isCopy);
    void (JNICALL *ReleaseByteArrayElements)(JNIEnv *env, jbyteArray array, jbyte *elems, jint mode);
    void (JNICALL *SetByteArrayRegion)(JNIEnv *env, jbyteArray array, jsize start, jsize len, const jbyte *buf);
    jint *(JNICALL *GetIntArrayElements)(JNIEnv *env, jintArray array, jboolean *isCopy);
    void (JNICALL *ReleaseIntArrayElements)(JNIEnv *env, jintArray array, jint *elems, jint mode);
    jlong *(JNICALL *GetLongArrayElements)(JNIEnv *env, jlongArray array, jboolean *isCopy);
    void (JNICALL *ReleaseLongArrayElements)(JNIEnv *env, jlongArray array, jlong *elems, jint mode);
    void (JNICALL *SetLongArrayRegion)(JNIEnv *env, jlongArray array, jsize start, jsize len, const jlong *buf);
    jobject (JNICALL *NewDirectByteBuffer)(JNIEnv *env, void *address, jlong capacity);
    void *(JNICALL *GetDirectBufferAddress)(JNIEnv *env, jobject buf);
    jlong (JNICALL *GetDirectBufferCapacity)(JNIEnv *env, jobject buf);
};

#endif
