/* internal.h -- shared between the translation units of libtroute_b200.so (not part of the C ABI) */
#pragma once
/* records `msg` as trt_last_error() of the calling thread and returns `code` (engine.cu) */
int trt_internal_fail(int code, const char* msg);
