// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <stdarg.h>
#include <stdio.h>

namespace mdb {
char* error_buffer();  // thread-local, 512 bytes (defined in capi_misc.cu)
inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace mdb

#define MDB_CHECK_LAUNCH(name)                                                                   \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) return mdb::set_error(MDB_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)
