// Shared plumbing of libsosfront.so: the context object, error reporting and launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "sosfront.h"

struct sos_prof_mark {
  const char* name;  // __func__ of the C-ABI entry point that launched the kernel (static storage)
  cudaEvent_t ev;    // recorded right after the launch
};

struct sos_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  void* arena = nullptr;  // grow-only scratch
  size_t arena_bytes = 0;
  int64_t launches = 0;
  int sm_count = 148;
  bool prof = false;      // per-launch CUDA-event timing (sos_ctx_profile_begin / _end)
  std::vector<sos_prof_mark> marks;
};

void sos_set_error(const char* fmt, ...);
void sos_prof_mark_launch(sos_ctx* ctx, const char* name);

#define SOS_CHECK_ARG(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      sos_set_error("%s: invalid argument: %s", __func__, msg);   \
      return SOS_ERR_INVALID;                                     \
    }                                                             \
  } while (0)

#define SOS_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      sos_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));        \
      return SOS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.  The profiling mark carries the
// kernel's name (SOS_LAUNCHED_AS) or, by default, the entry point that launched it.
#define SOS_LAUNCHED(ctx) SOS_LAUNCHED_AS(ctx, __func__)
#define SOS_LAUNCHED_AS(ctx, mark_name)                                                    \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    if ((ctx)->prof) sos_prof_mark_launch((ctx), (mark_name));                             \
    cudaError_t e__ = cudaPeekAtLastError();                                               \
    if (e__ != cudaSuccess) {                                                              \
      (void)cudaGetLastError();                                                            \
      sos_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__));    \
      return SOS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// Scratch arena: returns a pointer to at least `bytes` of device memory (256-byte aligned).
int sos_arena_get(sos_ctx* ctx, size_t bytes, void** out);

static inline int sos_div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t sos_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
