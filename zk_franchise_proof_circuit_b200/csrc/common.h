// Shared host-side declarations of libzkcensus_b200 (status codes mirror include/zkcensus_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <string>

#define ZKB_OK 0
#define ZKB_ERROR 1
#define ZKB_SHORT_BUFFER 2
#define ZKB_INVALID_WITNESS_LENGTH 3
#define ZKB_ASSERT_FAILED 4
#define ZKB_UNSUPPORTED_CIRCUIT 5
#define ZKB_INVALID_PROOF 6

extern "C" int zkb_device_count(void);
extern "C" const char *zkb_last_error(void);

namespace zkb {
void set_error(const std::string &s);
int cuda_fail(cudaError_t e, const char *what);
int require_device();
}  // namespace zkb
