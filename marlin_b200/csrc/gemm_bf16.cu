// bf16 block GEMM on tcgen05 — placeholder translation unit until the TMEM kernel lands.
#include "gemm_bf16.h"

namespace mb {

cudaError_t gemm_bf16(bool, bool, int, int, int, const void*, long long, const void*, long long, void*, long long,
                      bool, bool, int, cudaStream_t, int*) {
    return cudaErrorNotSupported;
}

}  // namespace mb
