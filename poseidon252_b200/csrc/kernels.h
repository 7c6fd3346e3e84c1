// Host-callable launchers of the sm_100a kernels (kernels.cu).  Internal to the library; the public
// boundary is include/poseidon252_b200.h.  All pointers are DEVICE pointers, 16-byte aligned;
// scalars are BlsScalar.0 (4 x u64 LE limbs, Montgomery form, < p).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace p252 {

cudaError_t launch_permute(void* states, size_t n, bool dense, cudaStream_t st);
cudaError_t launch_digest(const uint64_t tag[4], const void* in, size_t n, uint32_t in_len, void* out,
                          uint32_t out_len, bool truncate, cudaStream_t st);
cudaError_t launch_convert(const void* in, size_t n, void* out, uint8_t* ok, bool from_bytes, cudaStream_t st);
cudaError_t launch_encrypt(const uint64_t tag[4], const void* msg, size_t n, uint32_t L, const void* secret_uv,
                           const void* nonce, void* cipher, cudaStream_t st);
cudaError_t launch_decrypt(const uint64_t tag[4], const void* cipher, size_t n, uint32_t L, const void* secret_uv,
                           const void* nonce, void* msg, uint8_t* ok, cudaStream_t st);

}  // namespace p252
