// swb_gen.cu -- the portable counter-based sequence generator ON THE DEVICE.
//
// Bit-identical to concurrentproject_b200/rng.py (numpy) and oracle/gotoh_oracle.c (oracle_mix64 /
// oracle_random_acgt): symbol k of stream s = 2 bits of mix64(seed, s, k / 32); the planted-similarity recipes of
// BASELINE configs 4 and 5 (rng.mutate: i.i.d. substitutions and single-base insertions / deletions decided by
// mix64(seed, stream, position)) are restated here so that a 10 M-pair batch is born in HBM instead of crossing
// PCIe, while any single pair can be regenerated on the host for the oracle.  Replaces the reference harness's
// unseeded rand() % 4 (TestFileWithGPU.cpp:25-36).
#include "../../include/swb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t stream, uint64_t index) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1) + 0xD1B54A32D192ED03ull * stream;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint8_t nt(uint32_t code) { return (uint8_t)((0x54474341u >> (8 * (code & 3))) & 0xFFu); }   // "ACGT"

__device__ __forceinline__ uint8_t random_symbol(uint64_t seed, uint64_t stream, long long k) {
  return nt((uint32_t)(mix64(seed, stream, (uint64_t)(k >> 5)) >> (2 * (k & 31))));
}

__global__ void gen_random_kernel(uint64_t seed, uint64_t stream, long long length, uint8_t* __restrict__ out) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w * 32 < length; w += (long long)gridDim.x * blockDim.x) {
    const uint64_t bits = mix64(seed, stream, (uint64_t)w);
    for (int k = 0; k < 32 && w * 32 + k < length; ++k) out[w * 32 + k] = nt((uint32_t)(bits >> (2 * k)));
  }
}

struct Thresholds { uint64_t sub, del, ins; };     // on (r >> 11): u < rate  <=>  (r >> 11) < ceil(rate * 2^53)

// rng.mutate(src = random stream `src_stream` from `src_off`, decisions from stream `dec_stream`), first out_len
// symbols of the result; pads with 'A' if the mutated source is shorter (never happens with the margins used).
__device__ void mutate_into(uint64_t seed, uint64_t src_stream, long long src_off, int src_len, uint64_t dec_stream,
                            Thresholds th, uint8_t* __restrict__ out, int out_len) {
  int o = 0;
  for (int i = 0; i < src_len && o < out_len; ++i) {
    const uint8_t c = random_symbol(seed, src_stream, src_off + i);
    const uint64_t r = mix64(seed, dec_stream, (uint64_t)i);
    const uint64_t u = r >> 11;
    const uint32_t newc = (uint32_t)(r >> 3) & 3u;
    const int kind = u < th.sub ? 1 : (u < th.del ? 2 : (u < th.ins ? 3 : 0));    // substitute / delete / insert after
    if (kind != 2) {
      uint8_t b = c;
      if (kind == 1) { b = nt(newc); if (b == c) b = nt(newc + 1); }
      out[o++] = b;
    }
    if (kind == 3 && o < out_len) out[o++] = nt(newc + 2);
  }
  while (o < out_len) out[o++] = 'A';
}

// BASELINE config 4: pair k (global id) = window of wl random bases (stream 2k); even k: the read is a substring of the
// window (offset = mix64(seed, 2k+1, 2^40) % (wl - rl - 15)) with 5 % substitutions and 1 % indels, odd k: rl random
// bases (stream 2k+1).
__global__ void gen_read_pairs_kernel(uint64_t seed, long long first_pair, long long npairs, int rl, int wl, Thresholds th,
                                      uint8_t* __restrict__ reads, uint8_t* __restrict__ windows) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    const long long k = first_pair + p;
    uint8_t* w = windows + p * wl;
    for (int i = 0; i < wl; ++i) w[i] = random_symbol(seed, 2 * (uint64_t)k, i);
    uint8_t* r = reads + p * rl;
    if ((k & 1) == 0) {
      const long long off = (long long)(mix64(seed, 2 * (uint64_t)k + 1, 1ull << 40) % (uint64_t)(wl - rl - 15));
      mutate_into(seed, 2 * (uint64_t)k, off, rl + 16, 2 * (uint64_t)k + 1, th, r, rl);
    } else {
      for (int i = 0; i < rl; ++i) r[i] = random_symbol(seed, 2 * (uint64_t)k + 1, i);
    }
  }
}

// BASELINE config 5: pair k = len random bases (stream 2k) and the same stretch (plus 320 spare bases) with 10 %
// substitutions and 2 % short indels (decisions: stream 2k+1), cut to len: the optimum stays near the main diagonal.
__global__ void gen_long_pairs_kernel(uint64_t seed, long long first_pair, long long npairs, int len, Thresholds th,
                                      uint8_t* __restrict__ seq1, uint8_t* __restrict__ seq2) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
    const long long k = first_pair + p;
    uint8_t* a = seq1 + p * len;
    for (int i = 0; i < len; ++i) a[i] = random_symbol(seed, 2 * (uint64_t)k, i);
    mutate_into(seed, 2 * (uint64_t)k, 0, len + 320, 2 * (uint64_t)k + 1, th, seq2 + p * len, len);
  }
}

Thresholds thresholds(double sub_rate, double indel_rate) {
  // same float64 arithmetic as rng.mutate: kinds by u < sub_hi, u < del_hi, u < ins_hi with u = (r >> 11) / 2^53
  const double sub_hi = sub_rate, del_hi = sub_hi + indel_rate / 2.0, ins_hi = del_hi + indel_rate / 2.0;
  auto t = [](double rate) -> uint64_t {
    if (rate <= 0) return 0;
    if (rate >= 1) return 1ull << 53;
    const double x = rate * 9007199254740992.0;          // exact: a power of two
    const uint64_t f = (uint64_t)x;
    return (double)f < x ? f + 1 : f;                     // ceil
  };
  return Thresholds{t(sub_hi), t(del_hi), t(ins_hi)};
}

}  // namespace

extern "C" {

int swb200_gen_random_device(int device, unsigned long long seed, unsigned long long stream_id, long long length,
                             unsigned char* d_out, void* stream) {
  if (length < 0 || (length > 0 && !d_out)) return SWB200_ERR_ARG;
  if (length == 0) return SWB200_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return SWB200_ERR_CUDA;
  const long long words = (length + 31) / 32;
  const int blocks = (int)((words + 255) / 256 < 4096 ? (words + 255) / 256 : 4096);
  gen_random_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, length, d_out);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  return e == cudaSuccess ? SWB200_OK : SWB200_ERR_CUDA;
}

int swb200_gen_read_pairs_device(int device, unsigned long long seed, long long first_pair, long long npairs, int read_len,
                                 int window_len, unsigned char* d_reads, unsigned char* d_windows, void* stream) {
  if (npairs < 0 || first_pair < 0 || read_len < 1 || window_len < read_len + 16 || (npairs > 0 && (!d_reads || !d_windows)))
    return SWB200_ERR_ARG;
  if (npairs == 0) return SWB200_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return SWB200_ERR_CUDA;
  const int blocks = (int)((npairs + 127) / 128 < 8192 ? (npairs + 127) / 128 : 8192);
  gen_read_pairs_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(seed, first_pair, npairs, read_len, window_len,
                                                                  thresholds(0.05, 0.01), d_reads, d_windows);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  return e == cudaSuccess ? SWB200_OK : SWB200_ERR_CUDA;
}

int swb200_gen_long_pairs_device(int device, unsigned long long seed, long long first_pair, long long npairs, int len,
                                 unsigned char* d_seq1, unsigned char* d_seq2, void* stream) {
  if (npairs < 0 || first_pair < 0 || len < 1 || (npairs > 0 && (!d_seq1 || !d_seq2))) return SWB200_ERR_ARG;
  if (npairs == 0) return SWB200_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return SWB200_ERR_CUDA;
  const int blocks = (int)((npairs + 63) / 64 < 8192 ? (npairs + 63) / 64 : 8192);
  gen_long_pairs_kernel<<<blocks, 64, 0, (cudaStream_t)stream>>>(seed, first_pair, npairs, len, thresholds(0.10, 0.02), d_seq1, d_seq2);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  return e == cudaSuccess ? SWB200_OK : SWB200_ERR_CUDA;
}

}  // extern "C"
