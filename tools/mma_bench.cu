// Micro-benchmark: cycles for a burst of tcgen05.mma (M=128, K=16, bf16) as a function of N,
// operand source (A from SMEM vs TMEM) and accumulator reuse.  Garbage data; timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../kaldi_ctc_b200/csrc/tc_common.cuh"
using namespace b200::tc;

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(long long *out, int nmma, int reps) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 0) {
    constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, N);
    const uint32_t a0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0), b0 = a0 + 32768;
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
      const long long t0 = clock64();
      for (int kk = 0; kk < nmma; kk++) {
        const uint64_t ad = smem_desc(a0 + (kk & 3) * 32 + (kk >> 2) * 16384 % 32768, 0, 1024, kLayoutSw128);
        const uint64_t bd = smem_desc(b0 + (kk & 3) * 32, 0, 1024, kLayoutSw128);
        const uint32_t d = tm + (kk % NACC) * N;
        if (elect_one()) {
          if (TS) mma_bf16_ts(d, tm + 256 + (kk % 20) * 8, bd, idesc, kk >= NACC);
          else mma_bf16(d, ad, bd, idesc, kk >= NACC);
        }
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, r & 1);
      tot += clock64() - t0;
    }
    if (threadIdx.x == 0) out[0] = tot / reps;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// NW issuing warps, each with its own accumulator and its own share of the 20 K slices (TS mode, N = 16):
// is the burst bound by the tensor pipe or by the ~13 instructions ptxas emits around every tcgen05.mma?
template <int NW>
__global__ void __launch_bounds__(128, 1) kmw(long long *out, int nmma, int reps) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, NW); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
  if (warp < NW) {
    constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, 16);
    const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + 32768;
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
      const long long t0 = clock64();
      int first = 1;
      for (int kk = warp; kk < nmma; kk += NW) {
        const uint64_t bd = smem_desc(b0 + (kk & 3) * 32, 0, 1024, kLayoutSw128);
        if (elect_one()) mma_bf16_ts(tm + warp * 16, tm + 256 + (kk % 20) * 8, bd, idesc, first ? 0u : 1u);
        first = 0;
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, r & 1);
      tot += clock64() - t0;
    }
    if (threadIdx.x == 0) out[0] = tot / reps;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int NW>
void runmw(long long *d) {
  cudaFuncSetAttribute(kmw<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int nm : {20, 24}) {
    kmw<NW><<<1, 128, 70000>>>(d, nm, 50);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("TS N=16, %d issuing warps, nmma=%2d: cycles=%lld  %s\n", NW, nm, h, cudaGetErrorString(e));
  }
}


// Same burst with every MMA operand provably warp-uniform for ptxas: the issuing warp is a template parameter
// (one code copy per warp), the TMEM base is assumed to be 0 (checked by the caller), the shared-memory address
// comes from the extern array's shared-window offset, and ONE elect per burst guards the MMAs -- so that the
// descriptors live in uniform registers and no R2UR / ELECT / VOTEU sits between two UTCHMMA.
template <int NW, int W>
__device__ __forceinline__ void burst_uniform(uint32_t b0, int nmma, uint64_t *bar) {
  constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, 16);
  const uint64_t bd0 = smem_desc(b0, 0, 1024, kLayoutSw128);
  if (elect_one()) {
#pragma unroll
    for (int i = 0; i < 24 / NW; i++) {
      const int kk = NW * i + W;
      if (kk < nmma) mma_bf16_ts(W * 16, 256 + (kk % 20) * 8, bd0 + (uint64_t)(((kk & 3) * 32) >> 4), idesc, i ? 1u : 0u);
    }
    tc_commit(bar);
  }
  __syncwarp();
}
template <int NW>
__global__ void __launch_bounds__(128, 1) kmu(long long *out, int nmma, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_al[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(smem_al)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, NW); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) out[1] = tm;   // must be 0 for this variant
  const uint32_t b0 = ((smem_u32(smem_al) + 1023u) & ~1023u) + 32768;
  if (warp < NW && tm == 0) {
    long long tot = 0;
    for (int r = 0; r < reps; r++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
      const long long t0 = clock64();
      if (warp == 0) burst_uniform<NW, 0>(b0, nmma, &bar);
      else if (warp == 1) burst_uniform<NW, 1 % NW>(b0, nmma, &bar);
      else if (warp == 2) burst_uniform<NW, 2 % NW>(b0, nmma, &bar);
      else burst_uniform<NW, 3 % NW>(b0, nmma, &bar);
      mbar_wait(&bar, r & 1);
      tot += clock64() - t0;
    }
    if (threadIdx.x == 0) out[0] = tot / reps;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int NW>
void runmu(long long *d) {
  cudaFuncSetAttribute(kmu<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int nm : {20, 24}) {
    kmu<NW><<<1, 128, 70000>>>(d, nm, 50);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("TS N=16 uniform operands, %d issuing warps, nmma=%2d: cycles=%lld (tmem base %lld)  %s\n", NW, nm, h[0], h[1],
           cudaGetErrorString(e));
  }
}

template <int N, bool TS, int NACC>
void run(const char *name, long long *d) {
  cudaFuncSetAttribute(k<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int nm : {1, 20, 40}) {
    k<N, TS, NACC><<<1, 128, 70000>>>(d, nm, 50);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d nmma=%2d  cycles=%lld  (%.1f / mma)  %s\n", name, N, nm, h, (double)h / nm, cudaGetErrorString(e));
  }
}

int main() {
  long long *d;
  cudaMalloc(&d, 64);
  runmw<1>(d);
  runmw<2>(d);
  runmw<4>(d);
  runmu<1>(d);
  runmu<2>(d);
  runmu<4>(d);
  run<16, false, 1>("SS 1acc", d);
  run<16, true, 1>("TS 1acc", d);
  run<16, true, 4>("TS 4acc", d);
  run<32, true, 1>("TS 1acc", d);
  run<64, true, 1>("TS 1acc", d);
  run<128, true, 1>("TS 1acc", d);
  run<128, false, 1>("SS 1acc", d);
  run<256, false, 1>("SS 1acc", d);
  return 0;
}
