// Micro-benchmark: one-way latency of cluster signalling mechanisms on sm_100a.
// Two CTAs of a cluster ping-pong R times; reported = round trip / 2 (cycles).
//   A) st.async (16 B) with mbarrier::complete_tx::bytes on the peer
//   B) st.shared::cluster (16 B) + mbarrier.arrive.release.cluster on the peer, waiter acquire.cluster
//   C) as B but the sender has a global store in flight before the release (what a fused epilogue does)
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include "../kaldi_ctc_b200/csrc/tc_common.cuh"
namespace cg = cooperative_groups;
using namespace b200::tc;

__device__ __forceinline__ void wait_cl(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) k(long long *out, float *g, int reps) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  __shared__ __align__(16) uint32_t buf[4];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    if (MODE == 0) mbar_expect_tx(&bar, 16);
  }
  cl.sync();
  const uint32_t rbuf = mapa_u32(smem_u32(buf), rank ^ 1), rbar = mapa_u32(smem_u32(&bar), rank ^ 1);
  long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (rank == 0) {  // send, then wait for the reply
        if (MODE == 2) g[r & 63] = (float)r;
        if (MODE == 0) st_async_v4(rbuf, r, r, r, r, rbar);
        else {
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(rbuf), "r"(r) : "memory");
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        }
        if (MODE == 0) { mbar_wait(&bar, r & 1); mbar_expect_tx(&bar, 16); } else wait_cl(&bar, r & 1);
      } else {
        if (MODE == 0) { mbar_wait(&bar, r & 1); mbar_expect_tx(&bar, 16); } else wait_cl(&bar, r & 1);
        if (MODE == 2) g[64 + (r & 63)] = (float)r;
        if (MODE == 0) st_async_v4(rbuf, r, r, r, r, rbar);
        else {
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(rbuf), "r"(r) : "memory");
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        }
      }
    }
    if (rank == 0) out[0] = (clock64() - t0) / (2 * reps);
  }
  cl.sync();
}

int main() {
  long long *d; float *g;
  cudaMalloc(&d, 64); cudaMalloc(&g, 1024);
  const char *names[3] = {"st.async + complete_tx", "st.shared::cluster + arrive.release.cluster", "same, with a global store in flight"};
  for (int m = 0; m < 3; m++) {
    if (m == 0) k<0><<<2, 32>>>(d, g, 2000);
    if (m == 1) k<1><<<2, 32>>>(d, g, 2000);
    if (m == 2) k<2><<<2, 32>>>(d, g, 2000);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-50s one-way %lld cycles  (%s)\n", names[m], h, cudaGetErrorString(e));
  }
  return 0;
}
