// Micro-benchmark behind the ranking choice of k_radix_scatter / k_radix_agg: cycles per tuple per SM of
//   A  shared atomicAdd WITH the returned rank (random bins)         B  the same, result unused
//   C  ballot-match ranking (nbits ballots) + warp-private counters   D  __match_any_sync + warp-private counters
//   E  random 8-byte LDS / STS
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rank_ubench rank_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
constexpr int ITERS = 64, U = 8;
template <int MODE>
__global__ void k(unsigned* out, int nbins, int nbits, long long* cyc) {
  extern __shared__ unsigned sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16384; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  unsigned acc = 0;
  uint64_t seed = (uint64_t)blockIdx.x * 1315423911u + tid;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    unsigned d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) d[u] = (unsigned)(fmix64(seed + (uint64_t)(it * U + u) * 0x9e3779b97f4a7c15ULL) >> 40) & (unsigned)(nbins - 1);
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc ^= atomicAdd(&sm[d[u]], 1u);
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) atomicAdd(&sm[d[u]], 1u);
    } else if (MODE == 2) {
      unsigned short* wh = (unsigned short*)sm + warp * nbins;  // warp-private 16-bit counters
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned peers = 0xffffffffu;
        for (int b = 0; b < nbits; ++b) {
          const unsigned bit = (d[u] >> b) & 1u;
          const unsigned m = __ballot_sync(0xffffffffu, bit);
          peers &= bit ? m : ~m;
        }
        const int leader = __ffs(peers) - 1;
        const unsigned rank_in = __popc(peers & ((1u << lane) - 1u));
        unsigned old = 0;
        if (lane == leader) { old = wh[d[u]]; wh[d[u]] = (unsigned short)(old + __popc(peers)); }
        old = __shfl_sync(0xffffffffu, old, leader);
        acc ^= old + rank_in;
        __syncwarp();
      }
    } else if (MODE == 3) {
      unsigned short* wh = (unsigned short*)sm + warp * nbins;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned peers = __match_any_sync(0xffffffffu, d[u]);
        const int leader = __ffs(peers) - 1;
        const unsigned rank_in = __popc(peers & ((1u << lane) - 1u));
        unsigned old = 0;
        if (lane == leader) { old = wh[d[u]]; wh[d[u]] = (unsigned short)(old + __popc(peers)); }
        old = __shfl_sync(0xffffffffu, old, leader);
        acc ^= old + rank_in;
        __syncwarp();
      }
    } else if (MODE == 4) {
      const unsigned long long* s8 = (const unsigned long long*)sm;
#pragma unroll
      for (int u = 0; u < U; ++u) acc ^= (unsigned)s8[d[u]];
    } else if (MODE == 5) {
      unsigned long long* s8 = (unsigned long long*)sm;
#pragma unroll
      for (int u = 0; u < U; ++u) s8[d[u]] = seed + u;
    } else if (MODE == 6) {  // hash only (baseline to subtract)
#pragma unroll
      for (int u = 0; u < U; ++u) acc ^= d[u];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + tid] = acc + sm[tid];
}
template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, int nbins, int nbits) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm;
  unsigned* out; long long* cyc;
  cudaMalloc(&out, (size_t)grid * threads * 4);
  cudaMalloc(&cyc, grid * 8);
  const size_t smem = 65536;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<grid, threads, smem>>>(out, nbins, nbits, cyc);
  k<MODE><<<grid, threads, smem>>>(out, nbins, nbits, cyc);
  cudaDeviceSynchronize();
  long long h[2048];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)h[i];
  avg /= grid;
  const double tuples_per_sm = (double)threads * ctas_per_sm * ITERS * U;
  printf("%-28s thr=%4d x%d bins=%5d : %7.3f cyc/tuple/SM  (%s)\n", name, threads, ctas_per_sm, nbins, avg / tuples_per_sm,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int thr = cfg ? 1024 : 512, per = cfg ? 1 : 2;
    run<6>("hash only", thr, per, 256, 8);
    for (int nb : {256, 512, 2048}) {
      int bits = 0; while ((1 << bits) < nb) ++bits;
      run<0>("ATOMS with return", thr, per, nb, bits);
      run<1>("ATOMS no return", thr, per, nb, bits);
      if (nb <= 512) {
        run<2>("ballot-match + private ctr", thr, per, nb, bits);
        run<3>("match_any + private ctr", thr, per, nb, bits);
      }
      run<4>("LDS.64 random", thr, per, nb, bits);
      run<5>("STS.64 random", thr, per, nb, bits);
    }
  }
  return 0;
}
