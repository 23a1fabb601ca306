// Micro-benchmark behind the window scorer's design (profiles/README.md): cost of shared-memory atomics vs plain
// shared-memory accesses vs warp match on one SM with 32 resident warps (4 CTAs x 8 warps, like k3_score_*).
// Prints SM cycles per warp-level instruction (all 32 lanes active) for each access pattern.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/smem_atomics tools/microbench/smem_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096, WARPS = 8, SLOTS = 1024;

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int OP, int PATTERN>
__global__ void __launch_bounds__(WARPS * 32) bench(uint32_t* out) {
  __shared__ uint32_t tab[WARPS][SLOTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* t = tab[warp];
  for (int i = lane; i < SLOTS; i += 32) t[i] = 0;
  __syncwarp();
  uint32_t s = blockIdx.x * 977 + threadIdx.x * 31 + 1, acc = 0;
  for (int it = 0; it < ITER; ++it) {
    uint32_t r = lcg(s), a;
    if (PATTERN == 0) a = r & (SLOTS - 1);                                // random slots (random bank conflicts)
    else if (PATTERN == 1) a = (lane + 32 * (r & 31)) & (SLOTS - 1);      // conflict-free: bank = lane
    else if (PATTERN == 2) a = ((lane >> 3) + 4 * (r & 7)) & (SLOTS - 1); // 8 lanes share one address
    else a = (r % 24 < 12) ? (r & 7) : (r & (SLOTS - 1));                 // half the lanes in 8 hot slots (1D-like)
    if (OP == 0) {
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(t + a)), "r"((r & 256u) ? 0x10000u : 1u) : "memory");
    } else if (OP == 1) {
      acc += atomicAdd(t + a, 1u);
    } else if (OP == 2) {
      acc += atomicCAS(t + a, it & 1 ? 0u : 7u, 7u);
    } else if (OP == 3) {
      uint32_t v = t[a];
      __syncwarp();
      t[a] = v + 1;
      acc += v;
    } else if (OP == 4) {
      acc += __match_any_sync(0xffffffffu, a);
    } else if (OP == 5) {
      acc += t[a];
    }
  }
  if (acc == 0xFFFFFFFFu) out[0] = acc;
}

template <int OP, int PATTERN>
static void run(const char* name, uint32_t* d) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<OP, PATTERN><<<sms * 4, WARPS * 32>>>(d);
  cudaEventRecord(e0);
  for (int k = 0; k < 5; ++k) bench<OP, PATTERN><<<sms * 4, WARPS * 32>>>(d);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms / 5 * 1e-3 * khz * 1e3;           // SM cycles per launch (at the max clock)
  printf("%-34s %8.2f cycles per warp instruction per SM (32 warps resident)\n", name, cycles / (double(ITER) * 32));
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 64);
  const char* pat[4] = {"random", "conflict-free", "8 lanes/address", "half in 8 hot slots"};
#define ROW(OP, NAME)                                                                      \
  { char b[96];                                                                            \
    snprintf(b, sizeof b, "%s, %s", NAME, pat[0]); run<OP, 0>(b, d);                        \
    snprintf(b, sizeof b, "%s, %s", NAME, pat[1]); run<OP, 1>(b, d);                        \
    snprintf(b, sizeof b, "%s, %s", NAME, pat[2]); run<OP, 2>(b, d);                        \
    snprintf(b, sizeof b, "%s, %s", NAME, pat[3]); run<OP, 3>(b, d); }
  ROW(5, "LDS")
  ROW(3, "LDS+syncwarp+STS")
  ROW(0, "RED.add (no return)")
  ROW(1, "ATOMS.add (return)")
  ROW(2, "ATOMS.cas")
  ROW(4, "MATCH.ANY")
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
