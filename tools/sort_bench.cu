// Sort comparator (the reference's own bench compares its Vulkan sort with CUB: third_party/vulkan_radix_sort/bench/
// cuda_benchmark.cu:83-99): times cub::DeviceRadixSort::SortPairs and libvkgsb's vkgsb_sort_key_value_indirect on the
// same device and the same keys.  Not part of the product; build: tools/build_sort_bench.sh
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../include/vkgsb.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

int main(int argc, char** argv) {
  std::vector<uint32_t> sizes = {1u << 20, 2000000u, 4190000u, 1u << 24, 1u << 25};
  if (argc > 1) { sizes.clear(); for (int i = 1; i < argc; ++i) sizes.push_back(static_cast<uint32_t>(std::atoll(argv[i]))); }
  for (int dist = 0; dist < 2; ++dist)
    for (uint32_t n : sizes) {
      std::vector<uint32_t> hk(n), hv(n);
      std::mt19937 rng(1234 + n);
      for (uint32_t i = 0; i < n; ++i) {
        if (dist == 0) hk[i] = rng();  // uniform u32 (data_generator.cc:12-27)
        else {                         // depth-like: bits(1 - z), z = far-biased NDC depth (rank.comp:39)
          float d = 0.3f + 40.f * static_cast<float>(rng() & 0xffffff) / 16777216.f;
          float z = 1.0001f * (1.f - 0.01f / d);
          float k = 1.f - z; std::memcpy(&hk[i], &k, 4);
        }
        hv[i] = i;
      }
      uint32_t *dk, *dv, *dk2, *dv2, *dcount; void* tmp = nullptr; size_t tmp_bytes = 0;
      CK(cudaMalloc(&dk, n * 4ull)); CK(cudaMalloc(&dv, n * 4ull)); CK(cudaMalloc(&dk2, n * 4ull)); CK(cudaMalloc(&dv2, n * 4ull));
      CK(cudaMalloc(&dcount, 4)); CK(cudaMemcpy(dcount, &n, 4, cudaMemcpyHostToDevice));
      cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dk2, dv, dv2, static_cast<int>(n));
      CK(cudaMalloc(&tmp, tmp_bytes));
      size_t vbytes = 0; vkgsb_sort_storage_bytes(n, &vbytes);
      void* vstore; CK(cudaMalloc(&vstore, vbytes));
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      cudaStream_t s; CK(cudaStreamCreate(&s));
      float best_cub = 1e9f, best_ours = 1e9f;
      for (int it = 0; it < 12; ++it) {
        CK(cudaMemcpyAsync(dk, hk.data(), n * 4ull, cudaMemcpyHostToDevice, s)); CK(cudaMemcpyAsync(dv, hv.data(), n * 4ull, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(e0, s));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dk2, dv, dv2, static_cast<int>(n), 0, 32, s);
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it >= 2) best_cub = std::min(best_cub, ms);
      }
      std::vector<uint32_t> ref_k(n), ref_v(n), our_k(n), our_v(n);
      CK(cudaMemcpy(ref_k.data(), dk2, n * 4ull, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ref_v.data(), dv2, n * 4ull, cudaMemcpyDeviceToHost));
      for (int it = 0; it < 12; ++it) {
        CK(cudaMemcpyAsync(dk, hk.data(), n * 4ull, cudaMemcpyHostToDevice, s)); CK(cudaMemcpyAsync(dv, hv.data(), n * 4ull, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(e0, s));
        if (vkgsb_sort_key_value_indirect(s, n, dcount, dk, dv, vstore) != 0) { std::printf("vkgsb sort failed: %s\n", vkgsb_last_error()); return 1; }
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it >= 2) best_ours = std::min(best_ours, ms);
      }
      CK(cudaMemcpy(our_k.data(), dk, n * 4ull, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(our_v.data(), dv, n * 4ull, cudaMemcpyDeviceToHost));
      bool same = our_k == ref_k && our_v == ref_v;  // both stable: identical permutations
      std::printf("{\"keys\": \"%s\", \"n\": %u, \"cub_ms\": %.4f, \"cub_gkeys\": %.2f, \"vkgsb_ms\": %.4f, \"vkgsb_gkeys\": %.2f, \"identical\": %s}\n",
                  dist == 0 ? "uniform_u32" : "depth_like", n, best_cub, n / best_cub * 1e-6, best_ours, n / best_ours * 1e-6, same ? "true" : "false");
      cudaFree(dk); cudaFree(dv); cudaFree(dk2); cudaFree(dv2); cudaFree(dcount); cudaFree(tmp); cudaFree(vstore);
      cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
    }
  return 0;
}
