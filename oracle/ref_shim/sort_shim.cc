// C shim over the reference sort's own CPU statement, CpuBenchmark::SortKeyValue
// (third_party/vulkan_radix_sort/bench/cpu_benchmark.cc:29-51 = std::stable_sort by key), compiled from
// /root/reference by oracle/build_ref.py into oracle/_ref/libref_sort.so.  TEST INFRASTRUCTURE ONLY.
#include "cpu_benchmark.h"

#include <cstring>

namespace {
struct Access : CpuBenchmark {
  using CpuBenchmark::Results;
};
}  // namespace

extern "C" __attribute__((visibility("default"))) void ref_sort_key_value(unsigned n, const unsigned* keys,
                                                                           const unsigned* vals, unsigned* out_keys,
                                                                           unsigned* out_vals) {
  std::vector<uint32_t> k(keys, keys + n), v(vals, vals + n);
  CpuBenchmark b;
  Access::Results r = b.SortKeyValue(k, v);
  std::memcpy(out_keys, r.keys.data(), n * 4ull);
  std::memcpy(out_vals, r.values.data(), n * 4ull);
}
