// rlsb_count.cuh — counts kernel launches issued by this library (bench.py's `gpu_launches`).
#pragma once
#include <atomic>
namespace rlsb {
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace rlsb
