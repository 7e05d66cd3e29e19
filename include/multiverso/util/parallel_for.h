// ParallelFor: data-parallel loops for the host runtime's bulk operations (row gather / scatter /
// per-row updater application, whole-shard updates).
//
// The reference uses `#pragma omp parallel for` inside the default updater (updater.cpp:22-29).
// In this runtime the same loops run on THREE different threads of one process (the caller,
// the worker actor, the server actor); with OpenMP each of them owns a thread team, and idle
// teams busy-wait for a while after every region (libgomp's default wait policy), stealing the
// cores from whichever actor works next: a 20 000-row Add took 75 ms instead of 7 ms. A single
// shared pool whose idle threads block on a condition variable has no such interference.
#ifndef MULTIVERSO_UTIL_PARALLEL_FOR_H_
#define MULTIVERSO_UTIL_PARALLEL_FOR_H_
#include <cstddef>
#include <cstdint>
#include <functional>

namespace multiverso {

// Runs body(begin, end) over a partition of [0, n) on up to `threads` threads (the caller is one
// of them) and returns when all chunks are done. threads <= 1 or a small n: runs inline.
// Safe to call concurrently from several threads; must not be called from inside a body.
void ParallelFor(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& body);

// memcpy that splits copies of 8 MiB and more over up to `-omp_threads` pool threads (whole-table
// Get / Add payloads are hundreds of MB; one core cannot saturate the memory system).
void ParallelMemcpy(void* dst, const void* src, size_t bytes);

// Number of pool threads (hardware concurrency, at least 1).
int ParallelForCapacity();

}  // namespace multiverso
#endif
