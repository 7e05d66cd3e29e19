// Tooling, not product code: the reference's own matrix perf test (Test/test_matrix_perf.cpp) is
// not wired into its Test/main.cpp; this 10-line driver calls it unmodified so that
// tools/build_reference.sh can build baseline/_ref/bin/matrix_perf for a CPU-vs-CPU comparison
// with `mv_test dense_perf|sparse_perf` (bench/cpu_matrix_perf.py).
#include <cstring>
namespace multiverso { namespace test {
void TestDensePerf(int argc, char* argv[]);
void TestSparsePerf(int argc, char* argv[]);
} }
int main(int argc, char* argv[]) {
  if (argc >= 2 && std::strcmp(argv[1], "sparse") == 0) multiverso::test::TestSparsePerf(argc, argv);
  else multiverso::test::TestDensePerf(argc, argv);
  return 0;
}
