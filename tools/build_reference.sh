#!/bin/bash
# Build the UNMODIFIED reference (Microsoft/multiverso core + Applications/WordEmbedding) against
# the single-node MPI shim in baseline/mpi_shim -> baseline/_ref/bin/wordembedding.
# /root/reference is read-only, so sources are compiled in place into a scratch object dir.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REFERENCE_DIR:-/root/reference}
OUT=$ROOT/baseline/_ref
OBJ=/tmp/mv_ref_build_$$
mkdir -p $OUT/bin $OBJ
CXXFLAGS="-O3 -std=c++11 -fPIC -fopenmp -w -include cstddef -include cstdint -include cstdlib -include cstring -include string -include functional -DMULTIVERSO_USE_MPI -I$ROOT/baseline/mpi_shim -I$REF/include"
CORE="actor.cpp communicator.cpp controller.cpp dashboard.cpp multiverso.cpp net.cpp node.cpp server.cpp table.cpp table/array_table.cpp table/matrix_table.cpp table/sparse_matrix_table.cpp table/matrix.cpp timer.cpp updater/updater.cpp util/configure.cpp io/io.cpp io/local_stream.cpp util/log.cpp util/net_util.cpp worker.cpp zoo.cpp c_api.cpp util/allocator.cpp table_factory.cpp blob.cpp"
pids=()
for f in $CORE; do
  o=$OBJ/core_$(echo $f | tr '/' '_').o
  g++ $CXXFLAGS -c $REF/src/$f -o $o &
  pids+=($!)
done
for f in $REF/Applications/WordEmbedding/src/*.cpp; do
  o=$OBJ/we_$(basename $f).o
  g++ $CXXFLAGS -Wno-sign-compare -I$REF/Applications/WordEmbedding/src -c $f -o $o &
  pids+=($!)
done
g++ -O2 -std=c++11 -fPIC -I$ROOT/baseline/mpi_shim -c $ROOT/baseline/mpi_shim/mpi_shim.cpp -o $OBJ/mpi_shim.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -fopenmp -o $OUT/bin/wordembedding $OBJ/*.o -lpthread -ldl
echo "built $OUT/bin/wordembedding"
# the reference's matrix perf test (Test/test_matrix_perf.cpp, unmodified) with a tiny driver
g++ $CXXFLAGS -c $REF/Test/test_matrix_perf.cpp -o $OBJ/perf_test.o &
g++ -O2 -std=c++11 -c $ROOT/baseline/ref_matrix_perf_main.cpp -o $OBJ/perf_main.o &
wait
g++ -fopenmp -o $OUT/bin/matrix_perf $OBJ/core_*.o $OBJ/mpi_shim.o $OBJ/perf_test.o $OBJ/perf_main.o -lpthread -ldl \
  && echo "built $OUT/bin/matrix_perf"
# BASELINE config 2 through the reference's public API: MatrixTable rows x cols whole-table Add / Get
g++ $CXXFLAGS -c $ROOT/baseline/ref_matrix_bw_main.cpp -o $OBJ/bw_main.o \
  && g++ -fopenmp -o $OUT/bin/matrix_bw $OBJ/core_*.o $OBJ/mpi_shim.o $OBJ/bw_main.o -lpthread -ldl \
  && echo "built $OUT/bin/matrix_bw"
rm -rf $OBJ
