#!/bin/bash
# Build the product library (hand-written sm_100a CUDA behind the C ABI) in-tree.
#   fastqueeze_b200/libseqarc_b200.so
set -e
cd "$(dirname "$0")/.."
SRC=fastqueeze_b200/csrc
OUT=fastqueeze_b200/libseqarc_b200.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${SQ_NVCC_EXTRA}"
mkdir -p build
for f in sq_encode sq_decode sq_api; do
  if [ ! -f build/$f.o ] || [ -n "$(find $SRC include -newer build/$f.o -type f | head -1)" ]; then
    $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o &
  fi
done
wait
$NVCC -shared -o $OUT build/sq_encode.o build/sq_decode.o build/sq_api.o
echo "built $OUT"
# host driver (C++): FASTQ block cutter, .arc packer, CLI -> fastqueeze_b200/seqarc_b200
g++ -O2 -std=c++17 -Wall -o fastqueeze_b200/seqarc_b200 host/seqarc_b200_main.cpp host/fastq_block.cpp host/arc_file.cpp host/id_process.cpp \
    -Lfastqueeze_b200 -lseqarc_b200 -lz -pthread -Wl,-rpath,'$ORIGIN'
echo "built fastqueeze_b200/seqarc_b200"
g++ -O2 -std=c++17 -Wall -shared -fPIC -o fastqueeze_b200/libseqarc_host.so host/host_capi.cpp host/fastq_block.cpp host/arc_file.cpp host/id_process.cpp -lz
echo "built fastqueeze_b200/libseqarc_host.so"
