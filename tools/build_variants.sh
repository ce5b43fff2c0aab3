#!/bin/bash
# kernel-variant libraries for measurement sweeps (anyseq_b200/_build_<tag>/, git-ignored; selected with ANYSEQ_LIB)
cd "$(dirname "$0")/.."
build() { tag=$1; shift; ANYSEQ_BUILD_DIR=$PWD/anyseq_b200/_build_$tag ANYSEQ_NVCC_FLAGS="$*" python -m anyseq_b200.build -v 2>&1 | grep -E "error|bytes stack frame, [1-9]" | grep -v "Lb0ELb" | head -3; }
build f0 -DANYSEQ_CELL_FORM=0 &
build f1r1 -DANYSEQ_CELL_FORM=1 -DANYSEQ_ROWS_K32=1 -DANYSEQ_ROWS_K16=1 -DANYSEQ_ROWS_K8=2 &
wait
build f1p -DANYSEQ_CELL_FORM=1 -DANYSEQ_PROFILE &
build f0p -DANYSEQ_CELL_FORM=0 -DANYSEQ_PROFILE &
wait
python -m anyseq_b200.build > /dev/null 2>&1 || echo "main build FAILED"
ls -la anyseq_b200/_build*/libanyseq_b200.so
