#!/bin/bash
# kernel-variant libraries for measurement sweeps (anyseq_b200/_build_<tag>/, git-ignored; selected with ANYSEQ_LIB=<path to the .so>)
#   prof   -DANYSEQ_PROFILE: in-kernel cycle counters per batch, time of the strip kernel alone, per-strip timeline (ANYSEQ_TRACE_FILE)
#   r44    four-row K = 4 tiles and two-row K = 8 tiles (the shipped build has them the other way round)
#   f1r1   single-row K = 16/32 tiles, two-row K = 8
# Remove the directories before the end of a round: they travel to the GPU box with every gpurun snapshot (40 MB each).
cd "$(dirname "$0")/.."
build() { tag=$1; shift; ANYSEQ_BUILD_DIR=$PWD/anyseq_b200/_build_$tag ANYSEQ_NVCC_FLAGS="$*" python -m anyseq_b200.build -v 2>&1 | grep -E "error|bytes stack frame, [1-9]" | head -3; }
for tag in "${@:-prof}"; do
    case $tag in
        prof) build prof -DANYSEQ_PROFILE ;;
        r44)  build r44 -DANYSEQ_ROWS_K4=4 -DANYSEQ_ROWS_K8=2 ;;
        f1r1) build f1r1 -DANYSEQ_ROWS_K32=1 -DANYSEQ_ROWS_K16=1 -DANYSEQ_ROWS_K8=2 ;;
        *) echo "unknown variant $tag" ;;
    esac
done
python -m anyseq_b200.build > /dev/null 2>&1 || echo "main build FAILED"
ls -la anyseq_b200/_build*/libanyseq_b200.so
