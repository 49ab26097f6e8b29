#!/bin/bash
# Build several variants of the CUDA library HERE (nvcc cross-compiles without a GPU) and time
# them all in ONE gpurun call -- GPU minutes are the scarce resource, compile time is not.
#
#   tools/variants.sh build base="" mb7="-DTB_MINI_MIN_BLOCKS=7" ...   # -> _variants/<name>.so
#   gpurun --timeout 900 -- 'bash tools/variants.sh run 2>&1 | tee gpurun_out/variants.txt'
#   tools/variants.sh clean
#
# `run` swaps each prebuilt library in, runs tools/kernel_variants.py on the C2 shape for k = 21
# and 31 (pass rate, step and kernel time, and the count checksums, which must not change from
# variant to variant), and puts the original library back.  _variants/ is git-ignored but travels
# with the gpurun snapshot.
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LIB=$(ls "$ROOT"/trio_binning_b200/kmers_c*.so 2>/dev/null | head -1)
case "${1:-}" in
build)
    shift
    mkdir -p "$ROOT/_variants"
    for spec in "$@"; do
        name=${spec%%=*}; defs=${spec#*=}
        ( cd "$ROOT" && TB_NVCC_EXTRA="$defs" python -m trio_binning_b200._build --force >/dev/null 2>&1 ) || { echo "build failed: $name"; exit 1; }
        LIB=$(ls "$ROOT"/trio_binning_b200/kmers_c*.so | head -1)
        cp "$LIB" "$ROOT/_variants/$name.so" && echo "built $name ($defs)"
    done
    ( cd "$ROOT" && python -m trio_binning_b200._build --force >/dev/null 2>&1 ) && echo "default library restored"
    ;;
run)
    cp "$LIB" /tmp/tb_orig.so
    for so in "$ROOT"/_variants/*.so; do
        v=$(basename "$so" .so)
        cp "$so" "$LIB"
        for kk in ${KS:-21 31}; do
            echo -n "$v k=$kk: "
            ( cd "$ROOT" && K=$kk timeout 300 python tools/kernel_variants.py ${SHAPE:-10000000 10000 20000} 2>&1 ) | grep -v "^Found\|^Done\|^Reading" |
                sed -E 's/.*(passes=[0-9.]+).*(step [0-9.]+ ms, classify [0-9.]+ ms, [0-9.]+ Gbases\/s).*(counts=.*)/\1 \2 \3/'
        done
    done
    cp /tmp/tb_orig.so "$LIB"
    ;;
clean)
    rm -rf "$ROOT/_variants"
    ;;
*)
    sed -n 2,14p "$0"
    ;;
esac
