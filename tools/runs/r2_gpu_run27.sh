#!/bin/bash
# m-mers per window (w = k - m + 1) at k = 21: w = 7 (default, m = 15) .. 11 (m = 11), prebuilt by
#   tools/variants.sh build w7="" w8="-DTB_MINI_W=8" w9="-DTB_MINI_W=9" w10="-DTB_MINI_W=10" w11="-DTB_MINI_W=11"
# Fewer minimizer changes per window = fewer bitmap sector fetches (what binds C3), more ALU work
# per window (what binds C2), and fewer distinct minimizers to spread the keys over the sectors.
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
LIB=$(ls "$ROOT"/trio_binning_b200/kmers_c*.so | head -1)
cp "$LIB" /tmp/tb_orig.so
run() { ( cd "$ROOT" && K=21 timeout 300 python tools/kernel_variants.py "$@" 2>&1 ) | grep -o "filter=2^[0-9]*\|passes=[0-9.]*\|classify [0-9.]* ms, [0-9.]* Gbases/s\|counts=.*" | tr '\n' ' '; echo; }
for v in w7 w8 w9 w10 w11; do
    cp "$ROOT/_variants/$v.so" "$LIB"
    echo -n "$v C2: "; run 10000000 10000 20000
    for log2 in 32 33; do
        if [ $log2 = 33 ] && { [ $v = w7 ] || [ $v = w8 ]; }; then continue; fi
        echo -n "$v C3 log2=$log2: "; TB_FILTER_LOG2=$log2 run 150000000 10000 25000
    done
done
cp /tmp/tb_orig.so "$LIB"
