for spec in "c2 26" "c2 27" "c2 28" "c2 29" "c2 30" "c3 30" "c3 31" "c3 32" "c5 27" "c5 28" "c5 29" "c5 30"; do set -- $spec; echo -n "$1 log2=$2: "; TB_FILTER_LOG2=$2 timeout 300 python tools/config_bench.py $1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    x=json.loads(l); print('pass',round(x['pass_fraction'],4),'ms',round(x['classify_ms'],3),round(x['gbases_per_s_kernel'],1))"; done
for m in 27 28 29 30; do echo -n "k31 10M+10M log2=$m: "; K=31 TB_FILTER_LOG2=$m timeout 300 python tools/kernel_variants.py 10000000 10000 20000 2>&1 | grep -o "passes=[0-9.]* .*classify [0-9.]* ms, [0-9.]* Gbases/s"; done
for m in 31 32 33; do echo -n "k31 150M+150M log2=$m: "; K=31 TB_FILTER_LOG2=$m timeout 300 python tools/kernel_variants.py 150000000 10000 25000 2>&1 | grep -o "passes=[0-9.]* .*classify [0-9.]* ms, [0-9.]* Gbases/s"; done
